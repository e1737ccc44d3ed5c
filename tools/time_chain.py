#!/usr/bin/env python
"""Times MvdrMcsppChain.process_device (config 4: S streams x 10 s x 8 mics) with CUDA events; the small driver used
for A/B runs of kernel variants (DS_B200_LIB=build/variants/x.so) and as the ncu target.  S via the environment."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distantspeech_b200.beamformer.MicArray import MicArray
from distantspeech_b200.pipelines import MvdrMcsppChain
S = int(os.environ.get("S", 1024)); N = 256 * 625
mic = MicArray(arrayType="circular", r=0.05, M=8, n_fft=512)
ch = MvdrMcsppChain(mic, look_angle=(30, 0))
g = torch.Generator(device="cuda"); g.manual_seed(1)
x = torch.randn((S, 8, N), device="cuda", generator=g) * 0.1
y = torch.empty((S, N), device="cuda")
for _ in range(2):
    ch.reset_counters(); ch._state = None
    ch.process_device(x, out=y)
torch.cuda.synchronize()
ts = []
for _ in range(4):
    ch.reset_counters(); ch._state.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ch.process_device(x, out=y); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print(os.environ.get("DS_B200_LIB", "default"), os.environ.get("PRECISION", "f64"), "ms:", " ".join("%.2f" % t for t in ts), "finite" if torch.isfinite(y).all().item() else "nonfinite")
