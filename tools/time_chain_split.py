#!/usr/bin/env python
"""Kernel split of the config-4 step (analysis / per-bin / synthesis ms, CUDA events inside ds_chain_run_profiled) for A/B
runs of kernel variants (DS_B200_LIB=build/variants/x.so).  S via the environment."""
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
rows = []
for _ in range(4):
    ch.reset_counters(); ch._state.zero_()
    _, ms = ch.process_device_profiled(x, out=y)
    rows.append(ms)
print(os.environ.get("DS_B200_LIB", "default"), "analysis/perbin/synthesis ms:", " | ".join("%.3f %.3f %.3f" % tuple(r) for r in rows))
