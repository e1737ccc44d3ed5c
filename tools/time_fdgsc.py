#!/usr/bin/env python
"""Times FDGSC.process_device (config 3: S streams x 10 s x 6 mics) call by call with CUDA events (no NVML polling),
both implementations; S via the environment."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distantspeech_b200.beamformer.MicArray import MicArray
from distantspeech_b200.beamformer.FDGSC import FDGSC
S = int(os.environ.get("S", 4096)); N = 256 * 625
mic = MicArray(arrayType="linear", r=0.05, M=6, n_fft=256)
g = torch.Generator(device="cuda"); g.manual_seed(1)
x = torch.randn((S, 6, N), device="cuda", generator=g) * 0.1
xw = torch.empty_like(x); y = torch.empty((S, N), device="cuda")
for impl in (os.environ.get("IMPLS", "pipeline,fused").split(",")):
    fd = FDGSC(mic, frameLen=256, angle=[90, 0], impl=impl)
    ts, tc = [], []
    for it in range(4):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(); xw.copy_(x); e1.record()
        fd.reset_state(); fd.process_device(xw, out=y); e2.record(); torch.cuda.synchronize()
        tc.append(e0.elapsed_time(e1)); ts.append(e1.elapsed_time(e2))
    print(impl, "copy ms:", " ".join("%.2f" % t for t in tc), "| process ms:", " ".join("%.2f" % t for t in ts), flush=True)
