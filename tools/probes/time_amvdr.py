#!/usr/bin/env python
"""Config-1 step (adaptivebeamfomer.process, 4-mic linear array, S streams x 10 s) for A/B timing and as the ncu target."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from distantspeech_b200.beamformer.MicArray import MicArray
from distantspeech_b200.beamformer.adaptivebeamformer import adaptivebeamfomer
S = int(os.environ.get("S", 2048)); N = 256 * 625
mic = MicArray(arrayType="linear", r=0.032, M=4, n_fft=512)
ab = adaptivebeamfomer(mic, 512, 256, 512)
g = torch.Generator(device="cuda"); g.manual_seed(1)
x = torch.randn((S, 4, N), device="cuda", generator=g) * 0.1
ang = np.array([30, 0]) / 180 * np.pi
ts = []
for it in range(4):
    ab._state = None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); y = ab.process(x, ang, method=2)["data"]; e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print(os.environ.get("DS_B200_LIB", "default"), "ms:", " ".join("%.2f" % t for t in ts))
