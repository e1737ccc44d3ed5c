#!/usr/bin/env python
"""Times the multi-beam fixed beamformer (8 mics, n_fft 512 / hop 256, S streams x 10 s, B beams): tensor-core path
(STFT + pack + tcgen05 GEMMs + per-beam ISTFT) against the fused CUDA-core kernel; kernel split with CUDA events."""
import os, sys, ctypes as C, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distantspeech_b200 import _lib as L
from distantspeech_b200.beamformer.MicArray import MicArray
from distantspeech_b200.beamformer.fixedbeamformer import FixedBeamformer
S = int(os.environ.get("S", 32)); B = int(os.environ.get("B", 64)); N = 256 * 625
mic = MicArray(arrayType="circular", r=0.05, M=8, n_fft=512)
x = torch.randn((S, N, 8), device="cuda") * 0.1
angles = [(int(a), 0) for a in np.linspace(0, 359, B)]
for eng in os.environ.get("ENGINES", "tensor").split(","):
    fb = FixedBeamformer(mic, 512, 256, 512)
    ts = []
    for it in range(4):
        fb.reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); y = fb.process_multibeam(x, angles, engine=eng); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("S=%d B=%d engine=%s ms: %s | %.3g beam-audio-s/s" % (S, B, eng, " ".join("%.2f" % t for t in ts), S * B * (N / 16000) / (min(ts) / 1e3)), flush=True)
    del y
