#!/usr/bin/env python
"""Config-2 step (ds_fixedbf_run, 8 mics, S streams x 10 s, one SD beam) for A/B timing and as the ncu target."""
import os, sys, ctypes, numpy as np, torch as t
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from distantspeech_b200 import _lib as L
from distantspeech_b200.beamformer.MicArray import MicArray
from distantspeech_b200.beamformer.fixedbeamformer import FixedBeamformer
S = int(os.environ.get("S", 1024)); N = 256 * 625; M = 8
mic = MicArray(arrayType="circular", r=0.05, M=M, n_fft=512)
fb = FixedBeamformer(mic, 512, 256, 512)
W = fb.compute_weights([30, 0], "SD")[None]
Wd = L.to_device(np.asarray(W, dtype=np.complex64), t.complex64)
p = L.FixedBfParams(512, 256, S, M, N, 1, 0, 0, float(256 / fb.transform.W0))
state = t.zeros(L.lib().ds_fixedbf_state_bytes(ctypes.byref(p)), dtype=t.uint8, device="cuda")
x = t.randn((S, M, N), device="cuda") * 0.1
y = t.empty((S, 1, N), dtype=t.float32, device="cuda")
win = L.device_window(fb.transform.window, 512)
ts = []
for it in range(5):
    state.zero_()
    e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
    e0.record()
    L.check(L.lib().ds_fixedbf_run(ctypes.byref(p), L.ptr(win), L.ptr(Wd), L.ptr(state), L.ptr(x), L.ptr(y), L.stream_ptr()))
    e1.record(); t.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print(os.environ.get("DS_B200_LIB", "default"), "ms:", " ".join("%.3f" % v for v in ts))
