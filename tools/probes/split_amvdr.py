#!/usr/bin/env python
"""Where the config-1 step goes: the stages of adaptivebeamfomer.process timed one by one (CUDA events + wall clock)."""
import os, sys, time, ctypes as C, numpy as np, torch as t
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from distantspeech_b200 import _lib as L
from distantspeech_b200.beamformer.MicArray import MicArray
from distantspeech_b200.beamformer.adaptivebeamformer import adaptivebeamfomer
from distantspeech_b200.transform.transform import stft_device, istft_device
S = int(os.environ.get("S", 2048)); N = 256 * 625
mic = MicArray(arrayType="linear", r=0.032, M=4, n_fft=512)
ab = adaptivebeamfomer(mic, 512, 256, 512)
x = t.randn((S, 4, N), device="cuda") * 0.1
ang = np.array([30, 0]) / 180 * np.pi
ab.process(x, ang, method=2); t.cuda.synchronize()
for it in range(2):
    ab._state = None
    ab._ensure(S)
    ev = [t.cuda.Event(enable_timing=True) for _ in range(6)]
    w0 = time.perf_counter()
    win = L.device_window(ab.transformer.window, ab.nfft)
    a_dev = t.ones((4, 257), dtype=t.complex128, device="cuda")
    ev[0].record()
    X = stft_device(x, ab.nfft, ab.hop, win, L.DS_STFT_STREAMING, history=ab._hist)
    ev[1].record()
    T = X.shape[1]
    Y = t.empty((S, T, 1, 257), dtype=t.complex64, device="cuda")
    Hl = t.empty((S, 257, 4), dtype=t.complex128, device="cuda")
    pl = t.empty((S, T, 257), dtype=t.float64, device="cuda")
    prm = ab._params(S, T, 2)
    ev[2].record()
    L.check(L.lib().ds_amvdr_run(C.byref(prm), L.ptr(ab._state), L.ptr(a_dev), L.ptr(X), L.ptr(Y), L.ptr(Hl), L.ptr(pl), L.stream_ptr()))
    ev[3].record()
    y = istft_device(Y, ab.nfft, ab.hop, win, L.DS_STFT_STREAMING, tail=ab._tail, scale=1.0)[:, 0, :]
    ev[4].record()
    Hn = Hl.cpu().numpy(); pn = pl[:, -1, :].cpu().numpy()
    ev[5].record(); t.cuda.synchronize()
    w1 = time.perf_counter()
    print("stft %.2f | alloc %.2f | amvdr %.2f | istft %.2f | d2h %.2f | wall %.2f ms" % (
        ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3]), ev[3].elapsed_time(ev[4]),
        ev[4].elapsed_time(ev[5]), (w1 - w0) * 1e3))
    # the same kernel without the per-frame p tap
    e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
    ab._state.zero_()
    e0.record()
    L.check(L.lib().ds_amvdr_run(C.byref(prm), L.ptr(ab._state), L.ptr(a_dev), L.ptr(X), L.ptr(Y), L.ptr(Hl), None, L.stream_ptr()))
    e1.record(); t.cuda.synchronize()
    print("amvdr without p_out: %.2f ms" % e0.elapsed_time(e1))
