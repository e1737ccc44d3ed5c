#!/usr/bin/env python
"""Throughput of the non-headline configurations of BASELINE.json (configs 1, 2, 3, 5) on one GPU,
timed with CUDA events, next to the numpy oracle on one host core (reported baseline, small sample).
Prints one JSON object per config.  bench.py remains the contractual benchmark (config 4)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
FS = 16000


def timed(fn, warm=2, reps=3):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    import torch
    from oracle import np_oracle as O
    from distantspeech_b200.beamformer.MicArray import MicArray
    from distantspeech_b200.beamformer.fixedbeamformer import FixedBeamformer
    from distantspeech_b200.beamformer.adaptivebeamformer import adaptivebeamfomer
    from distantspeech_b200.beamformer.FDGSC import FDGSC
    from distantspeech_b200.doa.srp import srp
    small = "--small" in sys.argv
    out = []
    N = 160000 // 256 * 256

    # ---- config 2: fixed SD beamformer, 8-mic circular, 1024 streams x 10 s -------------------
    S = 128 if small else 1024
    mic = MicArray(arrayType="circular", r=0.05, M=8, n_fft=512)
    x = torch.randn((S, N, 8), device="cuda") * 0.1
    fb = FixedBeamformer(mic, 512, 256, 512)
    W = fb.compute_weights([30, 0], "SD")[None]
    xs = x.permute(0, 2, 1).contiguous()
    def run2():
        fb.reset()
        t = torch
        import ctypes as C
        from distantspeech_b200 import _lib as L
        p = L.FixedBfParams(512, 256, S, 8, N, 1, 0, 0, float(256 / fb.transform.W0))
        if fb._state is None:
            fb._state = t.zeros(L.lib().ds_fixedbf_state_bytes(C.byref(p)), dtype=t.uint8, device="cuda")
        L.check(L.lib().ds_fixedbf_run(C.byref(p), L.ptr(L.device_window(fb.transform.window, 512)), L.ptr(Wd),
                                       L.ptr(fb._state), L.ptr(xs), L.ptr(y2), L.stream_ptr()), "fixedbf")
    from distantspeech_b200 import _lib as L
    Wd = L.to_device(np.asarray(W, dtype=np.complex64), torch.complex64)
    y2 = torch.empty((S, 1, N), dtype=torch.float32, device="cuda")
    ms = timed(run2)
    geo = O.MicGeometry("circular", r=0.05, M=8, n_fft=512)
    xh = x[0, :16000 * 2].cpu().numpy().astype(np.float64)
    t0 = time.perf_counter(); O.fixed_beamform(xh, W[0], 512, 256); cpu = 2.0 / (time.perf_counter() - t0)
    out.append({"config": 2, "what": "fixed SD beamformer 8-mic, %d streams x 10 s" % S, "ms": ms,
                "audio_s_per_s": S * N / FS / (ms / 1e3), "hbm_GBps_algorithmic": S * N * 36 / (ms / 1e3) / 1e9,
                "cpu_oracle_1core_audio_s_per_s": cpu})
    del x, xs, y2

    # ---- config 1: online MVDR, 4-mic, one 10 s utterance (and a batch) -------------------------
    mic4 = MicArray(arrayType="circular", r=0.032, M=4, n_fft=512)
    ang = np.array([30, 0]) / 180 * np.pi
    for S1 in (1, 256 if small else 2048):
        x1 = torch.randn((S1, 4, N), device="cuda") * 0.1
        ab = adaptivebeamfomer(mic4, 512, 256, 512)
        def run1():
            ab._state = None
            return ab.process(x1, ang, method=2)
        ms = timed(run1, warm=1, reps=2)
        out.append({"config": 1, "what": "adaptivebeamfomer.process MVDR 4-mic, %d stream(s) x 10 s (incl. API hand-off)" % S1,
                    "ms": ms, "audio_s_per_s": S1 * N / FS / (ms / 1e3)})
        del x1
    geo4 = O.MicGeometry("circular", r=0.032, M=4, n_fft=512)
    xh = (np.random.default_rng(0).standard_normal((4, 16000 * 2)) * 0.1)
    t0 = time.perf_counter(); O.adaptive_mvdr(xh, geo4, ang, 512, 256); cpu = 2.0 / (time.perf_counter() - t0)
    out[-1]["cpu_oracle_1core_audio_s_per_s"] = cpu

    # ---- config 3: FDGSC 6-mic linear, 4096 streams --------------------------------------------
    S3 = 256 if small else 4096
    mic6 = MicArray(arrayType="linear", r=0.05, M=6, n_fft=256)
    x3 = torch.randn((S3, N, 6), device="cuda") * 0.1
    fd = FDGSC(mic6, frameLen=256, angle=[90, 0])
    def run3():
        fd.reset()
        return fd.process(x3)
    ms = timed(run3, warm=1, reps=2)
    geo6 = O.MicGeometry("linear", r=0.05, M=6, n_fft=256)
    xh = x3[0, :256 * 125].cpu().numpy().astype(np.float64)
    t0 = time.perf_counter(); O.FdgscOracle(geo6, 256, np.array([90, 0]) / 180 * np.pi).process(xh); cpu = 2.0 / (time.perf_counter() - t0)
    out.append({"config": 3, "what": "FDGSC 6-mic linear, %d streams x 10 s (incl. API hand-off + diagnostics outputs)" % S3,
                "ms": ms, "audio_s_per_s": S3 * N / FS / (ms / 1e3), "cpu_oracle_1core_audio_s_per_s": cpu})
    del x3

    # ---- config 5: SRP-PHAT 16-mic 48 kHz, n_fft 1024, 360 x 90 grid ------------------------------
    mic16 = MicArray(arrayType="circular", r=0.05, M=16, n_fft=1024)
    mic16.fs = 48000
    mic16.omega = 2 * np.pi * mic16.freq_bin * mic16.fs / mic16.n_fft
    secs = 2.0 if small else 10.0
    N5 = int(48000 * secs) // 512 * 512
    x5 = (torch.randn((N5, 16), device="cuda") * 0.1)
    az, el = np.arange(360), np.arange(90)
    tau = np.stack([mic16.compute_tau(np.array([a, e]) * np.pi / 180)[:, 0] for a in az for e in el])
    for eng in ("tensor", "simt"):
        sp = srp(mic16, engine=eng)
        X = sp._spectrum(x5)
        def run5():
            return sp._steered_response(X, tau, True)
        ms = timed(run5, warm=1, reps=2)
        D, T, K = tau.shape[0], X.shape[0], X.shape[2]
        out.append({"config": 5, "what": "SRP-PHAT 16-mic 48 kHz n_fft 1024, %d dirs x %d frames (%s engine)" % (D, T, eng),
                    "ms": ms, "audio_s_per_s": (N5 / 48000) / (ms / 1e3), "tflops_algorithmic": 8.0 * D * 16 * K * T / (ms / 1e3) / 1e12})
    # ---- f1: frequency-domain GSC with the McMcra postfilter, 8 mics, n_fft 256 / hop 128 ----
    from distantspeech_b200.beamformer.GSC import GSC
    Sg = 128 if small else 1024
    mic8g = MicArray(arrayType="circular", r=0.05, M=8)
    xg = torch.randn((Sg, 8, N), device="cuda") * 0.1
    gsc = GSC(mic8g, 256)
    ms = timed(lambda: gsc.process(xg, ang, method=2), warm=1, reps=2)
    geo8 = O.MicGeometry("circular", r=0.05, M=8, n_fft=256)
    xh = xg[0, :, :16000].cpu().numpy().astype(np.float64)
    t0 = time.perf_counter(); O.GscOracle(geo8, 256).process(xh, ang, method=2); cpu = 1.0 / (time.perf_counter() - t0)
    out.append({"config": "f1", "what": "frequency-domain GSC + McMcra postfilter 8-mic, %d streams x 10 s (incl. API hand-off)" % Sg,
                "ms": ms, "audio_s_per_s": Sg * N / FS / (ms / 1e3), "cpu_oracle_1core_audio_s_per_s": cpu})
    # ---- f1: SubbandGSC (STFT-domain NLMS blocking filters + canceller gated by McSpp), 4 mics ----------------
    from distantspeech_b200.beamformer.SubbandGSC import SubbandGSC
    Ss = 128 if small else 1024
    xs4 = torch.randn((Ss, 4, N), device="cuda") * 0.1
    sg = SubbandGSC(mic4, 256, angle=[30, 0])
    ms = timed(lambda: sg.process(xs4), warm=1, reps=2)
    geo4 = O.MicGeometry("circular", r=0.032, M=4, n_fft=256)
    xh = xs4[0, :, :16000].cpu().numpy().astype(np.float64)
    t0 = time.perf_counter(); O.SubbandGscOracle(geo4, 256, ang).process(xh); cpu = 1.0 / (time.perf_counter() - t0)
    out.append({"config": "f1", "what": "SubbandGSC 4-mic, %d streams x 10 s (incl. API hand-off + diagnostics outputs)" % Ss,
                "ms": ms, "audio_s_per_s": Ss * N / FS / (ms / 1e3), "cpu_oracle_1core_audio_s_per_s": cpu})
    # ---- f1: TDGSC (fixed blocking matrix + constrained FDAF canceller), 4 mics -----------------------------------
    from distantspeech_b200.beamformer.TDGSC import TDGSC
    xt = torch.randn((Ss, N, 4), device="cuda") * 0.1
    td = TDGSC(mic4, frameLen=256, angle=[30, 0])
    ms = timed(lambda: td.process(xt), warm=1, reps=2)
    xh = xt[0, :16000].cpu().numpy().astype(np.float64)
    t0 = time.perf_counter(); O.TdgscOracle(geo4, 256, ang).process(xh); cpu = 1.0 / (time.perf_counter() - t0)
    out.append({"config": "f1", "what": "TDGSC 4-mic, %d streams x 10 s (incl. API hand-off + diagnostics outputs)" % Ss,
                "ms": ms, "audio_s_per_s": Ss * N / FS / (ms / 1e3), "cpu_oracle_1core_audio_s_per_s": cpu})
    # ---- f3: mask-based MVDR / GEV (mvdr.ipynb cells 6, 8), 8 mics, mask from McSppBase on the device --------
    from distantspeech_b200.pipelines import MaskBeamformer
    Sm = 64 if small else 256
    xm = torch.randn((Sm, 8, N), device="cuda") * 0.1
    for method in ("mvdr", "gev"):
        mb = MaskBeamformer(8, method=method)
        ms = timed(lambda: mb.process_device(xm), warm=1, reps=2)
        pm = mb.p
        ms_w = timed(lambda: mb.weights_device(mb.Phi_xx, mb.Phi_vv), warm=1, reps=3)
        row = {"config": "f3", "what": "mask-based %s 8-mic, %d streams x 10 s (STFT + McSppBase mask from the output-only kernel's p tap + covariances + "
                                      "weights + apply + ISTFT)" % (method.upper(), Sm),
               "ms": ms, "audio_s_per_s": Sm * N / FS / (ms / 1e3), "ms_weights_only": ms_w,
               "eigenproblems_per_s": Sm * 257 / (ms_w / 1e3)}
        if method == "mvdr":
            xh = xm[0, :, :16000].cpu().numpy().astype(np.float64).T
            t0 = time.perf_counter(); O.mask_beamform(xh, method="mvdr"); row["cpu_oracle_1core_audio_s_per_s"] = 1.0 / (time.perf_counter() - t0)
        out.append(row)
        del mb
    del xm
    torch.cuda.empty_cache()
    # ---- f4: Idoa spatial presence probability, 8 mics, n_fft 512 -------------------------------------------
    from distantspeech_b200.doa.idoa import Idoa
    Si = 128 if small else 1024
    xi8 = torch.randn((Si, N, 8), device="cuda") * 0.1
    ido = Idoa(mic)
    ms = timed(lambda: ido.process(xi8, theta=30), warm=1, reps=2)
    geo8i = O.MicGeometry("circular", r=0.05, M=8, n_fft=512)
    xh = xi8[0, :16000].cpu().numpy().astype(np.float64)
    oi = O.IdoaOracle(geo8i)
    t0 = time.perf_counter(); oi.process(xh, theta=30); cpu = 1.0 / (time.perf_counter() - t0)
    out.append({"config": "f4", "what": "Idoa.process 8-mic (one look direction), %d streams x 10 s (incl. API hand-off)" % Si,
                "ms": ms, "audio_s_per_s": Si * N / FS / (ms / 1e3), "cpu_oracle_1core_audio_s_per_s_all_360_directions": cpu})
    Xi = ido.transform.stft(xi8[0])                                                      # [K, T, M] device
    ido2 = Idoa(mic)
    ms = timed(lambda: ido2.estimate(Xi), warm=1, reps=2)
    out.append({"config": "f4", "what": "Idoa.estimate 8-mic, full 360-direction map of one 10 s stream (p [257, 625, 360] on the device)",
                "ms": ms, "audio_s_per_s": N / FS / (ms / 1e3), "bin_direction_frame_updates_per_s": 257 * 360 * (N // 256) / (ms / 1e3)})
    del xi8
    torch.cuda.empty_cache()
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
