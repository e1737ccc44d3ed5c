#!/usr/bin/env python
"""CPU study behind the mixed-precision per-bin chain kernel: which parts of the McSppBase + MVDR +
OMLSA frame body tolerate float32 arithmetic.  Emulates in NumPy the kernel's formulation (xi / gamma through
A (Phi_vv + eps I) = I, packed real parts) with selected intermediates rounded to / evaluated in float32, and reports
the output SNR against the float64 oracle -- on the synthetic streams of SURVEY.md 8d, on a nearly noise-free mixture
(ill-conditioned Phi_vv) and, where the reference tree is present, on the shipped recording example/test_audio/rec1.

    python tools/mixed_precision_study.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import np_oracle as O  # noqa: E402

f32 = np.float32


def chain(D, a0, mode):
    """D [K, T, M] complex spectrum (complex64-representable), a0 [K, M].  mode: set of strings out of
    {"u", "z", "tr", "mvdr"}: which parts run in float32.  Returns Y [K, T] and p [K, T]."""
    K, T, M = D.shape
    eps, alpha, alpha_d = 1e-6, 0.92, 0.92
    Pyy = np.zeros((K, M, M))
    Pvv = np.zeros((K, M, M))
    mcra = O.Mcra(nfft=2 * (K - 1), L=15)
    Y = np.zeros((K, T), dtype=complex)
    P = np.zeros((K, T))
    I = np.eye(M)
    for n in range(T):
        y = D[:, n, :]
        A = np.linalg.inv(Pvv + eps * I)
        A32 = A.astype(f32)
        yr, yi = y.real, y.imag
        if "u" in mode:
            ur = np.einsum("kij,kj->ki", A32, yr.astype(f32)).astype(np.float64)
            ui = np.einsum("kij,kj->ki", A32, yi.astype(f32)).astype(np.float64)
        else:
            ur = np.einsum("kij,kj->ki", A, yr)
            ui = np.einsum("kij,kj->ki", A, yi)
        Pyy = alpha * Pyy + (1 - alpha) * (yr[:, :, None] * yr[:, None, :] + yi[:, :, None] * yi[:, None, :])
        if "tr" in mode:
            tr = np.einsum("kij,kij->k", A32, Pyy.astype(f32)).astype(np.float64)
        else:
            tr = np.einsum("kij,kij->k", A, Pyy)
        trA = np.trace(A, axis1=1, axis2=2)
        xi = tr - M + eps * trA
        if "z" in mode:
            u32r, u32i = ur.astype(f32), ui.astype(f32)
            Z = u32r[:, :, None] * u32r[:, None, :] + u32i[:, :, None] * u32i[:, None, :]
            g1 = np.einsum("kij,kij->k", Pyy.astype(f32), Z)
            syu = np.sum(yr.astype(f32) * u32r + yi.astype(f32) * u32i, axis=1, dtype=f32)
            uu = np.sum(u32r * u32r + u32i * u32i, axis=1, dtype=f32)
            gam = ((g1 - syu) + f32(eps) * uu).astype(np.float64)
        else:
            Z = ur[:, :, None] * ur[:, None, :] + ui[:, :, None] * ui[:, None, :]
            gam = np.einsum("kij,kij->k", Pyy, Z) - np.sum(yr * ur + yi * ui, axis=1) + eps * np.sum(ur * ur + ui * ui, axis=1)
        xi = np.clip(xi, 1e-6, 1e6)
        gam = np.clip(gam, 1e-6, 1e6)
        y0 = y[:, 0]
        mcra.estimation(np.abs(y0 * np.conj(y0)))
        q = np.clip(np.sqrt(1 - mcra.p), 0.01, 0.99)
        p = np.clip(1 / (1 + q / (1 - q) * (1 + xi) * np.exp(-gam / (1 + xi))), 0.01, 0.99)
        G = np.clip(np.power(xi / (1 + xi), p) * np.power(0.0631, 1 - p), 0.0631, 1.0)
        G[:2] = 0
        if "mvdr" in mode:
            a32r, a32i = a0.real.astype(f32), a0.imag.astype(f32)
            C = a32r[:, :, None] * a32r[:, None, :] + a32i[:, :, None] * a32i[:, None, :]
            den = np.einsum("kij,kij->k", A32, C).astype(np.float64)
            u32r, u32i = ur.astype(f32), ui.astype(f32)
            nr = np.sum(a32r * u32r + a32i * u32i, axis=1, dtype=f32).astype(np.float64)
            ni = np.sum(a32r * u32i - a32i * u32r, axis=1, dtype=f32).astype(np.float64)
            num = nr + 1j * ni
        else:
            den = np.einsum("ki,kij,kj->k", np.conj(a0), A, a0).real
            num = np.sum(np.conj(a0) * (ur + 1j * ui), axis=1)
        Y[:, n] = num / den * G
        at = (alpha_d + (1 - alpha_d) * p)[:, None, None]
        Pvv = at * Pvv + (1 - at) * (yr[:, :, None] * yr[:, None, :] + yi[:, :, None] * yi[:, None, :])
        P[:, n] = p
    return Y, P


def snr(ref, x):
    return 10 * np.log10(np.sum(np.abs(ref) ** 2) / max(np.sum(np.abs(ref - x) ** 2), 1e-300))


def study(name, x, geo, look, n_fft, hop):
    M = x.shape[1]
    D = O.Transform(n_fft=n_fft, hop_length=hop, channel=M).stft(x)
    a0 = O.steering_from_doa(geo, n_fft, look)
    tf = lambda Yk: O.Transform(n_fft=n_fft, hop_length=hop, channel=1).istft(Yk[:, :, None])      # noqa: E731
    y_ref = O.mvdr_mcspp_chain(x, geo, look, n_fft, hop)
    Y0, P0 = chain(D, a0, set())
    print("%s: kernel formulation in float64 vs oracle: %.1f dB" % (name, snr(y_ref, tf(Y0))))
    for mode in (("mvdr",), ("z",), ("tr",), ("z", "tr"), ("u",), ("u", "z", "tr"), ("u", "z", "tr", "mvdr")):
        Y1, P1 = chain(D, a0, set(mode))
        y1 = tf(Y1)
        print("   float32 in %-22s SNR %.1f dB   max-abs %.2e   max|dp| %.2e" % ("+".join(mode), snr(y_ref, y1),
                                                                                     np.max(np.abs(y_ref - y1)), np.max(np.abs(P1 - P0))))


if __name__ == "__main__":
    geo = O.MicGeometry("circular", r=0.05, M=8, n_fft=512)
    x = O.synth_streams(1, geo, 256 * 250)[0].T.astype(np.float64)
    study("synthetic (8d recipe), 4 s", x, geo, (30, 0), 512, 256)
    # nearly noise-free: sensor noise 1e-4 -> Phi_vv close to rank 1
    rng = np.random.default_rng(5)
    xs = O.synth_streams(1, geo, 256 * 250, seed0=9)[0].T.astype(np.float64)
    clean = xs - 0.0                                                       # same recipe, then shrink the sensor-noise part
    x2 = O.synth_streams(1, geo, 256 * 250, seed0=9)[0].T.astype(np.float64)
    tau = O.compute_tau(geo, np.array([200.0, 0]) / 180 * np.pi)[:, 0]
    itf = rng.standard_normal(256 * 250 + 64)
    F = np.fft.rfft(itf, 1 << 17)
    f = np.fft.rfftfreq(1 << 17, 1 / 16000)
    x3 = np.stack([np.fft.irfft(F * np.exp(-2j * np.pi * f * t), 1 << 17)[:256 * 250] for t in tau], axis=1) * 0.1
    x3 += 1e-4 * rng.standard_normal(x3.shape)
    x3 = x3.astype(f32).astype(np.float64)
    study("single coherent interferer + 1e-4 sensor noise (ill-conditioned Phi_vv)", x3, geo, (30, 0), 512, 256)
    rec = os.path.join(ROOT, "tests", "golden", "adaptive_mvdr_rec1.npz")
    if os.path.exists(rec):
        pcm = np.load(rec)["pcm"][:, :256 * 250]
        xr = (pcm.astype(f32) / f32(32767)).astype(np.float64).T
        geo4 = O.MicGeometry("circular", r=0.032, M=4, n_fft=512)
        study("recording rec1 (4 mics), 4 s", xr, geo4, (197, 0), 512, 256)
