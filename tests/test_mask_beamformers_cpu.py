"""CPU: SURVEY 8f.3 -- oracle restatement of steering / GEV / phase_correction / BAN / mask-based
beamformers against the golden fixture produced by the unmodified reference
(tests/golden/make_golden.py: mask_beamformers), and the host-compilable eigen-solver core that the
eig.cu kernels run per thread (csrc/eig_core.cuh) against LAPACK."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import golden, snr_db, ROOT
from oracle import np_oracle as O


def _same_up_to_sign(a, b, tol):
    """rows of a and b equal up to a factor +-1 per row"""
    s = np.sign(np.real(np.sum(a * np.conj(b), axis=-1, keepdims=True)))
    return np.max(np.abs(a - s * b)) <= tol * max(1.0, np.max(np.abs(b)))


def test_oracle_mask_beamformers_golden():
    g = golden("mask_beamformers.npz")
    x = g["x"].astype(np.float64)
    D = O.Transform(n_fft=512, hop_length=256, channel=6).stft(x)
    Pxx, Pvv = O.masked_covariances(D, g["p"])
    assert np.allclose(Pxx, g["Pxx"], rtol=1e-13, atol=1e-13) and np.allclose(Pvv, g["Pvv"], rtol=1e-13, atol=1e-13)
    assert np.allclose(O.steering_pca(g["Pxx"]), g["steer"], rtol=0, atol=1e-12)
    assert np.allclose(O.steering_pca(g["Ryy"] - g["Rvv"]), g["steer_pca"], rtol=0, atol=1e-12)
    assert np.allclose(O.mask_beamformer_weights(g["Pxx"], g["Pvv"], "mvdr"), g["w_mvdr"], rtol=1e-10, atol=1e-12)
    raw = O.gev_vector(g["Pxx"], g["Pvv"])
    assert np.allclose(raw, g["w_gev_raw"], rtol=1e-10, atol=1e-12)
    assert np.allclose(O.phase_correction(g["w_gev_raw"]), g["w_gev_pc"], rtol=1e-12, atol=1e-14)
    assert np.allclose(O.blind_analytic_normalization(g["w_gev_pc"], g["Pvv"]), g["w_gev"], rtol=1e-12, atol=1e-14)
    taps = {}
    y = O.mask_beamform(x, method="mvdr", taps=taps)
    assert np.max(np.abs(taps["p"] - g["p"])) < 1e-9                 # the oracle's McSppBase gives the reference's mask
    assert np.max(np.abs(y - g["y_mvdr"])) < 1e-7 and snr_db(g["y_mvdr"], y) > 120
    y = O.mask_beamform(x, p=g["p"], method="gev")
    assert np.max(np.abs(y - g["y_gev"])) < 1e-7 and snr_db(g["y_gev"], y) > 120


@pytest.fixture(scope="module")
def eig_host(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("eig") / "eig_host.so")
    src = os.path.join(ROOT, "tests", "host", "eig_host_shim.cpp")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", src, "-o", so], check=True)
    return ctypes.CDLL(so)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.mark.parametrize("M", [2, 3, 4, 6, 8])
def test_eig_core_matches_lapack(eig_host, M):
    import scipy.linalg as sl
    rng = np.random.default_rng(M)
    for trial in range(50):
        X = rng.standard_normal((M, M + 2)) + 1j * rng.standard_normal((M, M + 2))
        H = X @ X.conj().T
        N = rng.standard_normal((M, 3 * M)) + 1j * rng.standard_normal((M, 3 * M))
        B = N @ N.conj().T
        out = np.zeros(M, complex)
        assert eig_host.eig_host_steering(M, _p(H), _p(out)) <= 12
        assert np.max(np.abs(out - O.steering_pca(H))) < 1e-12
        assert eig_host.eig_host_gev(M, _p(H), _p(B), _p(out)) == 0
        ref = sl.eigh(H, B)[1][:, -1]
        assert _same_up_to_sign(out[None], ref[None], 1e-11)
        assert abs(np.vdot(out, B @ out).real - 1) < 1e-12
    # only the lower triangle is read, like LAPACK with UPLO='L'
    G = H.copy()
    G[np.triu_indices(M, 1)] = 123.0
    eig_host.eig_host_steering(M, _p(G), _p(out))
    assert np.max(np.abs(out - O.steering_pca(H))) < 1e-12
    # a noise matrix that is not positive definite is reported
    assert eig_host.eig_host_gev(M, _p(H), _p(-B), _p(out)) == 1


def test_eig_core_golden_matrices(eig_host):
    """the fixture's 257 6x6 covariance pairs (ill-conditioned low bins included)"""
    g = golden("mask_beamformers.npz")
    out = np.zeros(6, complex)
    for k in range(g["Pxx"].shape[0]):
        A, B = np.ascontiguousarray(g["Pxx"][k]), np.ascontiguousarray(g["Pvv"][k])
        eig_host.eig_host_steering(6, _p(A), _p(out))
        assert np.max(np.abs(out - g["steer"][k])) < 1e-9, k
        assert eig_host.eig_host_gev(6, _p(A), _p(B), _p(out)) == 0
        assert _same_up_to_sign(out[None], g["w_gev_raw"][k][None], 1e-8), k


def test_mask_beamformer_host_logic():
    """constructor-level validation of the pipeline class and the loud failure without a device (no CPU fallback)"""
    import torch
    from distantspeech_b200 import _lib
    from distantspeech_b200.pipelines import MaskBeamformer
    from distantspeech_b200.beamformer.beamformer import steering
    with pytest.raises(ValueError):
        MaskBeamformer(8, method="lcmv")
    bf = MaskBeamformer(6, n_fft=512, hop=256, method="gev")
    assert bf.K == 257 and abs(bf.W0 - 256.0) < 1e-9 and bf.w is None
    if not torch.cuda.is_available():
        with pytest.raises(_lib.DsError):
            steering(np.eye(4, dtype=complex)[None])
        with pytest.raises(_lib.DsError):
            bf.process(np.zeros((256 * 4, 6), np.float32))
