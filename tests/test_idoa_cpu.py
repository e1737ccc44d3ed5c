"""CPU: SURVEY 8f.4 -- the oracle's Idoa restatement against the golden fixture produced by the unmodified reference
(tests/golden/make_golden.py: idoa)."""
import numpy as np
import pytest

from conftest import golden, snr_db
from oracle import np_oracle as O

CASES = {"c4": ("circular", 4, 0.032, 256), "l6": ("linear", 6, 0.05, 512)}


@pytest.mark.parametrize("tag", ["c4", "l6"])
def test_idoa_oracle_golden(tag):
    g = golden("idoa.npz")
    arr, M, r, n_fft = CASES[tag]
    geo = O.MicGeometry(arr, r=r, M=M, n_fft=n_fft)
    x = g[tag + "_x"].astype(np.float64)
    n1 = int(g[tag + "_n1"])
    sel = g[tag + "_sel"]
    a = O.IdoaOracle(geo)
    assert np.allclose(a.Psi[:, :, sel], g[tag + "_Psi_sel"], rtol=0, atol=1e-12)
    y = np.concatenate([a.process(x[:n1], default_direction=30), a.process(x[n1:], default_direction=30)])
    assert np.max(np.abs(y - g[tag + "_y"])) < 1e-9 and snr_db(g[tag + "_y"], y) > 150
    b = O.IdoaOracle(geo)
    X = b.transform.stft(x)
    p = b.estimate(X)
    assert np.allclose(p[:, :, sel], g[tag + "_p_sel"], rtol=0, atol=1e-9, equal_nan=True)
    assert np.allclose(b.mu_Delta[:, sel], g[tag + "_mu_Delta_last"], rtol=0, atol=1e-9, equal_nan=True)
    assert np.allclose(b.var_Delta_h0[:, sel], g[tag + "_var_last"], rtol=0, atol=1e-9, equal_nan=True)
    c = O.IdoaOracle(geo)
    p1 = c.estimate(X, theta=40)
    assert np.allclose(p1[:, :, [40, 41]], g[tag + "_p_theta40"], rtol=0, atol=1e-9, equal_nan=True)


@pytest.mark.parametrize("tag", ["c4", "l6"])
def test_idoa_host_class_rtf_table(tag):
    """host logic of the drop-in class (no device work): direction grid and free-field RTFs Psi (idoa.py:41-44, 73-76)"""
    from distantspeech_b200.beamformer.MicArray import MicArray
    from distantspeech_b200.doa.idoa import Idoa
    g = golden("idoa.npz")
    arr, M, r, n_fft = CASES[tag]
    a = Idoa(MicArray(arrayType=arr, r=r, M=M, n_fft=n_fft))
    assert a.n_theta == (360 if arr == "circular" else 180) and a.idoa_dim == M - 1 and a.half_bin == n_fft // 2 + 1
    assert np.allclose(a.Psi[:, :, g[tag + "_sel"]], g[tag + "_Psi_sel"], rtol=0, atol=1e-12)
    assert a.p is None and a.mu_Delta is None                       # no state before the first call
    with pytest.raises(NotImplementedError):
        a.process(np.zeros((n_fft * 4, M)), pre_emphsis=True)
