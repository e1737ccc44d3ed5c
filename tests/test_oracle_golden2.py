"""CPU: oracle vs reference-generated goldens for the FDGSC / postfilter rows (a10, a15, a16, a17)."""
import numpy as np
import pytest

from conftest import golden, snr_db
from oracle import np_oracle as O


def test_fdgsc_golden():
    g = golden("fdgsc.npz")
    geo = O.MicGeometry("linear", r=0.05, M=6, n_fft=256)
    o = O.FdgscOracle(geo, 256, g["angle_deg"] / 180 * np.pi)
    assert np.allclose(o.h, g["delay_filter"], rtol=0, atol=1e-15)
    out, p, fix, bm, xn = o.process(g["x"].astype(np.float64))
    assert np.max(np.abs(out - g["y"])) < 1e-9 and snr_db(g["y"], out) > 150
    assert np.array_equal(xn.astype(np.float32), g["x_notched"])
    assert np.max(np.abs(p - g["p"])) < 1e-6
    assert np.max(np.abs(fix - g["fix_output"])) < 1e-6 and np.max(np.abs(bm - g["bm_output"])) < 1e-6
    assert np.allclose(o.aic.W, g["W_aic_last"], rtol=1e-6, atol=1e-9)


def test_fdgsc_postfilter_golden():
    g, gp = golden("fdgsc.npz"), golden("fdgsc_postfilter.npz")
    n1 = int(gp["n_first"])
    o = O.FdgscOracle(O.MicGeometry("linear", r=0.05, M=6, n_fft=256), 256, g["angle_deg"] / 180 * np.pi)
    ya = o.process(g["x"][:n1].astype(np.float64), postfilter=True)[0]
    yb = o.process(g["x"][n1:].astype(np.float64), postfilter=True)[0]
    assert np.max(np.abs(np.concatenate([ya, yb]) - gp["y"])) < 1e-9
    assert np.allclose(o.omlsa_multi.G, gp["G_last"], rtol=1e-9, atol=0)


def test_omlsa_multi_golden():
    g = golden("omlsa_multi.npz")
    Y, U = g["Y"].astype(np.float64), g["U"].astype(np.float64)
    o = O.OmlsaMulti(nfft=512, M=6, cal_weights=True)
    assert o.estimation(Y[:, 0], U[:, 0, :]) is None                  # first frame returns None (omlsa_multi.py:87-93)
    for n in range(1, Y.shape[1]):
        lam = o.estimation(Y[:, n], U[:, n, :])
        assert np.array_equal(o.G, g["G"][n]) and np.array_equal(o.p, g["p"][n]) and np.array_equal(lam, g["lambda_d"][n])


def test_zelinski_golden():
    g = golden("zelinski.npz")
    geo = O.MicGeometry("circular", r=0.05, M=8, n_fft=256)
    pf = O.ZelinskiPostFilter(8, 129, O.noise_msc(geo, 256))
    for n in range(g["Z"].shape[2]):
        W = pf.getweights(g["Z"][:, :, n].astype(complex))
        assert np.allclose(W, g["W"][n], rtol=1e-12, atol=0)
    assert np.allclose(pf.Pxii, g["Pxii"], rtol=1e-13) and np.allclose(pf.Pxij, g["Pxij"], rtol=1e-13, atol=1e-30)


def test_mcspp_cdr_golden():
    """a14: McSpp with the McCDR prior, 4 microphones, 240 frames (MCRA of the prior leaves its
    2 x 65-frame warm-up) -- the restatement reproduces the reference to the last bit."""
    g = golden("mcspp_cdr.npz")
    D = O.Transform(channel=4, n_fft=512, hop_length=256).stft(g["x"].astype(np.float64))      # [257, 240, 4]
    est = O.McSpp(nfft=512, channels=4)
    for n in range(D.shape[1]):
        p = est.estimation(D[:, n, :])
        assert np.array_equal(p, g["p"][:, n]) and np.array_equal(est.xi, g["xi"][:, n])
        assert np.array_equal(est.gamma, g["gamma"][:, n]) and np.array_equal(est.q, g["q"][:, n])
    assert np.array_equal(est.w, g["w_last"]) and np.array_equal(est.Phi_vv, g["Phi_vv_last"])
    assert np.array_equal(est.Phi_vv_inv, g["Phi_vv_inv_last"]) and np.array_equal(est.Phi_xx, g["Phi_xx_last"])
    assert np.array_equal(est.mccdr.mcra.p, g["mcra_p_last"])
    with pytest.raises(ValueError):
        O.McSpp(nfft=512, channels=3)                     # pair (1, 2) of the CDR prior is undefined below 4 channels


def test_gsc_mcmcra_golden():
    """f1: frequency-domain GSC with the McMcra postfilter, two consecutive calls (streaming state)."""
    g = golden("gsc.npz")
    geo = O.MicGeometry("circular", r=0.032, M=4, n_fft=256)
    o = O.GscOracle(geo, 256)
    n1 = int(g["n_first"])
    x = g["x"].astype(np.float64)
    y = np.concatenate([o.process(x[:, :n1], g["angle_rad"], method=2), o.process(x[:, n1:], g["angle_rad"], method=2)])
    assert np.max(np.abs(y - g["y"])) < 1e-12
    assert np.array_equal(o.spp.p, g["p_last"]) and np.array_equal(o.spp.q, g["q_last"])
    assert np.allclose(o.spp.G, g["Gpost_last"], rtol=1e-13, atol=0) and np.allclose(o.Gw, g["Gw_last"], rtol=1e-12, atol=1e-15)
    assert np.allclose(o.spp.Phi_vv, g["Phi_vv_last"], rtol=1e-13, atol=1e-18)


def test_subband_gsc_golden():
    """f1: SubbandGSC (STFT-domain NLMS blocking filters + canceller gated by McSpp), two calls with a ragged tail."""
    g = golden("subband_gsc.npz")
    o = O.SubbandGscOracle(O.MicGeometry("circular", r=0.032, M=4, n_fft=256), 256, np.array([30, 0]) / 180 * np.pi)
    n1 = int(g["n_first"])
    x = g["x"].astype(np.float64)
    a, b = o.process(x[:, :n1]), o.process(x[:, n1:])
    assert np.max(np.abs(np.concatenate([a[0], b[0]]) - g["y"])) < 1e-12
    assert np.max(np.abs(np.concatenate([a[2], b[2]]) - g["bm_output"])) < 1e-6
    assert np.allclose(o.aic.W, g["W_aic_last"], rtol=1e-9, atol=1e-14) and np.allclose(o.bm[0].W[:, :, 0], g["W_bm0_last"], rtol=1e-9, atol=1e-14)


def test_subband_rls_golden():
    g = golden("subband_rls.npz")
    o = O.SubbandRls(2, 512)
    x, d = g["x"].astype(np.float64), g["d"].astype(np.float64)
    err = np.concatenate([o.update(x[256 * n:256 * (n + 1)], d[256 * n:256 * (n + 1)]) for n in range(14)])
    assert np.array_equal(err, g["err"]) and np.array_equal(o.W, g["W_last"]) and np.array_equal(o.P, g["P_last"])


def test_tdgsc_postfilter_golden():
    # TDGSC.process(postfilter=True) (TDGSC.py:157-170): NsOmlsaMulti gain on the output, two calls on one object
    g = golden("tdgsc_postfilter.npz")
    o = O.TdgscOracle(O.MicGeometry("circular", r=0.032, M=4, n_fft=256), 256, np.array([30, 0]) / 180 * np.pi)
    n1 = int(g["n_first"])
    x = g["x"].astype(np.float64)
    a, b = o.process(x[:n1], postfilter=True), o.process(x[n1:], postfilter=True)
    assert np.max(np.abs(np.concatenate([a[0], b[0]]) - g["y"])) < 1e-12
    assert np.allclose(o.omlsa_multi.G, g["G_last"], rtol=1e-9) and np.allclose(o.omlsa_multi.lambda_d, g["lambda_last"], rtol=1e-9)


def test_tdgsc_golden():
    g = golden("tdgsc.npz")
    o = O.TdgscOracle(O.MicGeometry("circular", r=0.032, M=4, n_fft=256), 256, np.array([30, 0]) / 180 * np.pi)
    n1 = int(g["n_first"])
    x = g["x"].astype(np.float64)
    a, b = o.process(x[:n1]), o.process(x[n1:])
    assert np.max(np.abs(np.concatenate([a[0], b[0]]) - g["y"])) < 1e-12
    assert np.allclose(o.W, g["W_last"], rtol=1e-9, atol=1e-14)
