"""CPU: the numpy oracle reproduces the golden vectors dumped from the reference."""
import numpy as np

from conftest import golden, snr_db
from oracle import np_oracle as O


def test_stft_istft_golden():
    g = golden("stft_istft.npz")
    x = g["x"].astype(np.float64)
    D = O.stft(x, n_fft=512, hop_length=128, window=O.sqrt_hann(512), center=True)
    assert D.dtype == np.complex64 and np.array_equal(D, g["D_512_128_center"])
    y = O.istft(D, hop_length=128, window=O.sqrt_hann(512), center=True, length=6000)
    assert y.dtype == np.float32 and np.array_equal(y, g["y_512_128_center"])
    D2 = O.stft(x, n_fft=256, hop_length=128, window=O.sqrt_hann(256), center=False)
    assert np.array_equal(D2, g["D_256_128_plain"])
    assert np.array_equal(O.istft(D2, hop_length=128, window=O.sqrt_hann(256), center=False), g["y_256_128_plain"])


def test_stft_window_must_be_array():
    import pytest
    with pytest.raises(ValueError):
        O.stft(np.zeros(1024), n_fft=256, window=None)


def test_transform_streaming_golden():
    g = golden("transform_stream.npz")
    x = g["x"].astype(np.float64)
    tf = O.Transform(n_fft=512, hop_length=256, channel=3)
    tf2 = O.Transform(n_fft=512, hop_length=256, channel=3)
    cuts = [(0, 256 * 5), (256 * 5, 256 * 6), (256 * 6, 256 * 12)]
    for i, (a, b) in enumerate(cuts):
        Y = tf.stft(x[a:b])
        assert np.array_equal(Y, g["Y%d" % i])
        y = np.atleast_2d(tf2.istft(Y))
        assert np.array_equal(y, g["y%d" % i])
    assert np.array_equal(tf.previous_input, g["prev_in"])
    assert np.array_equal(tf2.previous_output, g["prev_out"])


def test_perfect_reconstruction_delay():
    # analytic: sqrt-Hann / sqrt-Hann at hop = n_fft/2 reconstructs with delay n_fft - hop
    rng = np.random.default_rng(0)
    x = rng.standard_normal(256 * 20)
    tf = O.Transform(n_fft=512, hop_length=256, channel=1)
    y = tf.istft(tf.stft(x))
    assert np.max(np.abs(y[256:] - x[:-256])) < 2e-6


def test_fixed_beamformer_golden():
    g = golden("fixedbf.npz")
    geo = O.MicGeometry("circular", r=0.05, M=8, n_fft=256)
    x = g["x"].astype(np.float64)
    W_sd = O.fixed_weights(geo, 256, g["angle"], "SD")
    W_ds = O.fixed_weights(geo, 256, g["angle"], "DS")
    assert np.allclose(W_sd, g["W_sd"], rtol=0, atol=1e-12)
    assert np.allclose(W_ds, g["W_ds"], rtol=0, atol=1e-15)
    assert np.max(np.abs(O.fixed_beamform(x, W_sd, 256, 128) - g["y_sd"])) < 1e-9
    assert np.max(np.abs(O.fixed_beamform(x, W_ds, 256, 128) - g["y_ds"])) < 1e-9
    # distortionless: w^H a = 1 for both designs
    a0 = O.steering_from_doa(geo, 256, g["angle"])
    assert np.allclose(np.sum(np.conj(W_sd) * a0, axis=1), 1.0, atol=1e-9)
    assert np.allclose(np.sum(np.conj(W_ds) * a0, axis=1), 1.0, atol=1e-12)


def test_mcra_golden_bit_exact():
    g = golden("mcra.npz")
    P = g["P"]
    m = O.Mcra(nfft=256)
    for n in range(P.shape[1]):
        lam = m.estimation(P[:, n])
        assert np.array_equal(lam, g["lambda_d"][:, n]), n
        assert np.array_equal(m.p, g["p"][:, n]), n
    assert np.array_equal(m.S, g["S"]) and np.array_equal(m.Smin, g["Smin"]) and np.array_equal(m.Stmp, g["Stmp"])
    assert m.ell == int(g["ell"]) and m.frm_cnt == int(g["frm_cnt"])


def test_mcra_batch_axis_matches_single():
    g = golden("mcra.npz")
    P = g["P"][:, :40]
    mb = O.Mcra(nfft=256, batch_shape=(3,))
    ms = O.Mcra(nfft=256)
    for n in range(P.shape[1]):
        lb = mb.estimation(np.stack([P[:, n], 2 * P[:, n], P[:, n]]))
        ls = ms.estimation(P[:, n])
        assert np.array_equal(lb[0], ls) and np.array_equal(lb[2], ls)


def test_chain_mcspp_mvdr_golden():
    g = golden("chain_mcspp_mvdr.npz")
    geo = O.MicGeometry("circular", r=0.05, M=8, n_fft=512)
    taps = {}
    y = O.mvdr_mcspp_chain(g["x"].astype(np.float64), geo, g["look"], 512, 256, taps=taps)
    assert np.max(np.abs(y - g["y"])) < 1e-7 and snr_db(g["y"], y) > 120
    assert np.allclose(taps["a0"], g["a0"], rtol=0, atol=1e-12)
    assert np.allclose(taps["xi"], g["xi"], rtol=1e-6)
    assert np.allclose(taps["p"], g["p"], atol=1e-6)
    assert np.allclose(taps["est"].Phi_vv, g["Phi_vv_last"], rtol=1e-9, atol=1e-15)
    assert np.allclose(taps["est"].w, g["w_pmwf_last"], rtol=1e-6, atol=1e-12)


def test_adaptive_mvdr_golden():
    g = golden("adaptive_mvdr.npz")
    geo = O.MicGeometry("circular", r=0.032, M=4, n_fft=256)
    y = O.adaptive_mvdr(g["x"].astype(np.float64), geo, g["angle_rad"], 256, 128)
    assert np.max(np.abs(y - g["y"])) < 1e-7 and snr_db(g["y"], y) > 120


def test_synth_is_deterministic_and_bounded():
    geo = O.MicGeometry("circular", r=0.05, M=8, n_fft=512)
    a = O.synth_streams(2, geo, 4096)
    b = O.synth_streams(1, geo, 4096, first_stream=1)
    assert a.dtype == np.float32 and a.shape == (2, 8, 4096)
    assert np.array_equal(a[1], b[0])
    assert np.abs(a).max() < 1.0
