"""CPU: round-2 parity pins -- goldens dumped from the unmodified reference by tests/golden/make_golden_r2.py.

  * a18: the SRP oracle (O.srp_angle_spectrum / O.srp_map) against srp.compute_angle_spectrum (doa/srp.py:17-53)
  * a19: the host-side diagnostics of the drop-in beamformer class against beamformer.py:435-534
  * config 1: the online-MVDR oracle on a 4-mic linear array (n_fft 512 / hop 256, 10 s) and on the recording the
    reference's example/run_MVDRbeamformer.py processes (example/test_audio/rec1, first 10 s)
"""
import numpy as np

from conftest import golden, snr_db
from oracle import np_oracle as O


def test_srp_oracle_golden():
    g = golden("srp.npz")
    geo = O.MicGeometry("circular", r=0.032, M=4, n_fft=256)
    x = g["x"].astype(np.float64)
    P, p = O.srp_angle_spectrum(x, geo)
    assert P.shape == g["angle_spectrum"].shape == (360, x.shape[0] // 128)
    assert np.max(np.abs(P - g["angle_spectrum"])) <= 1e-12 * np.max(np.abs(g["angle_spectrum"]))
    assert np.array_equal(p, g["p"])                                       # MCRA (L = 65) speech presence, bit-exact
    assert np.array_equal(np.argmax(P, axis=0), np.argmax(g["angle_spectrum"], axis=0))
    Pn, _ = O.srp_angle_spectrum(x, geo, phat=False)
    assert np.max(np.abs(Pn - g["angle_spectrum_nophat"])) <= 1e-12 * np.max(np.abs(g["angle_spectrum_nophat"]))


def test_diagnostics_golden():
    """compute_array_gain / compute_wng_di / compute_beampattern are host NumPy in the drop-in class."""
    from distantspeech_b200.beamformer.MicArray import MicArray
    from distantspeech_b200.beamformer.beamformer import beamformer
    g = golden("diagnostics.npz")
    mic = MicArray(arrayType="circular", r=0.05, M=6, n_fft=64)
    bf = beamformer(mic, frame_len=64, hop=32, nfft=64)
    assert np.allclose(bf.Fvv, g["Fvv"], rtol=0, atol=1e-15)
    a0 = bf.compute_steering_vector_from_doa((40, 0))
    assert np.allclose(a0, g["a0"], rtol=0, atol=1e-14)
    W_ds, W_sd = bf.compute_weights((40, 0), "DS"), bf.compute_weights((40, 0), "SD")
    assert np.allclose(W_ds, g["W_ds"], rtol=0, atol=1e-14) and np.allclose(W_sd, g["W_sd"], rtol=1e-9, atol=1e-12)
    # the reference's array gain is [bins, bins] (numerator [K, 1] against a [K, 1, 1] quadratic form, beamformer.py:456-458)
    G = bf.compute_array_gain(g["W_sd"], g["a0"], g["Fvv"])
    assert G.shape == g["gain_sd"].shape == (33, 33)
    assert np.allclose(G, g["gain_sd"], rtol=1e-12, atol=0)
    assert np.allclose(bf.compute_array_gain(g["W_sd"], g["a0"], g["Fvv"], return_db=True), g["gain_sd_db"], rtol=1e-12, atol=1e-12)
    for tag in ("ds", "sd"):
        wng, di = bf.compute_wng_di(g["W_" + tag], look_angle=[40, 0])
        assert np.allclose(wng, g["wng_" + tag], rtol=1e-10, atol=1e-10) and np.allclose(di, g["di_" + tag], rtol=1e-10, atol=1e-10)
    wng, di = bf.compute_wng_di(g["W_sd"], look_angle=[40, 0], return_db=False)
    assert np.allclose(wng, g["wng_sd_lin"], rtol=1e-12) and np.allclose(di, g["di_sd_lin"], rtol=1e-12)
    wng, di = bf.compute_wng_di(look_angle=[40, 0])
    assert np.allclose(wng, g["wng_default"], rtol=1e-10, atol=1e-10) and np.allclose(di, g["di_default"], rtol=1e-10, atol=1e-10)
    # distortionless designs: the diagonal of the white-noise gain of delay-and-sum is M
    assert np.allclose(np.diag(bf.compute_wng_di(g["W_ds"], look_angle=[40, 0], return_db=False)[0]), 6.0, rtol=1e-9)
    bp = bf.compute_beampattern(mic, weights=g["W_sd"].T.copy())
    assert bp.shape == (360, 33) and np.allclose(bp, g["bp_sd"], rtol=0, atol=1e-9)
    bp0 = bf.compute_beampattern(mic, look_angle=np.array([40, 0]))
    assert np.allclose(bp0, g["bp_default"], rtol=0, atol=1e-9)


def test_adaptive_mvdr_linear_golden():
    """Config 1 as BASELINE.json words it: 4-mic linear array, 16 kHz, n_fft 512 / hop 256, one 10 s utterance."""
    g = golden("adaptive_mvdr_linear.npz")
    geo = O.MicGeometry("linear", r=0.032, M=4, n_fft=512)
    with np.errstate(all="ignore"):
        y = O.adaptive_mvdr(g["x"].astype(np.float64), geo, g["angle_rad"], 512, 256)
    assert y.shape == g["y"].shape == (160000,)
    assert np.max(np.abs(y - g["y"])) < 1e-6 and snr_db(g["y"], y) > 110      # golden stored as float32


def test_adaptive_mvdr_rec1_golden():
    """Config 1 as the reference's example script runs it: the shipped recording, circular r = 0.032, n_fft 256 / hop 128,
    look direction 197 deg; int16 PCM scaled like load_audio (utils.py:182-187)."""
    g = golden("adaptive_mvdr_rec1.npz")
    x = (g["pcm"].astype(np.float32) / np.float32(32767.0)).astype(np.float64)
    geo = O.MicGeometry("circular", r=0.032, M=4, n_fft=256)
    with np.errstate(all="ignore"):
        y = O.adaptive_mvdr(x, geo, g["angle_rad"], 256, 128)
    assert y.shape == g["y"].shape == (160000,)
    assert np.max(np.abs(y - g["y"])) < 1e-6 and snr_db(g["y"], y) > 110


def test_mcspp_more_than_4_mics_golden():
    """8f.3: the McSpp oracle with 5 / 6 microphones is bit-identical to the reference run with McCDR(nfft, channels=M)
    handed in (oracle/ref_harness.make_mcspp), and so is estimation(repeat=True) at 4 microphones."""
    g = golden("mcspp_cdr_m68.npz")
    for tag, M, rep in (("m6", 6, False), ("m5", 5, False), ("m4r", 4, True)):
        D = O.Transform(channel=M, n_fft=256, hop_length=128).stft(g[tag + "_x"].astype(np.float64))
        est = O.McSpp(nfft=256, channels=M)
        with np.errstate(all="ignore"):
            for n in range(D.shape[1]):
                p = est.estimation(D[:, n, :], repeat=rep)
                assert np.array_equal(p, g[tag + "_p"][:, n]) and np.array_equal(est.xi, g[tag + "_xi"][:, n]), (tag, n)
                assert np.array_equal(est.q, g[tag + "_q"][:, n]) and np.array_equal(est.gamma, g[tag + "_gamma"][:, n]), (tag, n)
        assert np.array_equal(est.w, g[tag + "_w_last"]) and np.array_equal(est.Phi_vv_inv, g[tag + "_Phi_vv_inv_last"])
        assert np.array_equal(est.mccdr.Pxii, g[tag + "_Pxii_last"])


def test_wpe_golden():
    """8f.4: the WPE oracle reproduces the filter state of the reference's own update body (awpe.py:152-187) bit for bit."""
    g = golden("wpe.npz")
    C, Lf, nb, hop, D = (int(v) for v in g["params"])
    o = O.WpeOracle(channels=C, filter_len=Lf, num_bands=nb, delay=D, hop_length=hop)
    tx = O.Transform(n_fft=nb, hop_length=hop, channel=C)
    x = g["x"].astype(np.float64)
    for n in range(60):
        o.update_spec(tx.stft(x[n * hop:(n + 1) * hop])[:, 0, :])
        if n + 1 in (30, 60):
            assert np.array_equal(o.W, g["W%d" % (n + 1)]) and np.array_equal(o.P, g["P%d" % (n + 1)])
            assert np.array_equal(o.var, g["var%d" % (n + 1)])
    # the error signal is a dereverberated version of the input: energy drops once the filter has adapted
    y = O.WpeOracle(channels=C, filter_len=Lf, num_bands=nb, delay=D, hop_length=hop).process(x)
    assert y.shape == x.shape and np.sum(y[hop * 30:] ** 2) < np.sum(x[hop * 29:-hop] ** 2)


def test_realtime_chunk_api_host_logic():
    """8f.2: the chunk API mirrors realtime_processing.process / the capture loop's arithmetic (realtime_processing.py:78-84,
    :113-131): / 32768 scaling, channels 1..4 in, channel 5 out, int16 bytes back."""
    from distantspeech_b200.realtime.realtime_processing import realtime_processing
    rng = np.random.default_rng(2)
    pcm = rng.integers(-20000, 20000, size=(1024, 6), dtype=np.int16)

    class Mean(object):
        def process(self, data):
            return {"data": data.mean(axis=1)}
    rt = realtime_processing(EnhancementMehtod=Mean(), chunk=1024, channels=6)
    out = np.frombuffer(rt.process_pcm(pcm.tobytes()), dtype='<i2')
    f = pcm.astype(np.float32) / 32768.0
    assert np.array_equal(out, (f[:, 1:5].mean(axis=1) * 32768).astype('<i2'))
    passthrough = realtime_processing(EnhancementMehtod=None, chunk=1024, channels=6)
    assert np.array_equal(np.frombuffer(passthrough.process_pcm(pcm.tobytes()), dtype='<i2'), pcm[:, 2])      # data[:, 1] of channels 1..4
    rec = realtime_processing(EnhancementMehtod=None, chunk=1024, channels=6, save_rec_to_file=True)
    allch = np.frombuffer(rec.process_pcm(pcm.tobytes()), dtype='<i2').reshape(1024, 6)
    assert np.array_equal(allch[:, :5], pcm[:, :5]) and np.array_equal(allch[:, 5], pcm[:, 2])
