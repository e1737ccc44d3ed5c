"""GPU parity tests: CUDA path (through the C ABI) vs the numpy oracle and the
golden vectors dumped from the reference.  Tolerances: spectra are compared
relative to the frame peak (fp32 FFT), waveforms by the north_star contract
(max-abs <= 1e-4 and SNR >= 60 dB)."""
import numpy as np
import pytest

from conftest import golden, snr_db, assert_wave_parity
from oracle import np_oracle as O

pytestmark = pytest.mark.gpu


# ---------------------------------------------------------------- a1 / a2
@pytest.mark.parametrize("n_fft,hop,center", [(512, 128, True), (256, 128, False), (1024, 512, True), (128, 64, True),
                                              (2048, 512, False), (512, 200, True)])
@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_stft_vs_oracle(cuda, n_fft, hop, center, precision):
    from distantspeech_b200.transform.transform import stft
    rng = np.random.default_rng(n_fft + hop)
    x = (rng.standard_normal(9000) * 0.2).astype(np.float32)
    win = O.sqrt_hann(n_fft)
    ref = O.stft(x.astype(np.float64), n_fft=n_fft, hop_length=hop, window=win, center=center)
    out = stft(x, n_fft=n_fft, hop_length=hop, window=win, center=center, precision=precision)
    assert out.shape == ref.shape and out.dtype == np.complex64
    scale = np.max(np.abs(ref))
    tol = 3e-6 if precision == "fp32" else 2.5e-7          # fp64 path: complex64 rounding only
    assert np.max(np.abs(out - ref)) <= tol * scale


@pytest.mark.parametrize("hop,center,n", [(77, True, 9000), (256, False, 256 * 40 + 512), (256, True, 256 * 41), (129, False, 7001)])
def test_stft_512_half_warp_paths(cuda, hop, center, n):
    # n_fft = 512 / fp32 runs the half-warp-per-frame kernel (stft_sq_kernel): odd frame counts (one idle half-warp),
    # rows whose frames are not 8-byte aligned (scalar loader), reflect padding, several channels, complex128 output
    from distantspeech_b200.transform.transform import stft
    rng = np.random.default_rng(hop + n)
    x = (rng.standard_normal((3, n)) * 0.2).astype(np.float32)
    win = O.sqrt_hann(512)
    for c in range(3):
        ref = O.stft(x[c].astype(np.float64), n_fft=512, hop_length=hop, window=win, center=center)
        out = stft(x[c], n_fft=512, hop_length=hop, window=win, center=center, precision="fp32")
        assert out.shape == ref.shape
        assert np.max(np.abs(out - ref)) <= 3e-6 * np.max(np.abs(ref))


def test_stft_golden_and_errors(cuda):
    from distantspeech_b200.transform.transform import stft, istft
    g = golden("stft_istft.npz")
    D = stft(g["x"], n_fft=512, hop_length=128, window=O.sqrt_hann(512), center=True, precision="fp64")
    ref = g["D_512_128_center"]
    assert np.max(np.abs(D - ref)) <= 2.5e-7 * np.max(np.abs(ref))
    y = istft(ref, hop_length=128, window=O.sqrt_hann(512), center=True, length=6000)
    assert y.dtype == np.float32 and np.max(np.abs(y - g["y_512_128_center"])) < 2e-6
    y2 = istft(g["D_256_128_plain"], hop_length=128, window=O.sqrt_hann(256), center=False)
    assert np.max(np.abs(y2 - g["y_256_128_plain"])) < 2e-6
    with pytest.raises(AttributeError):
        stft(g["x"], n_fft=512)                            # default window="hann" is dead in the reference
    with pytest.raises(ValueError):
        stft(g["x"], n_fft=512, window=None)


@pytest.mark.parametrize("n_fft,hop", [(512, 256), (512, 128), (256, 128), (1024, 256)])
def test_istft_gain_quirk(cuda, n_fft, hop):
    # module-level istft applies no window-sum normalisation: gain 1 at hop=n/2, 2 at hop=n/4 (quirk 2)
    from distantspeech_b200.transform.transform import stft, istft
    rng = np.random.default_rng(3)
    x = (rng.standard_normal(8192) * 0.2).astype(np.float32)
    win = O.sqrt_hann(n_fft)
    y = istft(stft(x, n_fft=n_fft, hop_length=hop, window=win), hop_length=hop, window=win, length=8192)
    gain = (n_fft // hop) / 2.0
    assert np.max(np.abs(y[n_fft:-n_fft] - gain * x[n_fft:-n_fft])) < 5e-6
    ref = O.istft(O.stft(x.astype(np.float64), n_fft=n_fft, hop_length=hop, window=win), hop_length=hop, window=win, length=8192)
    assert np.max(np.abs(y - ref)) < 5e-6


# ---------------------------------------------------------------- a3
def test_transform_streaming_golden(cuda):
    from distantspeech_b200.transform.transform import Transform
    g = golden("transform_stream.npz")
    x = g["x"]
    tf = Transform(n_fft=512, hop_length=256, channel=3)
    tf2 = Transform(n_fft=512, hop_length=256, channel=3)
    cuts = [(0, 256 * 5), (256 * 5, 256 * 6), (256 * 6, 256 * 12)]
    for i, (a, b) in enumerate(cuts):
        Y = tf.stft(x[a:b])
        ref = g["Y%d" % i]
        assert Y.shape == ref.shape and Y.dtype == np.complex128
        assert np.max(np.abs(Y - ref)) <= 3e-6 * np.max(np.abs(ref))
        y = np.atleast_2d(tf2.istft(ref))
        assert y.shape == g["y%d" % i].shape
        assert np.max(np.abs(y - g["y%d" % i])) < 2e-6
    assert np.array_equal(tf.previous_input, g["prev_in"])
    assert np.max(np.abs(tf2.previous_output - g["prev_out"])) < 2e-6


def test_transform_rank_semantics_and_reconstruction(cuda):
    from distantspeech_b200.transform.transform import Transform
    rng = np.random.default_rng(5)
    x = (rng.standard_normal((256 * 16, 2)) * 0.2).astype(np.float32)
    tf = Transform(n_fft=512, hop_length=256, channel=2)
    Y = tf.stft(x)
    assert Y.shape == (257, 16, 2)
    y = tf.istft(Y)
    assert y.shape == (256 * 16, 2)
    assert np.max(np.abs(y[256:] - x[:-256])) < 3e-6          # delay = n_fft - hop
    # 2-D input to istft means ONE frame x channels (quirk 5), 1-D means one frame / one channel
    tf3 = Transform(n_fft=512, hop_length=256, channel=2)
    assert tf3.istft(Y[:, 0, :]).shape == (256, 2)
    assert tf3.istft(Y[:, 1, 0]).shape == (256,)
    # hop-by-hop streaming equals batch
    tfs = Transform(n_fft=512, hop_length=256, channel=2)
    tfo = Transform(n_fft=512, hop_length=256, channel=2)
    ys = np.concatenate([tfo.istft(tfs.stft(x[i:i + 256])) for i in range(0, x.shape[0], 256)])
    assert np.max(np.abs(ys - y)) < 1e-6


# ---------------------------------------------------------------- a9
@pytest.mark.parametrize("wt", ["SD", "DS"])
def test_fixed_beamformer_golden(cuda, wt):
    from distantspeech_b200.beamformer.MicArray import MicArray
    from distantspeech_b200.beamformer.fixedbeamformer import FixedBeamformer
    g = golden("fixedbf.npz")
    mic = MicArray(arrayType="circular", r=0.05, M=8, n_fft=256)
    fb = FixedBeamformer(mic, 256, 128, 256)
    y = fb.process(g["x"], tuple(g["angle"]), weightType=None if wt == "SD" else "DS")
    assert np.allclose(fb.W, g["W_sd" if wt == "SD" else "W_ds"], rtol=0, atol=1e-12)
    assert_wave_parity(g["y_sd" if wt == "SD" else "y_ds"], y, "fixed %s" % wt)


def test_fixed_beamformer_batch_stream_multibeam(cuda):
    from distantspeech_b200.beamformer.MicArray import MicArray
    from distantspeech_b200.beamformer.fixedbeamformer import FixedBeamformer
    geo = O.MicGeometry("circular", r=0.05, M=8, n_fft=512)
    xs = O.synth_streams(5, geo, 256 * 40, seed0=11)                       # [S, M, N]
    x_nm = np.ascontiguousarray(xs.transpose(0, 2, 1))                     # [S, N, M]
    mic = MicArray(arrayType="circular", r=0.05, M=8, n_fft=512)
    fb = FixedBeamformer(mic, 512, 256, 512)
    y = fb.process(x_nm, (30, 0))
    W = O.fixed_weights(geo, 512, (30, 0), "SD")
    for s in range(5):
        assert_wave_parity(O.fixed_beamform(x_nm[s].astype(np.float64), W, 512, 256), y[s], "stream %d" % s)
    # two chunks == one call (streaming state), segment split == no split (S small -> segmented launch)
    fb2 = FixedBeamformer(mic, 512, 256, 512)
    ya = fb2.process(x_nm[:, :256 * 15], (30, 0))
    yb = fb2.process(x_nm[:, 256 * 15:], (30, 0))
    assert np.max(np.abs(np.concatenate([ya, yb], axis=1) - y)) < 1e-6
    # multibeam
    fb3 = FixedBeamformer(mic, 512, 256, 512)
    ym = fb3.process_multibeam(x_nm, [(30, 0), (200, 0), (90, 10)], weightType="DS")
    for b, ang in enumerate([(30, 0), (200, 0), (90, 10)]):
        Wb = O.fixed_weights(geo, 512, ang, "DS")
        assert_wave_parity(O.fixed_beamform(x_nm[1].astype(np.float64), Wb, 512, 256), ym[1, b], "beam %d" % b)


def test_fixed_beamformer_quarter_hop(cuda):
    from distantspeech_b200.beamformer.MicArray import MicArray
    from distantspeech_b200.beamformer.fixedbeamformer import FixedBeamformer
    geo = O.MicGeometry("linear", r=0.04, M=4, n_fft=256)
    x = np.ascontiguousarray(O.synth_streams(1, geo, 64 * 100, seed0=5)[0].T)
    mic = MicArray(arrayType="linear", r=0.04, M=4, n_fft=256)
    fb = FixedBeamformer(mic, 256, 64, 256)
    y = fb.process(x, (60, 0), weightType="DS")
    ref = O.fixed_beamform(x.astype(np.float64), O.fixed_weights(geo, 256, (60, 0), "DS"), 256, 64)
    assert_wave_parity(ref, y, "hop n/4")


# ---------------------------------------------------------------- a12
def test_mcra_golden_bit_exact(cuda):
    from distantspeech_b200.noise_estimation.mcra import NoiseEstimationMCRA
    g = golden("mcra.npz")
    P = g["P"]
    m = NoiseEstimationMCRA(nfft=256)
    for n in range(20):                                       # per-frame API like the reference loop
        lam = m.estimation(P[:, n])
        assert np.array_equal(lam, g["lambda_d"][:, n]), n
        assert np.array_equal(m.p, g["p"][:, n]), n
    lam, p = m.estimation_frames(P[:, 20:].T.copy(), return_p=True)   # rest in one launch
    assert np.array_equal(lam.T, g["lambda_d"][:, 20:])
    assert np.array_equal(p.T, g["p"][:, 20:])
    assert np.array_equal(m.S, g["S"]) and np.array_equal(m.Smin, g["Smin"]) and np.array_equal(m.Stmp, g["Stmp"])
    assert m.ell == int(g["ell"]) and m.frm_cnt == int(g["frm_cnt"])


# ---------------------------------------------------------------- a13 + a6 + config 4
def test_mcspp_estimator_golden(cuda):
    from distantspeech_b200.noise_estimation.mcspp_base import McSppBase
    from distantspeech_b200.beamformer.beamformer import compute_mvdr_weight
    g = golden("chain_mcspp_mvdr.npz")
    D = O.Transform(n_fft=512, hop_length=256, channel=8).stft(g["x"].astype(np.float64))     # oracle spectrum
    est = McSppBase(nfft=512, channels=8)
    nT = D.shape[1]
    worst = {}
    for n in range(12):                                       # per-frame API
        p = est.estimation(D[:, n, :])
        est.compute_omlsa_weight(est.xi, est.p)
        for k, v in (("p", p), ("xi", est.xi), ("gamma", est.gamma), ("q", est.q), ("G", est.G)):
            ref = g[k][:, n].astype(np.float64)
            rel = np.max(np.abs(v - ref) / (np.abs(ref) + 1e-12))
            worst[k] = max(worst.get(k, 0), rel)
    assert worst["q"] < 1e-6 and worst["p"] < 1e-5 and worst["G"] < 1e-5, worst
    assert worst["xi"] < 1e-6 and worst["gamma"] < 1e-6, worst
    res = est.estimation_frames(D[:, 12:, :], a0=g["a0"])     # remaining frames in one launch
    assert np.allclose(res["xi"], g["xi"][:, 12:], rtol=1e-5)
    assert np.allclose(res["p"], g["p"][:, 12:], atol=1e-5)
    assert np.allclose(est.Phi_vv, g["Phi_vv_last"], rtol=1e-7, atol=1e-12)
    assert np.allclose(est.Phi_yy, g["Phi_yy_last"], rtol=1e-7, atol=1e-12)
    assert np.allclose(est.w, g["w_pmwf_last"], rtol=1e-5, atol=1e-9)
    assert np.allclose(est.Phi_vv_inv, g["Phi_vv_inv_last"], rtol=1e-6, atol=1e-9)
    Yref = g["Yspec"][:, 12:]
    assert np.max(np.abs(res["Y"] - Yref)) <= 1e-5 * np.max(np.abs(Yref))
    w = compute_mvdr_weight(g["a0"], est.Phi_vv_inv)
    wr = O.mvdr_weight(g["a0"], g["Phi_vv_inv_last"])
    assert np.allclose(w, wr, rtol=1e-5, atol=1e-9)
    assert np.allclose(np.sum(np.conj(w) * g["a0"], axis=1), 1.0, atol=1e-9)      # distortionless


# ---------------------------------------------------------------- a14
def test_mcspp_cdr_golden(cuda):
    """McSpp with the McCDR prior against the reference's own per-frame outputs (240 frames, 4 mics)."""
    from distantspeech_b200.noise_estimation.mcspp import McSpp
    from distantspeech_b200.noise_estimation.mccdr import McCDR
    from distantspeech_b200.transform.transform import Transform
    g = golden("mcspp_cdr.npz")
    D = O.Transform(channel=4, n_fft=512, hop_length=256).stft(g["x"].astype(np.float64))   # the reference's spectrum
    est = McSpp(nfft=512, channels=4)
    res = est.estimation_frames(D, want_Y=True)
    for k in ("p", "xi", "gamma", "q"):
        assert np.allclose(res[k], g[k], rtol=1e-8, atol=1e-12), k
    assert np.array_equal(est.mccdr.mcra.p, g["mcra_p_last"])                   # MCRA decisions of the prior identical
    scale = lambda a: np.max(np.abs(a))
    for nm, a in (("w_last", est.w), ("Phi_yy_last", est.Phi_yy), ("Phi_vv_last", est.Phi_vv),
                  ("Phi_vv_inv_last", est.Phi_vv_inv), ("Phi_xx_last", est.Phi_xx)):
        assert np.max(np.abs(a - g[nm])) <= 1e-10 * scale(g[nm]), nm
    assert np.allclose(res["cdr"][:, 10:], 1 - g["q"][:, 10:], rtol=0, atol=1e-12)   # q = 1 - McCDR.estimation (:117-118)
    Yref = np.einsum("ktm,ktm->kt", np.conj(res["w"]), D)
    assert np.max(np.abs(res["Y"] - Yref)) <= 2e-7 * scale(Yref)                # complex64 output
    # one frame per call == many frames per call; McCDR alone == the prior inside McSpp
    e2, cdr = McSpp(nfft=512, channels=4), McCDR(nfft=512)
    for n in range(25):
        p = e2.estimation(D[:, n, :])
        G = cdr.estimation(D[:, n, :])
    assert np.array_equal(p, res["p"][:, 24]) and np.array_equal(e2.w, res["w"][:, 24]) and np.array_equal(e2.q, res["q"][:, 24])
    assert np.array_equal(G, res["cdr"][:, 24])
    # from the waveform through the fp32 device STFT: same decisions, probabilities within 1e-4
    e3 = McSpp(nfft=512, channels=4)
    r3 = e3.estimation_frames(Transform(n_fft=512, hop_length=256, channel=4).stft(g["x"]))
    assert np.max(np.abs(r3["p"] - g["p"])) < 1e-4 and np.max(np.abs(r3["q"] - g["q"])) < 1e-4
    # streams are independent: batch of two == two singles
    rb = McSpp(nfft=512, channels=4).estimation_frames(np.stack([D[:, :60], D[:, 60:120]]))
    assert np.array_equal(rb["p"][0], res["p"][:, :60]) and rb["p"].shape == (2, 257, 60)
    with pytest.raises(ValueError):
        McSpp(nfft=512, channels=3)                                            # the CDR pair (1, 2) is undefined below 4 channels


@pytest.mark.parametrize("prec", ["fp32", "fp64"])
@pytest.mark.parametrize("full", [False, True])
def test_chain_golden(cuda, prec, full):
    from distantspeech_b200.beamformer.MicArray import MicArray
    from distantspeech_b200.pipelines import MvdrMcsppChain
    g = golden("chain_mcspp_mvdr.npz")
    mic = MicArray(arrayType="circular", r=0.05, M=8, n_fft=512)
    ch = MvdrMcsppChain(mic, look_angle=tuple(g["look"]), n_fft=512, hop=256, fft_precision=prec, full_state=full)
    y = ch.process(g["x"])
    err, s = assert_wave_parity(g["y"], y, "chain %s full=%s" % (prec, full))
    print("chain %s full=%s: max-abs %.2e SNR %.1f dB" % (prec, full, err, s))


def test_chain_batch_and_streaming(cuda):
    from distantspeech_b200.beamformer.MicArray import MicArray
    from distantspeech_b200.pipelines import MvdrMcsppChain
    geo = O.MicGeometry("circular", r=0.05, M=8, n_fft=512)
    xs = O.synth_streams(4, geo, 256 * 125, seed0=0x5EED)                  # 2 s
    mic = MicArray(arrayType="circular", r=0.05, M=8, n_fft=512)
    x_nm = np.ascontiguousarray(xs.transpose(0, 2, 1))
    ch = MvdrMcsppChain(mic, look_angle=(30, 0))
    y = ch.process(x_nm)
    worst = 1e9
    for s in range(4):
        ref = O.mvdr_mcspp_chain(x_nm[s].astype(np.float64), geo, (30, 0), 512, 256)
        err, snr = assert_wave_parity(ref, y[s], "chain stream %d" % s)
        worst = min(worst, snr)
    print("chain batch worst SNR %.1f dB" % worst)
    ch2 = MvdrMcsppChain(mic, look_angle=(30, 0))
    ya = ch2.process(x_nm[:, :256 * 50])
    yb = ch2.process(x_nm[:, 256 * 50:])
    assert np.max(np.abs(np.concatenate([ya, yb], axis=1) - y)) < 1e-6      # chunked == whole
    # ... down to single-hop chunks: the first frame of a chunk (history path) and interior frames
    # (register-fed path) round identically, so the hard MCRA decisions cannot flip between the two
    ch3 = MvdrMcsppChain(mic, look_angle=(30, 0))
    cuts = [0, 256, 256 * 3, 256 * 4, 256 * 40, 256 * 125]
    yc = np.concatenate([ch3.process(x_nm[:, a:b]) for a, b in zip(cuts[:-1], cuts[1:])], axis=1)
    assert np.array_equal(yc, y)
    # host-buffer pipeline == device path
    import torch
    yh = ch.process_host(torch.from_numpy(xs).pin_memory(), slice_frames=7)
    assert np.max(np.abs(yh.numpy() - y)) < 1e-6


@pytest.mark.parametrize("M", [2, 4, 6])
def test_chain_other_mic_counts(cuda, M):
    from distantspeech_b200.beamformer.MicArray import MicArray
    from distantspeech_b200.pipelines import MvdrMcsppChain
    geo = O.MicGeometry("circular" if M != 6 else "linear", r=0.04, M=M, n_fft=256)
    x = np.ascontiguousarray(O.synth_streams(1, geo, 128 * 120, seed0=M)[0].T)
    mic = MicArray(arrayType="circular" if M != 6 else "linear", r=0.04, M=M, n_fft=256)
    ch = MvdrMcsppChain(mic, look_angle=(30, 0), n_fft=256, hop=128)
    ref = O.mvdr_mcspp_chain(x.astype(np.float64), geo, (30, 0), 256, 128)
    assert_wave_parity(ref, ch.process(x), "chain M=%d" % M)


# ---------------------------------------------------------------- a11 (config 1)
def test_adaptive_mvdr_golden(cuda):
    from distantspeech_b200.beamformer.MicArray import MicArray
    from distantspeech_b200.beamformer.adaptivebeamformer import adaptivebeamfomer
    g = golden("adaptive_mvdr.npz")
    mic = MicArray(arrayType="circular", r=0.032, M=4, n_fft=256)
    ab = adaptivebeamfomer(mic, 256, 128, 256)
    out = ab.process(g["x"], g["angle_rad"], method=2)
    assert set(out.keys()) == {"data", "WNG", "DI", "beampattern"}
    err, s = assert_wave_parity(g["y"], out["data"], "adaptive MVDR")
    print("adaptive MVDR: max-abs %.2e SNR %.1f dB" % (err, s))
    assert np.array_equal(ab.mcra.p, g["p_last"])                      # MCRA decisions identical
    assert np.allclose(ab.Rvv, g["Rvv_last"], rtol=1e-6, atol=1e-12)
    assert np.allclose(ab.H, g["H_last"], rtol=1e-4, atol=1e-7)
    with pytest.raises(AttributeError):
        ab.process(g["x"], g["angle_rad"], method=2, retWNG=True)


def test_adaptive_mvdr_512_streaming_batch(cuda):
    from distantspeech_b200.beamformer.MicArray import MicArray
    from distantspeech_b200.beamformer.adaptivebeamformer import adaptivebeamfomer
    geo = O.MicGeometry("circular", r=0.032, M=4, n_fft=512)
    xs = O.synth_streams(3, geo, 256 * 100, seed0=21)                  # [S, M, N]
    mic = MicArray(arrayType="circular", r=0.032, M=4, n_fft=512)
    ang = np.array([30, 0]) / 180 * np.pi
    ab = adaptivebeamfomer(mic, 512, 256, 512)
    y = ab.process(xs, ang, method=2)["data"]
    for s in range(3):
        ref = O.adaptive_mvdr(xs[s].astype(np.float64), geo, ang, 512, 256)
        assert_wave_parity(ref, y[s], "adaptive MVDR stream %d" % s)
    ab2 = adaptivebeamfomer(mic, 512, 256, 512)
    ya = ab2.process(xs[:, :, :256 * 40], ang, method=2)["data"]
    yb = ab2.process(xs[:, :, 256 * 40:], ang, method=2)["data"]
    assert np.max(np.abs(np.concatenate([ya, yb], axis=1) - y)) < 1e-6


def test_adaptive_other_methods(cuda):
    from distantspeech_b200.beamformer.MicArray import MicArray
    from distantspeech_b200.beamformer.adaptivebeamformer import adaptivebeamfomer
    geo = O.MicGeometry("circular", r=0.032, M=4, n_fft=256)
    x = O.synth_streams(1, geo, 128 * 60, seed0=8)[0]
    mic = MicArray(arrayType="circular", r=0.032, M=4, n_fft=256)
    ang = np.array([50, 0]) / 180 * np.pi
    tao = -1 * geo.r * np.cos(ang[1]) * np.cos(ang[0] - geo.gamma) / geo.c
    a = np.exp(-1j * (2 * np.pi * np.arange(129) * 16000 / 256)[:, None] * tao[None, :])        # [K, M]
    y_ds = adaptivebeamfomer(mic, 256, 128, 256).process(x, ang, method=1)["data"]
    assert_wave_parity(O.fixed_beamform(x.T.astype(np.float64), a / 4, 256, 128), y_ds, "method DS")
    W0 = np.zeros_like(a)
    W0[:, 0] = a[:, 0]
    y_src = adaptivebeamfomer(mic, 256, 128, 256).process(x, ang, method=0)["data"]
    assert_wave_parity(O.fixed_beamform(x.T.astype(np.float64), W0, 256, 128), y_src, "method src")
    # TFGSC (beamformer.py:327-333): w = ((Rvv_inv Ryy) - I) u / (trace(Rvv_inv Ryy) - M), checked on the final state
    ab = adaptivebeamfomer(mic, 256, 128, 256)
    ab.process(x, ang, method=3)
    temp = ab.Rvv_inv @ ab.Ryy
    u = np.zeros((4, 1)); u[0] = 1
    w = ((temp - np.eye(4)) @ u)[:, :, 0] / (np.trace(temp, axis1=-2, axis2=-1) - 4)[:, None]
    assert np.allclose(ab.H.T, w, rtol=1e-6, atol=1e-9)


# ---------------------------------------------------------------- a10 / a17 (config 3)
@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_fdgsc_golden(cuda, precision):
    from distantspeech_b200.beamformer.MicArray import MicArray
    from distantspeech_b200.beamformer.FDGSC import FDGSC
    g = golden("fdgsc.npz")
    mic = MicArray(arrayType="linear", r=0.05, M=6, n_fft=256)
    fd = FDGSC(mic, frameLen=256, angle=[int(g["angle_deg"][0]), int(g["angle_deg"][1])], precision=precision)
    assert np.allclose(fd.time_alignment.delay_filter, g["delay_filter"], rtol=0, atol=1e-15)
    x = g["x"].copy()
    res = fd.process(x, postfilter=False, dc_notch=True)
    assert len(res) == 9
    # in-place DC notch (quirk 11): the time-parallel kernel reproduces the sequential filter to the last bit of float32
    # except where a double-precision last-bit difference of its carried state crosses a rounding boundary
    dn = np.abs(x - g["x_notched"])
    assert np.max(dn) <= 2 * np.spacing(np.abs(g["x_notched"]).max()) and np.mean(dn > 0) < 1e-4
    err, s = assert_wave_parity(g["y"], res[0], "FDGSC %s" % precision)
    print("FDGSC %s: max-abs %.2e SNR %.1f dB" % (precision, err, s))
    assert np.max(np.abs(res[2] - g["fix_output"])) < 2e-6
    assert np.max(np.abs(res[4] - g["bm_output"])) < 1e-4
    assert np.mean(np.abs(res[1] - g["p"]) > 1e-6) < 0.01                     # MCRA decisions (fp32 STFT input)
    rel = np.linalg.norm(fd.aic_filter.W - g["W_aic_last"]) / np.linalg.norm(g["W_aic_last"])
    assert rel < (1e-3 if precision == "fp32" else 1e-5), rel


def test_fdgsc_postfilter_golden(cuda):
    """postfilter=True reproduces the reference as written (hybrid analysis frames, buffer-wide reference
    spectra, sqrt(G) gain), over two consecutive calls so the carried Transform / OMLSA state is covered."""
    from distantspeech_b200.beamformer.MicArray import MicArray
    from distantspeech_b200.beamformer.FDGSC import FDGSC
    g, gp = golden("fdgsc.npz"), golden("fdgsc_postfilter.npz")
    n1 = int(gp["n_first"])
    mic = MicArray(arrayType="linear", r=0.05, M=6, n_fft=256)
    fd = FDGSC(mic, frameLen=256, angle=[int(g["angle_deg"][0]), int(g["angle_deg"][1])])
    ya = fd.process(g["x"][:n1].copy(), postfilter=True, dc_notch=True)[0]
    yb = fd.process(g["x"][n1:].copy(), postfilter=True, dc_notch=True)[0]
    err, s = assert_wave_parity(gp["y"], np.concatenate([ya, yb]), "FDGSC postfilter")
    print("FDGSC postfilter: max-abs %.2e SNR %.1f dB" % (err, s))
    # batch of two streams == singles; the postfiltered output differs from the plain one
    xs = np.stack([g["x"][:n1], g["x"][:n1][::-1].copy()])
    yb2 = FDGSC(mic, frameLen=256, angle=[60, 0]).process(xs.copy(), postfilter=True)[0]
    assert np.max(np.abs(yb2[0] - ya)) < 1e-6
    assert np.max(np.abs(ya - g["y"][:n1])) > 1e-3


def test_fdgsc_streaming_batch_and_time_alignment(cuda):
    from distantspeech_b200.beamformer.MicArray import MicArray
    from distantspeech_b200.beamformer.FDGSC import FDGSC, TimeAlignment
    geo = O.MicGeometry("linear", r=0.05, M=4, n_fft=256)
    xs = O.synth_streams(3, geo, 256 * 50, look_deg=(75.0, 0.0), interf_deg=(150.0, 0.0), seed0=44)
    x_nm = np.ascontiguousarray(xs.transpose(0, 2, 1))
    mic = MicArray(arrayType="linear", r=0.05, M=4, n_fft=256)
    ang = np.array([75, 0]) / 180 * np.pi
    fd = FDGSC(mic, frameLen=256, angle=[75, 0])
    y = fd.process(x_nm.copy())[0]
    for s in range(3):
        ref = O.FdgscOracle(geo, 256, ang).process(x_nm[s].astype(np.float64))[0]
        assert_wave_parity(ref, y[s], "FDGSC stream %d" % s)
    fd2 = FDGSC(mic, frameLen=256, angle=[75, 0])
    ya = fd2.process(x_nm[:, :256 * 20].copy())[0]
    yb = fd2.process(x_nm[:, 256 * 20:].copy())[0]
    assert np.max(np.abs(np.concatenate([ya, yb], axis=1) - y)) < 1e-6
    # the whole 9-tuple of the single-stream call, over two chunks (ragged tail: 100 samples beyond a block)
    orc = O.FdgscOracle(geo, 256, ang)
    fd3 = FDGSC(mic, frameLen=256, angle=[75, 0])
    for lo, hi in ((0, 256 * 12 + 100), (256 * 12 + 100, 256 * 30 + 100)):
        res = fd3.process(x_nm[0, lo:hi].copy())
        ref = orc.process(x_nm[0, lo:hi].astype(np.float64))
        assert len(res) == 9 and res[3].shape == ref[0].shape and res[5].shape == (hi - lo, 4)
        assert np.max(np.abs(res[3] - orc.diag["fix_output_delayed"])) < 2e-6
        assert np.max(np.abs(res[5] - orc.diag["aligned_output"])) < 1e-6
        assert np.max(np.abs(res[6] - orc.diag["aligned_output_delayed"])) < 1e-6
    assert fd.process(x_nm.copy())[5] is None and fd.process(x_nm.copy(), diagnostics=True)[5].shape == (3, 256 * 50, 4)
    # TimeAlignment.process == oracle FIR, streaming in two blocks
    ta = TimeAlignment(mic, angle=[75, 0])
    h = O.alignment_filters(geo, ang)
    assert np.allclose(ta.delay_filter, h, rtol=0, atol=1e-15)
    xin = x_nm[0].astype(np.float64)
    out = np.vstack([ta.process(xin[:1000]), ta.process(xin[1000:3000])])
    ref = np.stack([np.convolve(xin[:3000, m], h[:, m])[:3000] for m in range(4)], axis=1)
    assert np.max(np.abs(out - ref)) < 1e-12


def test_delay_samples_reference_unittest(cuda):
    # the reference's own unit test (tests/unittests/test_delay.py:5-23) against our DelaySamples
    from distantspeech_b200.beamformer.utils import DelaySamples
    rng = np.random.default_rng(0)
    for ch in (1, 2):
        for data_len in (1, 10, 100):
            for delay in (0, 1, 5, 50, 150):
                obj = DelaySamples(data_len, delay, channel=ch)
                x = rng.random((1000, ch))
                y = np.zeros((1000, ch))
                for n in range(1000 // data_len):
                    y[n * data_len:(n + 1) * data_len] = obj.delay(x[n * data_len:(n + 1) * data_len])
                if delay == 0:
                    assert np.sum(np.abs(y - x)) < 1e-5
                else:
                    assert np.sum(np.abs(y[delay:] - x[:-delay])) < 1e-5


# ---------------------------------------------------------------- a15 / a16
def test_omlsa_multi_golden(cuda):
    from distantspeech_b200.noise_estimation.omlsa_multi import NsOmlsaMulti
    g = golden("omlsa_multi.npz")
    Y, U = g["Y"].astype(np.float64), g["U"].astype(np.float64)
    om = NsOmlsaMulti(nfft=512, cal_weights=True, M=6)
    assert om.estimation(Y[:, 0], U[:, 0, :]) is None
    for n in range(1, 6):                                              # per-frame API
        lam = om.estimation(Y[:, n], U[:, n, :])
        assert np.allclose(lam, g["lambda_d"][n], rtol=1e-12, atol=1e-300)
        assert np.allclose(om.G, g["G"][n], rtol=1e-9) and np.allclose(om.p, g["p"][n], rtol=1e-9, atol=1e-15)
    res = om.estimation_frames(Y[:, 6:].T.copy(), U[:, 6:, :].transpose(1, 0, 2).copy())
    assert np.allclose(res["G"], g["G"][6:], rtol=1e-9)
    assert np.allclose(res["p"], g["p"][6:], rtol=1e-9, atol=1e-15)
    assert np.allclose(res["lambda_d"], g["lambda_d"][6:], rtol=1e-12, atol=1e-300)


def test_zelinski_postfilter_golden(cuda):
    from distantspeech_b200.beamformer.MicArray import MicArray
    from distantspeech_b200.postfilter.postfilter import PostFilter
    g = golden("zelinski.npz")
    pf = PostFilter(MicArray(arrayType="circular", r=0.05, M=8, n_fft=256), 256, 128, 256)
    for n in range(g["Z"].shape[2]):
        W = pf.getweights(g["Z"][:, :, n].astype(complex))
        assert W.shape == (129,)
        assert np.allclose(W, g["W"][n], rtol=1e-10, atol=1e-14)
    assert np.allclose(pf.Pxii, g["Pxii"], rtol=1e-12)
    assert np.allclose(pf.Pxij, g["Pxij"], rtol=1e-12, atol=1e-30)
    with pytest.raises(AttributeError):
        pf.process(None, None, None)


# ---------------------------------------------------------------- a18 (config 5)
@pytest.mark.parametrize("engine", ["simt", "tensor"])
def test_srp_angle_spectrum_vs_oracle(cuda, engine):
    from distantspeech_b200.beamformer.MicArray import MicArray
    from distantspeech_b200.doa.srp import srp
    geo = O.MicGeometry("circular", r=0.032, M=4, n_fft=512)
    x = np.ascontiguousarray(O.synth_streams(1, geo, 256 * 40, look_deg=(75.0, 0.0), interf_deg=(250.0, 0.0), seed0=9)[0].T)
    mic = MicArray(arrayType="circular", r=0.032, M=4, n_fft=512)
    P, p = srp(mic, engine=engine).compute_angle_spectrum(x)
    Pref, pref = O.srp_angle_spectrum(x.astype(np.float64), geo)
    assert P.shape == (360, 40) and p.shape == (257, 40)
    rel = np.max(np.abs(P - Pref) / np.abs(Pref))
    print("SRP %s: max rel err %.2e" % (engine, rel))
    assert rel <= 1e-3                                                       # SURVEY 8d tolerance for the map
    assert np.array_equal(np.argmax(P[:, 5:].sum(axis=1)), np.argmax(Pref[:, 5:].sum(axis=1)))
    assert np.mean(np.abs(p - pref) > 1e-9) < 0.01


def test_srp_grid_16mic_48k_tensor_vs_simt_vs_oracle(cuda):
    # config-5 shape at reduced grid: 16 mics, 48 kHz, n_fft 1024, az x el grid through compute_tau([az, el])
    from distantspeech_b200.beamformer.MicArray import MicArray
    from distantspeech_b200.doa.srp import srp
    mic = MicArray(arrayType="circular", r=0.05, M=16, n_fft=1024)
    mic.fs = 48000                                                           # the reference hard-wires 16 kHz (MicArray.py:27)
    mic.omega = 2 * np.pi * mic.freq_bin * mic.fs / mic.n_fft
    geo = O.MicGeometry("circular", r=0.05, M=16, n_fft=1024, fs=48000)
    assert np.allclose(geo.mic_loc, mic.mic_loc)
    x = np.ascontiguousarray(O.synth_streams(1, geo, 512 * 70, look_deg=(100.0, 20.0), interf_deg=(300.0, 5.0), seed0=12, fs=48000)[0].T)
    az, el = np.arange(0, 360, 3), np.arange(0, 90, 10)
    Pt = srp(mic, engine="tensor").compute_grid_spectrum(x, az, el)
    Ps = srp(mic, engine="simt").compute_grid_spectrum(x, az, el)
    assert Pt.shape == (120, 9, 70)
    assert np.max(np.abs(Pt - Ps) / np.abs(Ps)) <= 1e-3
    # oracle on 12 random frames x 40 random directions
    rng = np.random.default_rng(0)
    Y = O.Transform(channel=16, n_fft=1024, hop_length=512).stft(x.astype(np.float64))
    di = rng.choice(120 * 9, 40, replace=False)
    ti = rng.choice(70, 12, replace=False)
    tau = np.stack([O.method_tau(geo, np.array([az[i // 9], el[i % 9]]) * np.pi / 180)[:, 0] for i in di])
    Pref = O.srp_map(Y[:, ti, :], geo.omega, tau)
    got = Pt.reshape(-1, 70)[di][:, ti]
    rel = np.max(np.abs(got - Pref) / np.abs(Pref))
    print("SRP 16-mic grid: max rel err vs oracle %.2e" % rel)
    assert rel <= 1e-3
    flat = Pt.sum(axis=2)
    ia, ie = np.unravel_index(np.argmax(flat), flat.shape)
    assert abs(az[ia] - 100) <= 6                                            # finds the source azimuth


# ---------------------------------------------------------------- f2: PCM ingest / egress
def test_pcm16_ingest_matches_load_audio(cuda):
    import ctypes as C
    import torch
    from distantspeech_b200 import _lib as L
    from distantspeech_b200.beamformer.MicArray import MicArray
    from distantspeech_b200.pipelines import MvdrMcsppChain
    rng = np.random.default_rng(1)
    pcm = rng.integers(-32768, 32768, size=100003, dtype=np.int16)
    d = torch.from_numpy(pcm).cuda()
    out = torch.empty(pcm.shape[0], dtype=torch.float32, device="cuda")
    L.check(L.lib().ds_pcm16_to_float_run(pcm.shape[0], L.ptr(d), L.ptr(out), L.stream_ptr()))
    ref = pcm.astype(np.float32) / float(np.iinfo(np.int16).max)              # utils.py:184-185
    assert np.array_equal(out.cpu().numpy(), ref)
    back = torch.empty(pcm.shape[0], dtype=torch.int16, device="cuda")
    L.check(L.lib().ds_float_to_pcm16_run(pcm.shape[0], L.ptr(out), L.ptr(back), L.stream_ptr()))
    assert np.array_equal(back.cpu().numpy(), (ref * np.iinfo(np.int16).max).astype(np.int16))   # :193
    # load_audio / save_audio round trip through a wav file (utils.py:182-196)
    import os
    import tempfile
    from distantspeech_b200.beamformer.utils import load_audio, save_audio
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "rt")
        save_audio(path, ref[:4000].reshape(2000, 2), fs=16000)
        got = load_audio(path + ".wav")
    assert got.dtype == np.float32 and got.shape == (2000, 2)
    q = (ref[:4000] * 32767).astype(np.int16)
    assert np.array_equal(got.reshape(-1), q.astype(np.float32) / 32767.0)
    # int16 host buffers through the chain == float path on the dequantised signal
    geo = O.MicGeometry("circular", r=0.05, M=8, n_fft=512)
    xs = O.synth_streams(3, geo, 256 * 40, seed0=77)
    xi = np.round(xs * 32767).astype(np.int16)
    mic = MicArray(arrayType="circular", r=0.05, M=8, n_fft=512)
    ch = MvdrMcsppChain(mic, look_angle=(30, 0))
    y_pcm = ch.process_host(torch.from_numpy(xi).pin_memory(), slice_frames=9).numpy().copy()
    xf = xi.astype(np.float32) / 32767.0
    ref0 = O.mvdr_mcspp_chain(xf[0].T.astype(np.float64), geo, (30, 0), 512, 256)
    assert_wave_parity(ref0, y_pcm[0], "pcm16 chain")


# ---------------------------------------------------------------- the boundary from plain C
def test_capi_from_plain_c(cuda, tmp_path):
    """tests/capi/host_roundtrip.c: dlopen + cudaMalloc + POD structs, no torch -- STFT/ISTFT round trip and
    the error path through the C ABI."""
    import os
    import shutil
    import subprocess
    from distantspeech_b200 import _build
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    if shutil.which("gcc") is None or not os.path.exists(os.path.join(cuda_home, "include", "cuda_runtime_api.h")):
        pytest.skip("no C toolchain / CUDA headers on this box")
    exe = str(tmp_path / "host_roundtrip")
    subprocess.run(["gcc", "-O1", os.path.join(root, "tests", "capi", "host_roundtrip.c"), "-I", os.path.join(root, "include"),
                    "-I", os.path.join(cuda_home, "include"), "-L", os.path.join(cuda_home, "lib64"), "-lcudart", "-ldl", "-lm",
                    "-o", exe], check=True)
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(cuda_home, "lib64") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([exe, _build.LIB], capture_output=True, text=True, env=env, timeout=120)
    print(r.stdout, r.stderr)
    assert r.returncode == 0 and "PASS" in r.stdout


def test_transform_ragged_chunks_match_the_reference_quirk(cuda):
    """Quirk 4 (transform.py:441,451): a chunk that is not a multiple of hop loses its trailing samples but the
    history still takes the last `overlap` samples -- the device Transform must misalign exactly like the oracle;
    a chunk shorter than one hop raises (the reference fails inside util.frame)."""
    from distantspeech_b200.transform.transform import Transform
    rng = np.random.default_rng(17)
    x = (rng.standard_normal((4000, 2)) * 0.2).astype(np.float32)
    tf, ref = Transform(n_fft=512, hop_length=256, channel=2), O.Transform(channel=2, n_fft=512, hop_length=256)
    tfo, refo = Transform(n_fft=512, hop_length=256, channel=2), O.Transform(channel=2, n_fft=512, hop_length=256)
    pos = 0
    for n in (300, 1000, 256, 777, 1667):
        Y, Yr = tf.stft(x[pos:pos + n]), ref.stft(x[pos:pos + n].astype(np.float64))
        assert Y.shape == Yr.shape == (257, n // 256, 2)
        assert np.max(np.abs(Y - Yr)) <= 3e-6 * np.max(np.abs(Yr))
        assert np.array_equal(tf.previous_input, ref.previous_input)
        y, yr = tfo.istft(Yr), refo.istft(Yr)
        assert y.shape == yr.shape and np.max(np.abs(y - yr)) < 2e-6
        pos += n
    with pytest.raises(ValueError):
        Transform(n_fft=512, hop_length=256, channel=2).stft(x[:100])


# ---------------------------------------------------------------- f1: frequency-domain GSC + McMcra
def test_gsc_mcmcra_golden(cuda):
    from distantspeech_b200.beamformer.MicArray import MicArray
    from distantspeech_b200.beamformer.GSC import GSC
    from distantspeech_b200.noise_estimation.mc_mcra import McMcra
    g = golden("gsc.npz")
    n1 = int(g["n_first"])
    mic = MicArray(arrayType="circular", r=0.032, M=4)
    gsc = GSC(mic, 256)
    ya = gsc.process(g["x"][:, :n1], g["angle_rad"], method=2)
    yb = gsc.process(g["x"][:, n1:], g["angle_rad"], method=2)
    assert set(ya.keys()) == {"data", "WNG", "DI", "beampattern"}
    err, snr = assert_wave_parity(g["y"], np.concatenate([ya["data"], yb["data"]]), "GSC")
    print("GSC: max-abs %.2e SNR %.1f dB" % (err, snr))
    assert np.max(np.abs(gsc.spp.p - g["p_last"])) < 1e-4 and np.max(np.abs(gsc.spp.G - g["Gpost_last"])) < 1e-4
    assert np.linalg.norm(gsc.G - g["Gw_last"]) <= 1e-4 * np.linalg.norm(g["Gw_last"])
    assert np.linalg.norm(gsc.spp.Phi_vv - g["Phi_vv_last"]) <= 1e-5 * np.linalg.norm(g["Phi_vv_last"])
    # McMcra on its own, fed the oracle's spectrum: per-frame p / q / xi / gamma / G against the oracle
    geo = O.MicGeometry("circular", r=0.032, M=4, n_fft=256)
    D = O.Transform(channel=4, n_fft=256, hop_length=128).stft(g["x"].astype(np.float64).T)[:, :60]
    ref, taps = O.McMcra(nfft=256, channels=4), {k: [] for k in ("p", "q", "xi", "gamma", "G")}
    for n in range(D.shape[1]):
        ref.estimation(D[:, n, :])
        for k in taps:
            taps[k].append(getattr(ref, k).copy())
    est = McMcra(nfft=256, channels=4)
    res = est.estimation_frames(D)
    for k in taps:
        assert np.allclose(res[k], np.array(taps[k]).T, rtol=1e-8, atol=1e-12), k
    assert np.allclose(est.Phi_yy, ref.Phi_yy, rtol=1e-10, atol=1e-18) and est.Phi_vv.shape == (4, 4, 129)
    e2 = McMcra(nfft=256, channels=4)
    for n in range(8):
        e2.estimation(D[:, n, :])
    assert np.array_equal(e2.p, res["p"][:, 7])                                # frame by frame == many frames per call
    # method 0 passes channel 0 through (GSC.py:242); batch of streams == singles
    y0 = GSC(mic, 256).process(g["x"][:, :n1], g["angle_rad"], method=0)["data"]
    assert np.max(np.abs(y0[128:] - g["x"][0, :n1 - 128])) < 2e-6
    yb2 = GSC(mic, 256).process(np.stack([g["x"][:, :n1], g["x"][::-1, :n1]]), g["angle_rad"], method=2)["data"]
    assert np.max(np.abs(yb2[0] - ya["data"])) < 1e-6


@pytest.mark.parametrize("M", [2, 6, 8])
def test_gsc_other_mic_counts(cuda, M):
    from distantspeech_b200.beamformer.MicArray import MicArray
    from distantspeech_b200.beamformer.GSC import GSC
    geo = O.MicGeometry("circular", r=0.04, M=M, n_fft=256)
    x = O.synth_streams(1, geo, 128 * 90, seed0=40 + M)[0]
    ang = np.array([75, 0]) / 180 * np.pi
    ref = O.GscOracle(geo, 256).process(x.astype(np.float64), ang, method=2)
    y = GSC(MicArray(arrayType="circular", r=0.04, M=M), 256).process(x, ang, method=2)["data"]
    assert_wave_parity(ref, y, "GSC M=%d" % M)


def test_subband_gsc_golden(cuda):
    from distantspeech_b200.beamformer.MicArray import MicArray
    from distantspeech_b200.beamformer.SubbandGSC import SubbandGSC
    g = golden("subband_gsc.npz")
    n1 = int(g["n_first"])
    sg = SubbandGSC(MicArray(arrayType="circular", r=0.032, M=4), 256, angle=[30, 0])
    xa, xb = g["x"][:, :n1].copy(), g["x"][:, n1:].copy()
    ra, rb = sg.process(xa), sg.process(xb)
    assert len(ra) == 5 and ra[2].shape == (n1, 4) and ra[3].shape[0] == 257
    err, snr = assert_wave_parity(g["y"], np.concatenate([ra[0], rb[0]]), "SubbandGSC")
    print("SubbandGSC: max-abs %.2e SNR %.1f dB" % (err, snr))
    assert np.max(np.abs(np.concatenate([ra[1], rb[1]]) - g["fix_output"])) < 2e-6
    assert np.max(np.abs(np.concatenate([ra[2], rb[2]]) - g["bm_output"])) < 1e-4
    assert np.mean(np.abs(np.concatenate([ra[3], rb[3]], axis=1) - g["p"]) > 1e-3) < 0.01
    rel = np.linalg.norm(sg.aic_filter.W - g["W_aic_last"]) / np.linalg.norm(g["W_aic_last"])
    assert rel < 1e-3, rel
    assert np.linalg.norm(sg._bm_W(0) - g["W_bm0_last"]) < 1e-3 * np.linalg.norm(g["W_bm0_last"])
    assert not np.array_equal(xa, g["x"][:, :n1])                       # the caller's array was DC-notched in place
    # batch of two streams == singles
    rb2 = SubbandGSC(MicArray(arrayType="circular", r=0.032, M=4), 256, angle=[30, 0]).process(
        np.stack([g["x"][:, :n1], g["x"][::-1, :n1]]).copy())
    assert np.max(np.abs(rb2[0][0] - ra[0])) < 1e-6


def test_subband_lms_classes(cuda):
    """SubbandLMS / SubbandLmsMc block by block (time-domain in and out) against the oracle's filters."""
    from distantspeech_b200.adaptivefilter.SubbandLMS import SubbandLMS
    from distantspeech_b200.adaptivefilter.SubbandLmsMc import SubbandLmsMc
    rng = np.random.default_rng(5)
    x = rng.standard_normal((256 * 12, 3)) * 0.2
    d = 0.6 * np.roll(x[:, 0], 7) + 0.1 * rng.standard_normal(256 * 12)
    p = rng.uniform(0.1, 1.0, 257)
    f1, o1 = SubbandLMS(filter_len=2, num_bands=512, mu=0.1), O.SubbandNlms(2, 512, 1, mu=0.1)
    f3, o3 = SubbandLmsMc(filter_len=2, num_bands=512, channel=3, mu=0.01, alpha=0.8), O.SubbandNlms(2, 512, 3, mu=0.01, alpha=0.8)
    for n in range(12):
        sl = slice(256 * n, 256 * (n + 1))
        e1, W1 = f1.update(x[sl, 0], d[sl], p=p)
        e3, W3 = f3.update(x[sl], d[sl], p=p[:, None])
        r1, r3 = o1.update(x[sl, 0], d[sl], p), o3.update(x[sl], d[sl], p)
        assert np.max(np.abs(e1 - r1)) < 5e-6 and np.max(np.abs(e3 - r3)) < 5e-6
    fl, ol = SubbandLMS(filter_len=3, num_bands=512, mu=0.002, normalization=False), O.SubbandNlms(3, 512, 1, mu=0.002, normalization=False)
    for n in range(12):
        sl = slice(256 * n, 256 * (n + 1))
        assert np.max(np.abs(fl.update(x[sl, 0], d[sl])[0] - ol.update(x[sl, 0], d[sl], np.ones(257)))) < 5e-6
    assert W1.shape == (257, 2) and W3.shape == (257, 2, 3)
    assert np.linalg.norm(W1 - o1.W[:, :, 0]) < 1e-4 * np.linalg.norm(o1.W) and np.linalg.norm(W3 - o3.W) < 1e-4 * np.linalg.norm(o3.W)


def test_subband_rls_golden(cuda):
    from distantspeech_b200.adaptivefilter.SubbandRLS import SubbandRLS
    g = golden("subband_rls.npz")
    f = SubbandRLS(filter_len=2, num_bands=512)
    errs = [f.update(g["x"][256 * n:256 * (n + 1)], g["d"][256 * n:256 * (n + 1)])[0] for n in range(14)]
    assert_wave_parity(g["err"], np.concatenate(errs), "SubbandRLS")
    assert np.linalg.norm(f.W - g["W_last"]) < 1e-4 * np.linalg.norm(g["W_last"])
    assert np.linalg.norm(f.P - g["P_last"]) < 1e-4 * np.linalg.norm(g["P_last"])
    with pytest.raises(ValueError):
        SubbandRLS(filter_len=5)


def test_tdgsc_postfilter_golden(cuda):
    # postfilter=True (TDGSC.py:157-170): NsOmlsaMulti gain on the output through the streaming transforms, state carried
    # across two calls; batch == single
    from distantspeech_b200.beamformer.MicArray import MicArray
    from distantspeech_b200.beamformer.TDGSC import TDGSC
    g = golden("tdgsc_postfilter.npz")
    n1 = int(g["n_first"])
    td = TDGSC(MicArray(arrayType="circular", r=0.032, M=4), frameLen=256, angle=[30, 0])
    ra, rb = td.process(g["x"][:n1].copy(), postfilter=True), td.process(g["x"][n1:].copy(), postfilter=True)
    err, snr = assert_wave_parity(g["y"], np.concatenate([ra[0], rb[0]]), "TDGSC postfilter")
    print("TDGSC postfilter: max-abs %.2e SNR %.1f dB" % (err, snr))
    assert np.max(np.abs(td.omlsa_multi.G - g["G_last"])) < 1e-3
    xb = np.stack([g["x"], g["x"][::-1].copy()])
    yb = TDGSC(MicArray(arrayType="circular", r=0.032, M=4), frameLen=256, angle=[30, 0]).process(xb.copy(), postfilter=True)[0]
    y0 = TDGSC(MicArray(arrayType="circular", r=0.032, M=4), frameLen=256, angle=[30, 0]).process(g["x"].copy(), postfilter=True)[0]
    assert np.max(np.abs(yb[0] - y0)) < 1e-6


def test_tdgsc_golden(cuda):
    from distantspeech_b200.beamformer.MicArray import MicArray
    from distantspeech_b200.beamformer.TDGSC import TDGSC
    g = golden("tdgsc.npz")
    n1 = int(g["n_first"])
    td = TDGSC(MicArray(arrayType="circular", r=0.032, M=4), frameLen=256, angle=[30, 0])
    ra, rb = td.process(g["x"][:n1].copy()), td.process(g["x"][n1:].copy())
    assert len(ra) == 3 and ra[2].shape == (n1, 3) and ra[1].shape[0] == 257
    err, snr = assert_wave_parity(g["y"], np.concatenate([ra[0], rb[0]]), "TDGSC")
    print("TDGSC: max-abs %.2e SNR %.1f dB" % (err, snr))
    assert np.max(np.abs(np.concatenate([ra[2], rb[2]]) - g["bm_output"])) < 2e-6
    assert np.mean(np.abs(np.concatenate([ra[1], rb[1]], axis=1) - g["p"]) > 1e-6) < 0.01
    assert np.linalg.norm(td.aic_filter.W - g["W_last"]) < 1e-3 * np.linalg.norm(g["W_last"])
    # GSC.process1: the same chain ungated and causal (GSC.py:151-172)
    from distantspeech_b200.beamformer.GSC import GSC
    geo4 = O.MicGeometry("circular", r=0.032, M=4, n_fft=256)
    o1 = O.TdgscOracle(geo4, 256, np.array([30, 0]) / 180 * np.pi, gated=False, non_causal=False)
    g1 = GSC(MicArray(arrayType="circular", r=0.032, M=4), 256, angle=[30, 0])
    y1 = np.concatenate([g1.process1(g["x"][:n1].copy()), g1.process1(g["x"][n1:].copy())])
    r1 = np.concatenate([o1.process(g["x"][:n1].astype(np.float64))[0], o1.process(g["x"][n1:].astype(np.float64))[0]])
    assert_wave_parity(r1, y1, "GSC.process1")
    # other microphone counts against the oracle; batch == singles
    geo = O.MicGeometry("linear", r=0.04, M=6, n_fft=256)
    x6 = np.ascontiguousarray(O.synth_streams(2, geo, 256 * 40, look_deg=(75.0, 0.0), seed0=61).transpose(0, 2, 1))
    ang = np.array([75, 0]) / 180 * np.pi
    yb = TDGSC(MicArray(arrayType="linear", r=0.04, M=6), frameLen=256, angle=[75, 0]).process(x6.copy())[0]
    for s in range(2):
        assert_wave_parity(O.TdgscOracle(geo, 256, ang).process(x6[s].astype(np.float64))[0], yb[s], "TDGSC 6 mics stream %d" % s)
