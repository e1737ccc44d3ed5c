"""GPU: round-2 parity pins -- the CUDA path against goldens dumped from the unmodified reference
(tests/golden/make_golden_r2.py): config 1 on a 4-mic linear array and on the shipped recording, SRP-PHAT map."""
import numpy as np
import pytest

from conftest import golden, assert_wave_parity
from oracle import np_oracle as O

pytestmark = pytest.mark.gpu


def test_adaptive_mvdr_linear_golden_gpu(cuda):
    """Config 1 as BASELINE.json words it: 4-mic linear array, n_fft 512 / hop 256, 10 s (adaptivebeamformer.py:44-128)."""
    from distantspeech_b200.beamformer.MicArray import MicArray
    from distantspeech_b200.beamformer.adaptivebeamformer import adaptivebeamfomer
    g = golden("adaptive_mvdr_linear.npz")
    mic = MicArray(arrayType="linear", r=0.032, M=4, n_fft=512)
    ab = adaptivebeamfomer(mic, 512, 256, 512)
    out = ab.process(g["x"], g["angle_rad"], method=2)
    err, s = assert_wave_parity(g["y"], out["data"], "adaptive MVDR, linear array")
    print("config 1 (linear, 10 s): max-abs %.2e SNR %.1f dB" % (err, s))
    assert np.array_equal(ab.mcra.p, g["p_last"])                       # MCRA decisions identical after 625 frames
    assert np.allclose(ab.H, g["H_last"], rtol=1e-3, atol=1e-6)


def test_adaptive_mvdr_rec1_golden_gpu(cuda):
    """Config 1 as example/run_MVDRbeamformer.py runs it: the shipped 4-channel recording (first 10 s), int16 PCM
    scaled on the device like load_audio (utils.py:182-187), circular r = 0.032, n_fft 256 / hop 128, look 197 deg."""
    from distantspeech_b200.beamformer.MicArray import MicArray
    from distantspeech_b200.beamformer.adaptivebeamformer import adaptivebeamfomer
    from distantspeech_b200.beamformer.utils import pcm16_to_float
    g = golden("adaptive_mvdr_rec1.npz")
    x = pcm16_to_float(g["pcm"])                                         # [4, N] float32
    assert np.array_equal(x, g["pcm"].astype(np.float32) / np.float32(32767.0))
    mic = MicArray(arrayType="circular", r=0.032, M=4)
    ab = adaptivebeamfomer(mic, 256, 128, 256)
    out = ab.process(x, g["angle_rad"], method=2)
    err, s = assert_wave_parity(g["y"], out["data"], "adaptive MVDR, rec1")
    print("config 1 (rec1, 10 s): max-abs %.2e SNR %.1f dB" % (err, s))
    assert np.array_equal(ab.mcra.p, g["p_last"])


@pytest.mark.parametrize("engine", ["tensor", "simt"])
def test_srp_reference_golden_gpu(cuda, engine):
    """a18 against the reference itself: srp.compute_angle_spectrum (doa/srp.py:17-53) on a 4-mic circular array."""
    from distantspeech_b200.beamformer.MicArray import MicArray
    from distantspeech_b200.doa.srp import srp
    g = golden("srp.npz")
    mic = MicArray(arrayType="circular", r=0.032, M=4, n_fft=256)
    P, p = srp(mic, engine=engine).compute_angle_spectrum(g["x"])
    ref = g["angle_spectrum"]
    assert P.shape == ref.shape and p.shape == g["p"].shape
    rel = np.max(np.abs(P - ref)) / np.max(np.abs(ref))
    print("SRP %s vs reference golden: max err / max %.2e" % (engine, rel))
    assert rel <= 1e-3 and np.max(np.abs(P - ref) / np.abs(ref)) <= 2e-3
    assert np.mean(np.argmax(P, axis=0) == np.argmax(ref, axis=0)) >= 0.95      # per-frame argmax (ties within 1e-3 aside)
    assert np.array_equal(np.argmax(P.sum(axis=1)), np.argmax(ref.sum(axis=1)))
    assert np.mean(np.abs(p - g["p"]) > 1e-9) < 0.01
    Pn, _ = srp(mic, engine=engine).compute_angle_spectrum(g["x"], phat=False)
    reln = np.max(np.abs(Pn - g["angle_spectrum_nophat"])) / np.max(np.abs(g["angle_spectrum_nophat"]))
    assert reln <= 1e-3


# ---------------------------------------------------------------- f2: int16 ingest / egress fused into the transform kernels
def test_fused_pcm16_transforms_match_the_separate_conversions(cuda):
    """ds_stft_pcm16_run == ds_pcm16_to_float_run + ds_stft_run bit for bit (every mode, chunked streaming, odd offsets);
    ds_istft_pcm16_run == ds_istft_run + save_audio's expression on the float64 array Transform.istft returns."""
    import ctypes as C
    import torch
    from distantspeech_b200 import _lib as L
    from distantspeech_b200.transform.transform import _sqrt_hann
    rng = np.random.default_rng(5)
    for n_fft, hop, mode, N in ((512, 256, L.DS_STFT_STREAMING, 256 * 9), (256, 64, L.DS_STFT_STREAMING, 64 * 21),
                                (512, 128, L.DS_STFT_CENTER, 3001), (256, 128, L.DS_STFT_PLAIN, 1999)):
        S, M, K = 3, 2, n_fft // 2 + 1
        pcm = torch.from_numpy(rng.integers(-30000, 30000, size=(S, M, N), dtype=np.int16)).cuda()
        xf = torch.empty((S, M, N), dtype=torch.float32, device="cuda")
        L.check(L.lib().ds_pcm16_to_float_run(pcm.numel(), L.ptr(pcm), L.ptr(xf), L.stream_ptr()))
        assert np.array_equal(xf.cpu().numpy(), pcm.cpu().numpy().astype(np.float32) / np.float32(32767.0))
        win = L.device_window(_sqrt_hann(n_fft), n_fft)
        p = L.StftParams(n_fft, hop, S, M, N, mode, 0, 0)
        T = L.lib().ds_stft_num_frames(C.byref(p))
        ov = n_fft - hop
        h1 = torch.full((S, M, ov), 0.25, dtype=torch.float32, device="cuda")
        h2 = h1.clone()
        X1 = torch.empty((S, T, M, K), dtype=torch.complex64, device="cuda")
        X2 = torch.empty_like(X1)
        L.check(L.lib().ds_stft_run(C.byref(p), L.ptr(win), L.ptr(h1), L.ptr(xf), L.ptr(X1), L.stream_ptr()))
        L.check(L.lib().ds_stft_pcm16_run(C.byref(p), L.ptr(win), L.ptr(h2), L.ptr(pcm), L.ptr(X2), L.stream_ptr()))
        assert torch.equal(torch.view_as_real(X1), torch.view_as_real(X2)) and torch.equal(h1, h2), (n_fft, hop, mode)
        if mode == L.DS_STFT_CENTER:
            continue
        # synthesis: int16 store == float store followed by (float64(y) * 32767).astype(int16)
        ip = L.IstftParams(n_fft, hop, S, M, T, mode, 0, 0, 0.5 * hop / float(np.sum(_sqrt_hann(n_fft) ** 2)))
        n_out = T * hop if mode == L.DS_STFT_STREAMING else n_fft + hop * (T - 1)
        t1 = torch.zeros((S, M, ov), dtype=torch.float32, device="cuda")
        t2 = t1.clone()
        yf = torch.empty((S, M, n_out), dtype=torch.float32, device="cuda")
        yi = torch.empty((S, M, n_out), dtype=torch.int16, device="cuda")
        L.check(L.lib().ds_istft_run(C.byref(ip), L.ptr(win), L.ptr(t1), L.ptr(X1), L.ptr(yf), L.stream_ptr()))
        L.check(L.lib().ds_istft_pcm16_run(C.byref(ip), L.ptr(win), L.ptr(t2), L.ptr(X1), L.ptr(yi), L.stream_ptr()))
        ref = (yf.cpu().numpy().astype(np.float64) * np.iinfo(np.int16).max).astype(np.int16)       # utils.py:193
        assert np.abs(yf.cpu().numpy()).max() < 1.0
        assert np.array_equal(yi.cpu().numpy(), ref) and torch.equal(t1, t2), (n_fft, hop, mode)
    # every int16 value through the fused scaling (non-overlapping frames, so each value is read exactly once)
    allv = torch.arange(-32768, 32768, dtype=torch.int32).to(torch.int16).reshape(1, 1, 65536).cuda()
    xf = torch.empty((1, 1, 65536), dtype=torch.float32, device="cuda")
    L.check(L.lib().ds_pcm16_to_float_run(65536, L.ptr(allv), L.ptr(xf), L.stream_ptr()))
    p = L.StftParams(128, 128, 1, 1, 65536, L.DS_STFT_PLAIN, 1, 1)               # fp64 FFT, complex128 out: injective enough
    win = L.device_window(np.ones(128), 128)
    Xa = torch.empty((1, 512, 1, 65), dtype=torch.complex128, device="cuda")
    Xb = torch.empty_like(Xa)
    L.check(L.lib().ds_stft_run(C.byref(p), L.ptr(win), None, L.ptr(xf), L.ptr(Xa), L.stream_ptr()))
    L.check(L.lib().ds_stft_pcm16_run(C.byref(p), L.ptr(win), None, L.ptr(allv), L.ptr(Xb), L.stream_ptr()))
    assert torch.equal(torch.view_as_real(Xa), torch.view_as_real(Xb))
    # float64 save path: product in double
    y64 = torch.from_numpy(rng.standard_normal(50001) * 0.3).cuda()
    out = torch.empty(50001, dtype=torch.int16, device="cuda")
    L.check(L.lib().ds_double_to_pcm16_run(y64.numel(), L.ptr(y64), L.ptr(out), L.stream_ptr()))
    assert np.array_equal(out.cpu().numpy(), (y64.cpu().numpy() * 32767).astype(np.int16))


def test_chain_pcm16_in_and_out(cuda):
    """MvdrMcsppChain with int16 PCM on both sides (device and host entry points): identical to the float path on the
    dequantised signal followed by save_audio's quantisation; within the waveform contract of the oracle."""
    import torch
    from distantspeech_b200.beamformer.MicArray import MicArray
    from distantspeech_b200.pipelines import MvdrMcsppChain
    geo = O.MicGeometry("circular", r=0.05, M=8, n_fft=512)
    xs = O.synth_streams(5, geo, 256 * 40, seed0=78)
    xi = np.round(xs * 32767).astype(np.int16)
    xf = xi.astype(np.float32) / np.float32(32767.0)
    mic = MicArray(arrayType="circular", r=0.05, M=8, n_fft=512)
    yf = MvdrMcsppChain(mic, look_angle=(30, 0)).process_device(torch.from_numpy(xf).cuda())
    yi = torch.empty((5, 256 * 40), dtype=torch.int16, device="cuda")
    MvdrMcsppChain(mic, look_angle=(30, 0)).process_device(torch.from_numpy(xi).cuda(), out=yi)
    q = (yf.cpu().numpy().astype(np.float64) * 32767).astype(np.int16)
    assert np.array_equal(yi.cpu().numpy(), q)
    ch = MvdrMcsppChain(mic, look_angle=(30, 0))
    y_host = torch.empty((5, 256 * 40), dtype=torch.int16).pin_memory()
    for _ in range(2):                                                   # second call reuses the staging pipeline
        ch.process_host(torch.from_numpy(xi).pin_memory(), y_host, slice_frames=11)
        assert np.array_equal(y_host.numpy(), q)
    yh32 = ch.process_host(torch.from_numpy(xi).pin_memory(), slice_frames=40)         # int16 in, float32 out
    assert np.array_equal(yh32.numpy(), yf.cpu().numpy())
    ref0 = O.mvdr_mcspp_chain(xf[3].T.astype(np.float64), geo, (30, 0), 512, 256)
    assert_wave_parity(ref0, yf[3].cpu().numpy(), "pcm16 chain (float out)")
    err = np.max(np.abs(q[3].astype(np.float64) / 32767.0 - ref0))
    assert err <= 1.0 / 32767 + 1e-4                                     # one LSB of the 16-bit output format
    with pytest.raises(ValueError):
        ch.process_device_profiled(torch.from_numpy(xf[:, :, :100]).cuda())     # profiled entry validates its input
    with pytest.raises(ValueError):
        ch.process_device(torch.from_numpy(xf).cuda().double())


# ---------------------------------------------------------------- f3: McSpp beyond 4 microphones, repeat=True
@pytest.mark.parametrize("tag,M,rep", [("m6", 6, False), ("m5", 5, False), ("m4r", 4, True)])
def test_mcspp_more_than_4_mics_golden_gpu(cuda, tag, M, rep):
    """McSpp with 5 / 6 microphones against the reference run with McCDR(nfft, channels=M) handed in
    (oracle/ref_harness.make_mcspp; mcspp.py:54 hard-wires 4), and estimation(repeat=True) (mcspp.py:282-284)."""
    from distantspeech_b200.noise_estimation.mcspp import McSpp
    g = golden("mcspp_cdr_m68.npz")
    D = O.Transform(channel=M, n_fft=256, hop_length=128).stft(g[tag + "_x"].astype(np.float64))     # the reference's spectrum
    est = McSpp(nfft=256, channels=M)
    res = est.estimation_frames(D, repeat=rep)
    for k in ("p", "xi", "gamma", "q"):
        assert np.allclose(res[k], g[tag + "_" + k], rtol=1e-7, atol=1e-11), (tag, k, np.max(np.abs(res[k] - g[tag + "_" + k])))
    scale = lambda a: np.max(np.abs(a))                                                           # noqa: E731
    for nm, a in (("w_last", est.w), ("Phi_vv_last", est.Phi_vv), ("Phi_vv_inv_last", est.Phi_vv_inv)):
        assert a.shape == g[tag + "_" + nm].shape
        assert np.max(np.abs(a - g[tag + "_" + nm])) <= 1e-8 * scale(g[tag + "_" + nm]), (tag, nm)
    # frame by frame == many frames per launch
    e2 = McSpp(nfft=256, channels=M)
    for n in range(12):
        p = e2.estimation(D[:, n, :], repeat=rep)
    assert np.array_equal(p, res["p"][:, 11])


def test_mcspp_8_mics_matches_the_oracle_gpu(cuda):
    """8 microphones: the patched reference raises LinAlgError in frames 5..6 (unloaded fallback inverse of a rank-deficient
    Phi_yy, mcspp.py:224-227); the drop-in and the oracle keep the loading until Phi_yy has full rank (deviations)."""
    from distantspeech_b200.noise_estimation.mcspp import McSpp
    geo = O.MicGeometry("circular", r=0.05, M=8, n_fft=512)
    x = np.ascontiguousarray(O.synth_streams(1, geo, 256 * 160, seed0=0xC8)[0].T)
    D = O.Transform(channel=8, n_fft=512, hop_length=256).stft(x.astype(np.float64))
    ref, taps = O.McSpp(nfft=512, channels=8), {k: [] for k in ("p", "xi", "q")}
    with np.errstate(all="ignore"):
        for n in range(D.shape[1]):
            ref.estimation(D[:, n, :])
            for k in taps:
                taps[k].append(getattr(ref, k).copy())
    est = McSpp(nfft=512, channels=8)
    res = est.estimation_frames(D)
    assert np.all(np.isfinite(res["p"]))
    for k in taps:
        assert np.allclose(res[k], np.array(taps[k]).T, rtol=1e-6, atol=1e-10), (k, np.max(np.abs(res[k] - np.array(taps[k]).T)))
    assert np.max(np.abs(est.w - ref.w)) <= 1e-7 * np.max(np.abs(ref.w))


# ---------------------------------------------------------------- f4: adaptive WPE
def test_wpe_golden_gpu(cuda):
    """Wpe (dereverberation/awpe.py:128-191) against the filter state of the reference's own update body (ref_harness.make_wpe)
    and against the oracle's dereverberated waveform; block by block == one call; batch == singles."""
    from distantspeech_b200.dereverberation.awpe import Wpe
    g = golden("wpe.npz")
    C, Lf, nb, hop, D = (int(v) for v in g["params"])
    x = g["x"]
    w = Wpe(channels=C, filter_len=Lf, num_bands=nb, delay=D, hop_length=hop)
    outs = []
    for n in range(60):
        y0, W = w.update(x[n * hop:(n + 1) * hop])
        outs.append(y0)
        if n + 1 in (30, 60):
            for nm, a in (("W", w.W), ("P", w.P), ("var", w.var)):
                ref = g["%s%d" % (nm, n + 1)]
                assert a.shape == ref.shape, (nm, a.shape, ref.shape)
                assert np.max(np.abs(a - ref)) <= 2e-5 * np.max(np.abs(ref)), (nm, n + 1, np.max(np.abs(a - ref)) / np.max(np.abs(ref)))
    ref_y = O.WpeOracle(channels=C, filter_len=Lf, num_bands=nb, delay=D, hop_length=hop).process(x.astype(np.float64))
    assert_wave_parity(ref_y[:, 0], np.concatenate(outs), "WPE, block by block")
    w2 = Wpe(channels=C, filter_len=Lf, num_bands=nb, delay=D, hop_length=hop)
    y = w2.process(x)
    assert y.shape == x.shape
    for c in range(C):
        assert_wave_parity(ref_y[:, c], y[:, c], "WPE channel %d" % c)
    assert np.max(np.abs(y[:, 0] - np.concatenate(outs))) < 1e-6
    # fed the reference's spectrum (complex128) the recursion itself agrees to rounding
    import torch
    from distantspeech_b200 import _lib as L
    Dspec = O.Transform(n_fft=nb, hop_length=hop, channel=C).stft(x.astype(np.float64))              # [K, T, C]
    w3 = Wpe(channels=C, filter_len=Lf, num_bands=nb, delay=D, hop_length=hop)
    w3._ensure(1)
    X = torch.as_tensor(np.ascontiguousarray(Dspec.transpose(1, 2, 0)[None])).cuda()                 # [1, T, C, K] c128
    Err = torch.empty_like(X)
    L.check(L.lib().ds_wpe_run(1, nb // 2 + 1, X.shape[1], C, Lf, D, 0.998, 0.98, L.ptr(w3._state), L.ptr(X), 1, L.ptr(Err), L.stream_ptr()))
    assert np.max(np.abs(w3.W - g["W60"])) <= 1e-9 * np.max(np.abs(g["W60"]))
    assert np.max(np.abs(w3.P - g["P60"])) <= 1e-9 * np.max(np.abs(g["P60"]))
    yb = Wpe(channels=C, filter_len=Lf, num_bands=nb, delay=D, hop_length=hop).process(np.stack([x, x[::-1].copy()]))
    assert np.max(np.abs(yb[0] - y)) < 1e-7


def test_realtime_chunks_equal_one_call_gpu(cuda):
    """8f.2 chunk API: 1024-frame int16 capture buffers through realtime_processing.process_pcm with the GSC pipeline as the
    enhancement method == the same signal processed in one call (state carries over from chunk to chunk)."""
    from distantspeech_b200.beamformer.MicArray import MicArray
    from distantspeech_b200.beamformer.GSC import GSC
    from distantspeech_b200.realtime.realtime_processing import realtime_processing
    geo = O.MicGeometry("circular", r=0.032, M=4, n_fft=256)
    x4 = O.synth_streams(1, geo, 1024 * 12, seed0=0x77)[0]                                # [4, N] float32
    pcm = np.zeros((1024 * 12, 6), dtype=np.int16)
    pcm[:, 1:5] = np.round(x4.T * 32768).astype(np.int16)
    ang = np.array([30, 0]) / 180 * np.pi

    class Wrap(object):                                  # the reference calls EnhancementMethod.process(data) with one argument
        def __init__(self):
            self.g = GSC(MicArray(arrayType="circular", r=0.032, M=4), 256)

        def process(self, data):
            return self.g.process(np.ascontiguousarray(data.T), ang, method=2)
    rt = realtime_processing(EnhancementMehtod=Wrap(), chunk=1024, channels=6)
    chunks = [np.frombuffer(rt.process_pcm(pcm[i * 1024:(i + 1) * 1024].tobytes()), dtype='<i2') for i in range(12)]
    whole = Wrap().process(pcm[:, 1:5].astype(np.float32) / 32768.0)["data"]
    ref = (whole.astype(np.float32) * 32768).astype('<i2')
    assert np.max(np.abs(np.concatenate(chunks).astype(np.int32) - ref.astype(np.int32))) <= 1


# ---------------------------------------------------------------- a17: the FDGSC kernel pipeline against the fused kernel and the goldens
@pytest.mark.parametrize("precision", ["fp32", "fp64"])
def test_fdgsc_pipeline_equals_fused_and_golden_gpu(cuda, precision):
    from distantspeech_b200.beamformer.MicArray import MicArray
    from distantspeech_b200.beamformer.FDGSC import FDGSC
    g = golden("fdgsc.npz")
    mic = MicArray(arrayType="linear", r=0.05, M=6, n_fft=256)
    ang = [int(g["angle_deg"][0]), int(g["angle_deg"][1])]
    res = {}
    for impl in ("pipeline", "fused"):
        fd = FDGSC(mic, frameLen=256, angle=ang, precision=precision, impl=impl)
        r = fd.process(g["x"].copy(), postfilter=False, dc_notch=True)
        err, s = assert_wave_parity(g["y"], r[0], "FDGSC %s %s" % (impl, precision))
        print("FDGSC %s %s: max-abs %.2e SNR %.1f dB" % (impl, precision, err, s))
        res[impl] = (r, fd.aic_filter.W.copy(), np.stack([f.W[:, 0] for f in fd.bm]))
    tol = 2e-6 if precision == "fp32" else 1e-10
    a, b = res["pipeline"], res["fused"]
    assert np.max(np.abs(a[0][0] - b[0][0])) < tol                      # output
    assert np.max(np.abs(a[0][2] - b[0][2])) < tol and np.max(np.abs(a[0][4] - b[0][4])) < tol        # fixed beam, blocking outputs
    assert np.array_equal(a[0][1], b[0][1])                              # adaptation-control p: same decisions
    assert np.linalg.norm(a[1] - b[1]) <= (1e-4 if precision == "fp32" else 1e-9) * np.linalg.norm(b[1])
    assert np.linalg.norm(a[2] - b[2]) <= (1e-4 if precision == "fp32" else 1e-9) * np.linalg.norm(b[2])
    # same state blob: a stream may switch implementation between calls
    n1 = 256 * 25
    fd = FDGSC(mic, frameLen=256, angle=ang, precision=precision, impl="pipeline")
    ya = fd.process(g["x"][:n1].copy())[0]
    fd.impl = "fused"
    yb = fd.process(g["x"][n1:].copy())[0]
    assert np.max(np.abs(np.concatenate([ya, yb]) - a[0][0])) < 10 * tol
    fd2 = FDGSC(mic, frameLen=256, angle=ang, precision=precision, impl="fused")
    ya = fd2.process(g["x"][:n1].copy())[0]
    fd2.impl = "pipeline"
    yb = fd2.process(g["x"][n1:].copy())[0]
    assert np.max(np.abs(np.concatenate([ya, yb]) - a[0][0])) < 10 * tol


def test_fdgsc_pipeline_batch_device_entry_gpu(cuda):
    """process_device (output only) on a batch with 4, 6 and 8 microphones and a look direction with non-zero alignment delays."""
    import torch
    from distantspeech_b200.beamformer.MicArray import MicArray
    from distantspeech_b200.beamformer.FDGSC import FDGSC
    for M, look in ((4, 75.0), (6, 20.0), (8, 140.0)):
        geo = O.MicGeometry("linear", r=0.05, M=M, n_fft=256)
        xs = O.synth_streams(3, geo, 256 * 40, look_deg=(look, 0.0), interf_deg=(look + 70.0, 0.0), seed0=60 + M)   # [S, M, N]
        mic = MicArray(arrayType="linear", r=0.05, M=M, n_fft=256)
        fd = FDGSC(mic, frameLen=256, angle=[int(look), 0])
        xd = torch.from_numpy(xs).cuda()
        y = fd.process_device(xd).cpu().numpy()
        for s in (0, 2):
            ref = O.FdgscOracle(geo, 256, np.array([look, 0]) / 180 * np.pi).process(xs[s].T.astype(np.float64))
            assert_wave_parity(ref[0], y[s], "FDGSC pipeline M=%d stream %d" % (M, s))
            assert np.max(np.abs(xd[s].cpu().numpy().T - ref[4])) < 1e-6      # the caller's tensor holds the notched signal


# ---------------------------------------------------------------- n1: multi-beam weighting on the tensor cores
@pytest.mark.parametrize("M,B,weight", [(8, 64, "SD"), (4, 16, "DS"), (16, 130, "SD")])
def test_multibeam_tensor_core_gpu(cuda, M, B, weight):
    """FixedBeamformer.process_multibeam(engine="tensor"): per-bin GEMMs on tcgen05 (tf32 head + tail operands) against
    O.fixed_beamform per beam (fixedbeamformer.py:109-165), against the CUDA-core path, and chunked == one call."""
    from distantspeech_b200.beamformer.MicArray import MicArray
    from distantspeech_b200.beamformer.fixedbeamformer import FixedBeamformer
    geo = O.MicGeometry("circular", r=0.05, M=M, n_fft=512)
    xs = O.synth_streams(3, geo, 256 * 70, seed0=0x4B + M)
    x_nm = np.ascontiguousarray(xs.transpose(0, 2, 1))                    # [S, N, M]
    mic = MicArray(arrayType="circular", r=0.05, M=M, n_fft=512)
    angles = [(int(a), 0) for a in np.linspace(0, 359, B)]
    fb = FixedBeamformer(mic, 512, 256, 512)
    y = fb.process_multibeam(x_nm, angles, weightType=weight, engine="tensor")
    assert y.shape == (3, B, 256 * 70)
    worst = 1e9
    for s, b in ((0, 0), (1, B // 2), (2, B - 1), (0, B // 3)):
        W = O.fixed_weights(geo, 512, angles[b], weight)
        ref = O.fixed_beamform(x_nm[s].astype(np.float64), W, 512, 256)
        err, snr = assert_wave_parity(ref, y[s, b], "tensor-core beam %d of stream %d" % (b, s))
        worst = min(worst, snr)
    print("multi-beam tcgen05 M=%d B=%d: worst SNR vs oracle %.1f dB" % (M, B, worst))
    assert worst >= 100.0                                                 # head + tail operands: fp32-level accuracy, not tf32's ~70 dB
    if B <= 16:                                                           # the fused CUDA-core kernel holds every beam in shared memory
        ys = FixedBeamformer(mic, 512, 256, 512).process_multibeam(x_nm, angles, weightType=weight, engine="simt")
        assert np.max(np.abs(ys - y)) < 2e-6
    fb2 = FixedBeamformer(mic, 512, 256, 512)
    ya = fb2.process_multibeam(x_nm[:, :256 * 30], angles, weightType=weight, engine="tensor")
    yb = fb2.process_multibeam(x_nm[:, 256 * 30:], angles, weightType=weight, engine="tensor")
    assert np.max(np.abs(np.concatenate([ya, yb], axis=2) - y)) < 1e-6
    with pytest.raises(ValueError):
        FixedBeamformer(MicArray(arrayType="circular", r=0.05, M=6, n_fft=512), 512, 256, 512).process_multibeam(
            x_nm[:, :, :6], angles, engine="tensor")
