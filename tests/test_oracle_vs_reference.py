"""LIVE pinning of the oracle: the NumPy restatement (oracle/np_oracle.py) against the UNMODIFIED reference imported from
/root/reference through oracle/ref_harness.py, on fresh random inputs (the committed fixtures in tests/golden were made the
same way).  Runs only where the reference tree exists (the build container); skipped on the GPU box."""
import contextlib
import io
import warnings

import numpy as np
import pytest

from oracle import np_oracle as O
from oracle import ref_harness as H

pytestmark = [pytest.mark.skipif(not H.reference_available(), reason="reference tree not present"),
              pytest.mark.filterwarnings("ignore")]
warnings.filterwarnings("ignore")


@pytest.fixture(scope="module")
def ref():
    H.install()
    return H


def _quiet():
    return contextlib.redirect_stdout(io.StringIO())


def test_transform_and_mcra_live(ref):
    from DistantSpeech.transform.transform import Transform, stft, istft
    from DistantSpeech.noise_estimation.mcra import NoiseEstimationMCRA
    rng = np.random.default_rng(101)
    x = rng.standard_normal((256 * 20, 3)) * 0.2
    tr, to = Transform(n_fft=512, hop_length=256, channel=3), O.Transform(channel=3, n_fft=512, hop_length=256)
    for lo, hi in ((0, 256 * 7), (256 * 7, 256 * 20)):
        Yr, Yo = tr.stft(x[lo:hi]), to.stft(x[lo:hi])
        assert np.array_equal(Yr, Yo)
        assert np.array_equal(tr.istft(Yr), to.istft(Yo))
    w = O.sqrt_hann(512)
    D = stft(x[:, 0], n_fft=512, hop_length=128, window=w, center=True)
    assert np.array_equal(D, O.stft(x[:, 0], n_fft=512, hop_length=128, window=w, center=True))
    assert np.array_equal(istft(D, hop_length=128, window=w, center=True, length=4000),
                          O.istft(D, hop_length=128, window=w, center=True, length=4000))
    mr, mo = NoiseEstimationMCRA(nfft=512), O.Mcra(nfft=512)
    P = np.abs(Yr[:, :, 0]) ** 2
    for n in range(P.shape[1]):
        assert np.array_equal(mr.estimation(P[:, n]), mo.estimation(P[:, n])) and np.array_equal(mr.p, mo.p)


def test_chain_b_and_mcspp_live(ref):
    from DistantSpeech.transform.transform import Transform
    from DistantSpeech.noise_estimation.mcspp_base import McSppBase
    from DistantSpeech.noise_estimation.mcspp import McSpp
    geo = O.MicGeometry("circular", r=0.05, M=8, n_fft=512)
    x = O.synth_streams(1, geo, 256 * 40, seed0=202)[0].T.astype(np.float64)
    D = Transform(n_fft=512, hop_length=256, channel=8).stft(x)
    er, eo = McSppBase(nfft=512, channels=8), O.McSppBase(nfft=512, channels=8)
    for n in range(D.shape[1]):
        pr, po = er.estimation(D[:, n, :]), eo.estimation(D[:, n, :])
        assert np.allclose(pr, po, rtol=1e-9, atol=1e-12)
    assert np.allclose(er.Phi_vv, eo.Phi_vv, rtol=1e-9, atol=1e-15)
    with _quiet():
        cr = McSpp(nfft=512, channels=4)
    co = O.McSpp(nfft=512, channels=4)
    for n in range(D.shape[1]):
        assert np.array_equal(cr.estimation(D[:, n, :4]), co.estimation(D[:, n, :4]))
    assert np.array_equal(cr.w, co.w)


def test_gsc_family_live(ref):
    from DistantSpeech.beamformer.MicArray import MicArray
    from DistantSpeech.beamformer.TDGSC import TDGSC
    geo = O.MicGeometry("circular", r=0.032, M=4, n_fft=256)
    mic = MicArray(arrayType="circular", r=0.032, M=4)
    ang = np.array([30, 0]) / 180 * np.pi
    x = O.synth_streams(1, geo, 256 * 24 + 33, seed0=303)[0].astype(np.float64)          # [M, N]
    g = H.make_gsc(mic, 256)
    assert np.max(np.abs(g.process(x.copy(), ang, method=2)["data"] - O.GscOracle(geo, 256).process(x, ang, method=2))) < 1e-12
    sg = H.make_subband_gsc(mic, 256, angle=[30, 0])
    with _quiet():
        r = sg.process(x.copy())
    o = O.SubbandGscOracle(geo, 256, ang).process(x)
    assert np.max(np.abs(r[0] - o[0])) < 1e-12 and np.max(np.abs(r[2] - o[2])) < 1e-12
    with _quiet():
        td = TDGSC(mic, frameLen=256, angle=[30, 0])
        rt = td.process(x.T.copy())
    ot = O.TdgscOracle(geo, 256, ang).process(x.T)
    assert np.max(np.abs(rt[0] - ot[0])) < 1e-12


def test_fdgsc_live(ref):
    from DistantSpeech.beamformer.MicArray import MicArray
    from DistantSpeech.beamformer.FDGSC import FDGSC
    geo = O.MicGeometry("linear", r=0.05, M=6, n_fft=256)
    x = O.synth_streams(1, geo, 256 * 20, look_deg=(60.0, 0.0), seed0=404)[0].T.astype(np.float64)
    for post in (False, True):
        with _quiet():
            fd = FDGSC(MicArray(arrayType="linear", r=0.05, M=6), frameLen=256, angle=[60, 0])
            r = fd.process(x.copy(), postfilter=post, dc_notch=True)
        o = O.FdgscOracle(geo, 256, np.array([60, 0]) / 180 * np.pi).process(x, postfilter=post)
        assert np.max(np.abs(r[0] - o[0])) < 1e-9


def test_steering_gev_mask_live(ref):
    """8f.3: steering / get_gev_vector / phase_correction / blind_analytic_normalization and the cell-6 accumulation"""
    from DistantSpeech.beamformer.beamformer import (steering, get_gev_vector, phase_correction,
                                                     blind_analytic_normalization, compute_mvdr_weight)
    rng = np.random.default_rng(77)
    K, T, M = 33, 40, 5
    D = rng.standard_normal((K, T, M)) + 1j * rng.standard_normal((K, T, M))
    p = rng.uniform(0.01, 0.99, (K, T))
    Pxx_r = np.zeros((K, M, M), dtype=complex)
    Pvv_r = np.zeros((K, M, M), dtype=complex)
    for n in range(T):                                               # example/mvdr.ipynb cell 6, verbatim arithmetic
        y = D[:, n, :]
        Pxx_r = Pxx_r + np.einsum('ij,il->ijl', y, y.conj()) * p[:, n:n + 1, None]
        Pvv_r = Pvv_r + np.einsum('ij,il->ijl', y, y.conj()) * (1 - p[:, n:n + 1, None])
    Pxx, Pvv = O.masked_covariances(D, p)
    assert np.array_equal(Pxx, Pxx_r) and np.array_equal(Pvv, Pvv_r)
    assert np.array_equal(O.steering_pca(Pxx), steering(Pxx))
    raw = get_gev_vector(Pxx, Pvv)
    assert np.array_equal(O.gev_vector(Pxx, Pvv), raw)
    assert np.array_equal(O.phase_correction(raw), phase_correction(raw))
    assert np.array_equal(O.blind_analytic_normalization(raw, Pvv), blind_analytic_normalization(raw, Pvv))
    assert np.array_equal(O.mask_beamformer_weights(Pxx, Pvv, "mvdr"), compute_mvdr_weight(steering(Pxx), np.linalg.inv(Pvv)))
    # noise matrix that is not positive definite: the reference's fallback branch
    with _quiet():
        assert np.array_equal(O.gev_vector(Pxx[:2], -Pvv[:2]), get_gev_vector(Pxx[:2], -Pvv[:2]))


def test_idoa_live(ref):
    """8f.4: Idoa.estimate / Idoa.process, circular (360 directions) and linear (180) arrays"""
    from DistantSpeech.doa.idoa import Idoa
    from DistantSpeech.beamformer.MicArray import MicArray
    for arr, M, r in (("circular", 4, 0.032), ("linear", 5, 0.04)):
        geo = O.MicGeometry(arr, r=r, M=M, n_fft=256)
        x = np.ascontiguousarray(O.synth_streams(1, geo, 128 * 40, seed0=11 + M)[0].T).astype(np.float64)
        with _quiet():
            mic = MicArray(arrayType=arr, r=r, M=M, n_fft=256)
            rf, rf2 = Idoa(mic), Idoa(mic)
        o, o2 = O.IdoaOracle(geo), O.IdoaOracle(geo)
        assert np.max(np.abs(rf.Psi - o.Psi)) < 1e-13
        yr = np.concatenate([rf.process(x[:128 * 15].copy(), default_direction=30), rf.process(x[128 * 15:].copy(), default_direction=30)])
        yo = np.concatenate([o.process(x[:128 * 15].copy(), default_direction=30), o.process(x[128 * 15:].copy(), default_direction=30)])
        assert np.max(np.abs(yr - yo)) < 1e-12
        X = rf2.transform.stft(x)
        assert np.max(np.abs(rf2.estimate(X, theta=40) - o2.estimate(X, theta=40))) < 1e-12


def test_srp_live(ref):
    """a18: O.srp_angle_spectrum against srp.compute_angle_spectrum (doa/srp.py:17-53) on a fresh input, PHAT on and off"""
    from DistantSpeech.beamformer.MicArray import MicArray
    from DistantSpeech.doa.srp import srp
    import DistantSpeech.doa.srp as srp_mod
    srp_mod.tqdm = lambda it, *a, **k: it
    geo = O.MicGeometry("circular", r=0.04, M=6, n_fft=128)
    x = np.ascontiguousarray(O.synth_streams(1, geo, 64 * 12, look_deg=(75.0, 0.0), seed0=909)[0].T).astype(np.float64)
    with _quiet():
        mic = MicArray(arrayType="circular", r=0.04, M=6, n_fft=128)
    for phat in (True, False):
        with _quiet():
            Pr, pr = srp(mic).compute_angle_spectrum(x.copy(), phat=phat)
        Po, po = O.srp_angle_spectrum(x, geo, phat=phat)
        assert np.max(np.abs(Pr - Po)) <= 1e-12 * np.max(np.abs(Pr)) and np.array_equal(pr, po)


def test_diagnostics_live(ref):
    """a19: the drop-in class's host diagnostics against beamformer.py:435-534 (array gain is [bins, bins] in the reference)"""
    from DistantSpeech.beamformer.MicArray import MicArray
    from DistantSpeech.beamformer.beamformer import beamformer
    from distantspeech_b200.beamformer.MicArray import MicArray as OurMic
    from distantspeech_b200.beamformer.beamformer import beamformer as our_bf
    with _quiet():
        mic = MicArray(arrayType="linear", r=0.03, M=5, n_fft=32)
    rb, ob = beamformer(mic, frame_len=32, hop=16, nfft=32), our_bf(OurMic(arrayType="linear", r=0.03, M=5, n_fft=32), 32, 16, 32)
    W = rb.compute_weights((70, 0), "SD")
    assert np.allclose(ob.compute_weights((70, 0), "SD"), W, rtol=1e-9, atol=1e-12)
    a0 = rb.compute_steering_vector_from_doa((70, 0))
    Gr, Go = rb.compute_array_gain(W, a0, rb.Fvv), ob.compute_array_gain(W, a0, ob.Fvv)
    assert Gr.shape == Go.shape == (17, 17) and np.allclose(Gr, Go, rtol=1e-11)
    for a, b in zip(rb.compute_wng_di(W, look_angle=[70, 0]), ob.compute_wng_di(W, look_angle=[70, 0])):
        assert np.allclose(a, b, rtol=1e-10, atol=1e-10)
    assert np.allclose(rb.compute_beampattern(mic, weights=W.T.copy()), ob.compute_beampattern(ob.MicArray, weights=W.T.copy()),
                       rtol=0, atol=1e-9)


def test_mcspp_more_than_4_mics_live(ref):
    """8f.3: McSpp with 6 microphones = the reference with McCDR(nfft, channels=6) handed in; as shipped it raises IndexError
    (mcspp.py:54); with 8 microphones even the patched reference fails in frames 5..6 (singular unloaded fallback inverse)"""
    from DistantSpeech.transform.transform import Transform
    from DistantSpeech.noise_estimation.mcspp import McSpp
    geo = O.MicGeometry("circular", r=0.04, M=6, n_fft=256)
    x = np.ascontiguousarray(O.synth_streams(1, geo, 128 * 30, seed0=505)[0].T).astype(np.float64)
    D = Transform(n_fft=256, hop_length=128, channel=6).stft(x)
    with _quiet():
        shipped = McSpp(nfft=256, channels=6)
    with pytest.raises(IndexError):
        shipped.estimation(D[:, 0, :])
    r, o = H.make_mcspp(256, 6), O.McSpp(nfft=256, channels=6)
    with np.errstate(all="ignore"):
        for n in range(D.shape[1]):
            assert np.array_equal(r.estimation(D[:, n, :]), o.estimation(D[:, n, :]))
    assert np.array_equal(r.w, o.w)
    geo8 = O.MicGeometry("circular", r=0.04, M=8, n_fft=256)
    x8 = np.ascontiguousarray(O.synth_streams(1, geo8, 128 * 12, seed0=7)[0].T).astype(np.float64)
    D8 = Transform(n_fft=256, hop_length=128, channel=8).stft(x8)
    r8 = H.make_mcspp(256, 8)
    with pytest.raises(np.linalg.LinAlgError), np.errstate(all="ignore"):
        for n in range(D8.shape[1]):
            r8.estimation(D8[:, n, :])


def test_wpe_live(ref):
    """8f.4: the WPE oracle against the reference's own update body (awpe.py:152-187) made executable by ref_harness.make_wpe
    (missing check_input_data supplied, Subband bank replaced by Transform, unassigned `output` swallowed), fresh input"""
    rng = np.random.default_rng(606)
    C, Lf, nb, hop, D = 2, 3, 128, 64, 3
    x = rng.standard_normal((hop * 25, C)) * 0.3
    x[:, 1] += 0.6 * np.roll(x[:, 0], 211)
    r = H.make_wpe(channels=C, filter_len=Lf, num_bands=nb, delay=D, hop_length=hop)
    o = O.WpeOracle(channels=C, filter_len=Lf, num_bands=nb, delay=D, hop_length=hop)
    tx = O.Transform(n_fft=nb, hop_length=hop, channel=C)
    for n in range(25):
        blk = x[n * hop:(n + 1) * hop]
        W, P, var = r.step(blk)
        o.update_spec(tx.stft(blk)[:, 0, :])
        assert np.array_equal(W, o.W) and np.array_equal(P, o.P) and np.array_equal(var, o.var)
    # (the class as shipped is not instantiated here: its Subband bank writes pickles under /home/wangwei, awpe.py:62, subband.py:50)
