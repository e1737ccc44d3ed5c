"""GPU parity, SURVEY 8f.3: data-driven steering, GEV weights and the mask-based beamformers of
example/mvdr.ipynb cells 2 / 6 / 8 -- CUDA path (through the C ABI) against the golden fixture dumped from
the unmodified reference and against the NumPy oracle."""
import numpy as np
import pytest

from conftest import golden, snr_db, assert_wave_parity
from oracle import np_oracle as O

pytestmark = pytest.mark.gpu


def _signs(a, b):
    return np.sign(np.real(np.sum(a * np.conj(b), axis=-1, keepdims=True)))


def test_steering_golden(cuda):
    from distantspeech_b200.beamformer.beamformer import steering
    g = golden("mask_beamformers.npz")
    v = steering(g["Pxx"])
    assert v.shape == g["steer"].shape and np.max(np.abs(v - g["steer"])) < 1e-9
    assert np.max(np.abs(np.linalg.norm(v, axis=-1) - 1)) < 1e-12 and np.max(np.abs(v[:, 0].imag)) < 1e-15
    # cell 2: PCA steering from frame-range averages
    assert np.max(np.abs(steering(g["Ryy"] - g["Rvv"]) - g["steer_pca"])) < 1e-9
    # leading batch axes, torch in -> torch out, lower triangle only
    X = cuda.as_tensor(np.stack([g["Pxx"], g["Pvv"]])).cuda()
    vb = steering(X)
    assert vb.is_cuda and tuple(vb.shape) == (2, 257, 6)
    assert np.max(np.abs(vb[0].cpu().numpy() - g["steer"])) < 1e-9
    low = g["Pxx"].copy()
    low[:, np.triu_indices(6, 1)[0], np.triu_indices(6, 1)[1]] = 7.0
    assert np.max(np.abs(steering(low) - g["steer"])) < 1e-9
    with pytest.raises(Exception):
        steering(np.zeros((4, 9, 9), complex))                       # > 8 sensors: not compiled


def test_gev_phase_correction_ban_golden(cuda):
    from distantspeech_b200.beamformer.beamformer import get_gev_vector, phase_correction, blind_analytic_normalization
    g = golden("mask_beamformers.npz")
    w = get_gev_vector(g["Pxx"], g["Pvv"])
    s = _signs(w, g["w_gev_raw"])                                    # LAPACK's free sign per bin (include/ds_b200.h)
    assert np.max(np.abs(w - s * g["w_gev_raw"])) < 1e-8 * np.max(np.abs(g["w_gev_raw"]))
    quad = np.einsum("ka,kab,kb->k", w.conj(), g["Pvv"], w)
    assert np.max(np.abs(quad - 1)) < 1e-10                          # scipy.linalg.eigh normalisation
    # given the reference's own raw vectors the two follow-up steps are deterministic
    pc = phase_correction(g["w_gev_raw"])
    assert np.max(np.abs(pc - g["w_gev_pc"])) < 1e-12 * np.max(np.abs(g["w_gev_pc"]))
    ban = blind_analytic_normalization(g["w_gev_pc"], g["Pvv"])
    assert np.allclose(ban, g["w_gev"], rtol=1e-11, atol=1e-14)
    # the whole chain from our vectors equals the reference's up to ONE global sign (bin 0's)
    full = blind_analytic_normalization(phase_correction(w), g["Pvv"])
    s0 = np.sign(np.real(np.vdot(g["w_gev"][0], full[0])))
    assert np.max(np.abs(full - s0 * g["w_gev"])) < 1e-7 * np.max(np.abs(g["w_gev"]))
    # not positive definite -> the reference's LinAlgError fallback (beamformer.py:94-96)
    bad = -g["Pvv"][:3]
    fb = get_gev_vector(g["Pxx"][:3], bad)
    ref = np.stack([np.ones(6) / np.trace(bad[k]) * 6 for k in range(3)])
    assert np.allclose(fb, ref, rtol=1e-12)


def test_masked_covariances_golden(cuda):
    from distantspeech_b200.beamformer.beamformer import masked_covariances
    g = golden("mask_beamformers.npz")
    D = O.Transform(n_fft=512, hop_length=256, channel=6).stft(g["x"].astype(np.float64))
    Pxx, Pvv = masked_covariances(D, g["p"])
    sc = np.max(np.abs(g["Pxx"]))
    assert np.max(np.abs(Pxx - g["Pxx"])) < 1e-12 * sc and np.max(np.abs(Pvv - g["Pvv"])) < 1e-12 * sc
    # cell 2: plain averages over frame ranges
    Rvv, none = masked_covariances(D, frames=(0, 10), scale=1 / 10)
    assert none is None and np.max(np.abs(Rvv - g["Rvv"])) < 1e-12 * np.max(np.abs(g["Rvv"]))
    Ryy, _ = masked_covariances(D, frames=(30, 80), scale=1 / 50)
    assert np.max(np.abs(Ryy - g["Ryy"])) < 1e-12 * np.max(np.abs(g["Ryy"]))
    with pytest.raises(Exception):
        masked_covariances(D, frames=(5, 1000))


@pytest.mark.parametrize("prec", ["fp32", "fp64"])
def test_mask_mvdr_and_gev_golden(cuda, prec):
    from distantspeech_b200.pipelines import MaskBeamformer
    g = golden("mask_beamformers.npz")
    bf = MaskBeamformer(6, n_fft=512, hop=256, method="mvdr", fft_precision=prec)
    y = bf.process(g["x"], p=g["p"])
    err, s = assert_wave_parity(g["y_mvdr"], y, "mask-MVDR %s" % prec)
    print("mask-MVDR %s: max-abs %.2e SNR %.1f dB" % (prec, err, s))
    w = bf.w[0].cpu().numpy()
    assert np.max(np.abs(w - g["w_mvdr"])) < 1e-4 * np.max(np.abs(g["w_mvdr"]))
    bg = MaskBeamformer(6, n_fft=512, hop=256, method="gev", fft_precision=prec)
    y = bg.process(g["x"], p=g["p"])
    sgn = np.sign(np.sum(y * g["y_gev"]))                             # the one sign LAPACK leaves open
    err, s = assert_wave_parity(g["y_gev"], sgn * y, "mask-GEV %s" % prec)
    print("mask-GEV %s: max-abs %.2e SNR %.1f dB" % (prec, err, s))


def test_mask_beamformer_own_mask_batch(cuda):
    """mask estimated on the device (McSppBase per frame, cell 4's loop), 3 streams at once, 8 microphones"""
    from distantspeech_b200.pipelines import MaskBeamformer
    geo = O.MicGeometry("circular", r=0.05, M=8, n_fft=512)
    xs = O.synth_streams(3, geo, 256 * 100, seed0=0x3A5)
    x_nm = np.ascontiguousarray(xs.transpose(0, 2, 1))
    bf = MaskBeamformer(8, method="mvdr")
    y = bf.process(x_nm)
    assert y.shape == (3, 256 * 100)
    for s in range(3):
        taps = {}
        ref = O.mask_beamform(x_nm[s].astype(np.float64), method="mvdr", taps=taps)
        err, snr = assert_wave_parity(ref, y[s], "mask-MVDR own mask, stream %d" % s)
        dp = np.max(np.abs(bf.p[s].cpu().numpy().T - taps["p"]))
        print("stream %d: max-abs %.2e SNR %.1f dB, max |dp| %.2e" % (s, err, snr, dp))
    y1 = MaskBeamformer(8, method="mvdr").process(x_nm[1])
    assert np.array_equal(y1, y[1])                                   # batching does not change a stream
    with pytest.raises(ValueError):
        bf.process(x_nm[0][:1000])                                    # not a multiple of hop
    with pytest.raises(ValueError):
        MaskBeamformer(8, method="lcmv")
