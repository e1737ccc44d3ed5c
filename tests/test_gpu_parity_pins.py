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
