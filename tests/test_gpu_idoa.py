"""GPU parity, SURVEY 8f.4: Idoa.estimate / Idoa.process (doa/idoa.py) -- CUDA path through the C ABI against the golden
fixture dumped from the unmodified reference and against the NumPy oracle."""
import numpy as np
import pytest

from conftest import golden, assert_wave_parity
from oracle import np_oracle as O

pytestmark = pytest.mark.gpu

CASES = {"c4": ("circular", 4, 0.032, 256), "l6": ("linear", 6, 0.05, 512)}


@pytest.mark.parametrize("tag", ["c4", "l6"])
def test_idoa_golden(cuda, tag):
    from distantspeech_b200.beamformer.MicArray import MicArray
    from distantspeech_b200.doa.idoa import Idoa
    g = golden("idoa.npz")
    arr, M, r, n_fft = CASES[tag]
    mic = MicArray(arrayType=arr, r=r, M=M, n_fft=n_fft)
    x = g[tag + "_x"]
    n1 = int(g[tag + "_n1"])
    sel = g[tag + "_sel"]
    a = Idoa(mic)
    assert a.n_theta == (360 if arr == "circular" else 180)
    assert np.allclose(a.Psi[:, :, sel], g[tag + "_Psi_sel"], rtol=0, atol=1e-12)
    # process: two chunks, the state (RTF estimate, statistics, STFT history, overlap tail) carries over like the reference's
    y = np.concatenate([a.process(x[:n1], default_direction=30), a.process(x[n1:], default_direction=30)])
    err, s = assert_wave_parity(g[tag + "_y"], y, "Idoa.process %s" % tag)
    print("Idoa.process %s: max-abs %.2e SNR %.1f dB" % (tag, err, s))
    # estimate on the reference's own spectrum: every direction of the grid, compared on the stored columns
    X = O.Transform(channel=M, n_fft=n_fft, hop_length=n_fft // 2).stft(x.astype(np.float64))
    b = Idoa(mic)
    p = b.estimate(X)
    assert p.shape == (n_fft // 2 + 1, X.shape[1], a.n_theta)
    dp = np.max(np.abs(p[:, :, sel] - g[tag + "_p_sel"]))
    assert dp < 1e-8, dp
    assert np.max(np.abs(b.mu_Delta[:, sel] - g[tag + "_mu_Delta_last"])) < 1e-9
    assert np.max(np.abs(b.var_Delta_h0[:, sel] - g[tag + "_var_last"])) < 1e-9
    c = Idoa(mic)
    p1 = c.estimate(X, theta=40)
    assert np.max(np.abs(p1[:, :, [40, 41]] - g[tag + "_p_theta40"])) < 1e-8
    print("Idoa.estimate %s: max |dp| %.2e" % (tag, dp))


def test_idoa_batch_and_errors(cuda):
    from distantspeech_b200.beamformer.MicArray import MicArray
    from distantspeech_b200.doa.idoa import Idoa
    geo = O.MicGeometry("circular", r=0.05, M=8, n_fft=512)
    xs = np.ascontiguousarray(O.synth_streams(3, geo, 256 * 40, seed0=77).transpose(0, 2, 1))     # [3, N, 8]
    mic = MicArray(arrayType="circular", r=0.05, M=8, n_fft=512)
    y = Idoa(mic).process(xs, theta=120)
    assert y.shape == (3, 256 * 40)
    for s in range(3):
        ref = O.IdoaOracle(geo).process(xs[s].astype(np.float64), theta=120)
        assert_wave_parity(ref, y[s], "Idoa batch stream %d" % s)
    assert np.array_equal(Idoa(mic).process(xs[2], theta=120), y[2])
    with pytest.raises(NotImplementedError):
        Idoa(mic).process(xs[0], pre_emphsis=True)
    with pytest.raises(Exception):                                   # n_fft 128: fewer than the 128 bins the reference indexes
        Idoa(MicArray(arrayType="circular", r=0.05, M=4, n_fft=128)).process(np.zeros((64 * 10, 4), np.float32))
