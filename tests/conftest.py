import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "needs_reference: imports the reference from /root/reference (build container only)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def snr_db(ref, x):
    ref = np.asarray(ref, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    num = np.sum(np.abs(ref) ** 2)
    den = np.sum(np.abs(ref - x) ** 2)
    return 10 * np.log10(num / max(den, 1e-300))


# Parity contract of BASELINE.json:north_star for waveforms
WAVE_MAXABS = 1e-4
WAVE_SNR_DB = 60.0


def assert_wave_parity(ref, out, what=""):
    ref = np.asarray(ref, dtype=np.float64)
    out = np.asarray(out, dtype=np.float64)
    assert ref.shape == out.shape, (what, ref.shape, out.shape)
    assert np.all(np.isfinite(out)), what
    err = np.max(np.abs(ref - out))
    s = snr_db(ref, out)
    assert err <= WAVE_MAXABS, "%s: max-abs error %.3e > %.1e (SNR %.1f dB)" % (what, err, WAVE_MAXABS, s)
    assert s >= WAVE_SNR_DB, "%s: SNR %.1f dB < %.0f dB (max-abs %.3e)" % (what, s, WAVE_SNR_DB, err)
    return err, s


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from distantspeech_b200 import _lib
    _lib.ensure_init()          # raises loudly if libds_b200.so is missing
    return torch
