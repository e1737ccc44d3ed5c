"""CPU: libds_b200.so loads and exports every symbol include/ds_b200.h declares."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    txt = open(os.path.join(ROOT, "include", "ds_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ds_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_entry_points():
    names = _declared()
    for must in ("ds_stft_run", "ds_istft_run", "ds_fixedbf_run", "ds_mcra_run", "ds_mcspp_run", "ds_chain_run"):
        assert must in names


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from distantspeech_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, missing
    lib.ds_version.restype = ctypes.c_int
    assert lib.ds_version() >= 100


def test_host_side_helpers_without_gpu():
    from distantspeech_b200 import _lib
    lib = _lib.lib()
    p = _lib.StftParams(512, 256, 1, 1, 16000, _lib.DS_STFT_STREAMING, 0, 0)
    assert lib.ds_stft_num_frames(ctypes.byref(p)) == 62
    p.mode = _lib.DS_STFT_CENTER
    assert lib.ds_stft_num_frames(ctypes.byref(p)) == 63
    f, e = ctypes.c_int32(0), ctypes.c_int32(1)
    lib.ds_mcra_advance(15, 30, ctypes.byref(f), ctypes.byref(e))
    assert (f.value, e.value) == (30, 1)          # window reset at frames 14 and 29 (mcra.py:52-56)
    mp = _lib.McsppParams()
    lib.ds_mcspp_default_params(ctypes.byref(mp), 512, 4, 8, 10)
    assert mp.alpha == 0.92 and mp.mcra_L == 15 and mp.Gmin == 0.0631
    assert lib.ds_mcspp_state_bytes(ctypes.byref(mp)) == 4 * (2 * 64 + 5) * 257 * 8


def test_no_cpu_fallback():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    from distantspeech_b200 import _lib
    from distantspeech_b200.transform.transform import Transform
    with pytest.raises(_lib.DsError):
        Transform(n_fft=256, hop_length=128).stft(np.zeros(1024))
