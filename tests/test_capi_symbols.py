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


def test_argument_validation_precedes_any_device_work():
    """Every *_run rejects null pointers / unsupported shapes with DS_EINVAL / DS_EUNSUPPORTED and a message before it
    touches the device -- so this runs without a GPU."""
    from distantspeech_b200 import _lib
    lib = _lib.lib()
    lib.ds_last_error.restype = ctypes.c_char_p
    null = ctypes.c_void_p(0)
    dummy = (ctypes.c_double * 4)()
    one = ctypes.cast(dummy, ctypes.c_void_p)
    EINVAL, EUNSUP = -1, -2

    sp = _lib.StftParams(512, 256, 1, 1, 4096, _lib.DS_STFT_STREAMING, 0, 0)
    assert lib.ds_stft_run(ctypes.byref(sp), null, null, null, null, null) == EINVAL and b"null" in lib.ds_last_error()
    sp.hop = 1024                                                            # hop > n_fft
    assert lib.ds_stft_run(ctypes.byref(sp), one, one, one, one, null) == EINVAL and b"hop" in lib.ds_last_error()
    sp.hop, sp.n_fft = 250, 500                                              # not a power of two
    assert lib.ds_stft_run(ctypes.byref(sp), one, one, one, one, null) == EUNSUP and b"n_fft" in lib.ds_last_error()
    ip = _lib.IstftParams(512, 256, 1, 1, 4, 1, 0, 0, 1.0)                   # DS_STFT_CENTER is not a synthesis mode
    assert lib.ds_istft_run(ctypes.byref(ip), one, one, one, one, null) == EINVAL

    mp = _lib.McsppParams()
    lib.ds_mcspp_default_params(ctypes.byref(mp), 512, 1, 9, 4)              # 9 microphones: outside 2..8
    assert lib.ds_mcspp_run(ctypes.byref(mp), one, null, one, 0, null, 0, None, null) == EUNSUP
    cp = _lib.McsppCdrParams()
    lib.ds_mcspp_cdr_default_params(ctypes.byref(cp), 512, 1, 3, 4)          # McSpp needs 4..8 channels
    assert lib.ds_mcspp_cdr_run(ctypes.byref(cp), one, one, one, one, 0, null, None, null) == EUNSUP
    assert b"n_mics must be 4..8" in lib.ds_last_error()
    assert lib.ds_wpe_run(1, 257, 4, 8, 3, 2, 0.998, 0.98, one, one, 0, one, null) == EINVAL      # 8 channels x 3 taps > 16
    assert lib.ds_wpe_state_bytes(1, 257, 0, 2, 2) == 0
    gp = _lib.GscParams()
    lib.ds_gsc_default_params(ctypes.byref(gp), 256, 1, 4, 4)
    assert lib.ds_gsc_run(ctypes.byref(gp), one, null, one, 0, one, None, null) == EINVAL     # output without propagation vectors
    fp = _lib.FdafParams(128, 1, 3, 1024, 30, 1, 1, 0, 0.01, 0.9)            # frame_len 128 is not compiled
    assert lib.ds_fdaf_run(ctypes.byref(fp), one, one, one, null, one, null) == EUNSUP
    np_ = _lib.SubbandNlmsParams(257, 1, 1, 4, 5, 4, 0, 0, 0.1, 0.9, 1e-4)   # 4 taps x 5 channels: not compiled
    assert lib.ds_subband_nlms_run(ctypes.byref(np_), one, one, one, null, one, null) == EUNSUP
    assert lib.ds_srp_run(10, 4, 5, 513, 48000.0, 1024, one, one, null, one, 1, null) == EUNSUP   # tensor path: 4, 8, 16 mics
    assert lib.ds_srp_workspace_bytes(937, 16, 513, 1) == 513 * 8 * 256 * 32 * 4 and lib.ds_srp_workspace_bytes(937, 16, 513, 0) == 0
    # 8f.3 / 8f.4 entry points
    assert lib.ds_steering_run(4, 6, null, one, null) == EINVAL and b"null" in lib.ds_last_error()
    assert lib.ds_steering_run(4, 9, one, one, null) == EUNSUP and b"n_mics" in lib.ds_last_error()    # > 8 sensors
    assert lib.ds_gev_run(0, 6, one, one, one, null) == EINVAL
    assert lib.ds_mvdr_from_cov_run(4, 12, one, one, one, null) == EUNSUP
    assert lib.ds_ban_run(4, 6, one, null, 0.0, one, null) == EINVAL
    assert lib.ds_phase_correction_run(1, 0, 6, one, null) == EINVAL
    assert lib.ds_masked_cov_run(1, 10, 6, 257, 5, 20, one, 0, one, 1.0, one, one, null) == EINVAL and b"frame range" in lib.ds_last_error()
    assert lib.ds_masked_cov_run(1, 10, 6, 257, 0, 10, one, 0, null, 1.0, one, one, null) == EINVAL    # Phi_vv without a mask
    assert lib.ds_masked_cov_run(1, 10, 9, 257, 0, 10, one, 0, one, 1.0, one, one, null) == EUNSUP
    assert lib.ds_apply_stream_weights_run(1, 0, 6, 257, one, 0, one, one, null) == EINVAL
    assert lib.ds_idoa_rtf_run(1, 4, 9, 257, 0.02, one, 0, one, one, null) == EUNSUP
    assert lib.ds_idoa_spp_run(1, 4, 4, 65, 1, one, -1, one, one, one, null, null, 0, null, null) == EINVAL   # fewer than 128 bins
    assert b"72..127" in lib.ds_last_error()
    assert lib.ds_idoa_spp_run(1, 4, 4, 129, 2, one, -1, one, one, one, null, one, 0, one, null) == EINVAL    # gain output needs one direction
    assert lib.ds_idoa_rtf_state_bytes(2, 4, 129) == 2 * 7 * 129 * 8 and lib.ds_idoa_spp_state_bytes(2, 3, 129) == 2 * 3 * 4 * 129 * 8


def test_bench_static_inputs():
    """bench.py parses on this interpreter and the ncu-derived figures it quotes are where it looks for them"""
    import ast
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ast.parse(open(os.path.join(root, "bench.py")).read())
    tj = json.load(open(os.path.join(root, "profiles", "traffic.json")))["mcspp_fast_kernel"]
    assert tj["dram_bytes_per_stream_10s"] > 5e6 and 1500 < tj["fp64_flop_per_bin_frame"] < 2500 and 0 < tj["ncu_pipe_fp64_pct"] < 100
    base = json.load(open(os.path.join(root, "BASELINE.json")))
    assert "audio-s/s" in base["metric"]
