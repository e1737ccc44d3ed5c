"""ctypes binding of libds_b200.so (see include/ds_b200.h).

PyTorch is used for device memory and streams only; every computation on the
hot path happens in the hand-written sm_100a kernels of the shared library.
There is no CPU fallback: if the library or a CUDA device is missing the
product raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DS_B200_LIB", os.path.join(_HERE, "libds_b200.so"))   # override: A/B testing of builds

DS_OK = 0
DS_STFT_STREAMING, DS_STFT_CENTER, DS_STFT_PLAIN = 0, 1, 2


class DsError(RuntimeError):
    pass


class StftParams(C.Structure):
    _fields_ = [("n_fft", C.c_int32), ("hop", C.c_int32), ("n_streams", C.c_int32), ("n_ch", C.c_int32),
                ("n_samples", C.c_int32), ("mode", C.c_int32), ("fft_fp64", C.c_int32), ("out_c128", C.c_int32)]


class IstftParams(C.Structure):
    _fields_ = [("n_fft", C.c_int32), ("hop", C.c_int32), ("n_streams", C.c_int32), ("n_ch", C.c_int32),
                ("n_frames", C.c_int32), ("mode", C.c_int32), ("fft_fp64", C.c_int32), ("in_c128", C.c_int32),
                ("scale", C.c_double)]


class FixedBfParams(C.Structure):
    _fields_ = [("n_fft", C.c_int32), ("hop", C.c_int32), ("n_streams", C.c_int32), ("n_mics", C.c_int32),
                ("n_samples", C.c_int32), ("n_beams", C.c_int32), ("reserved0", C.c_int32), ("reserved1", C.c_int32),
                ("scale", C.c_double)]


class McraParams(C.Structure):
    _fields_ = [("n_bins", C.c_int32), ("n_streams", C.c_int32), ("n_frames", C.c_int32), ("L", C.c_int32),
                ("frm_cnt", C.c_int32), ("ell", C.c_int32), ("reserved0", C.c_int32), ("reserved1", C.c_int32),
                ("alpha_d", C.c_double), ("alpha_s", C.c_double), ("delta_s", C.c_double), ("alpha_p", C.c_double),
                ("p_min", C.c_double), ("p_max", C.c_double)]


class McsppParams(C.Structure):
    _fields_ = [("n_fft", C.c_int32), ("n_streams", C.c_int32), ("n_mics", C.c_int32), ("n_frames", C.c_int32),
                ("frm_cnt", C.c_int32), ("ell", C.c_int32), ("mcra_L", C.c_int32), ("full_state", C.c_int32),
                ("alpha", C.c_double), ("alpha_d", C.c_double), ("diag_eps", C.c_double),
                ("q_min", C.c_double), ("q_max", C.c_double), ("p_min", C.c_double), ("p_max", C.c_double),
                ("snr_min", C.c_double), ("snr_max", C.c_double), ("Gmin", C.c_double),
                ("mcra_alpha_d", C.c_double), ("mcra_alpha_s", C.c_double), ("mcra_delta_s", C.c_double),
                ("mcra_alpha_p", C.c_double), ("mcra_p_min", C.c_double), ("mcra_p_max", C.c_double)]


class McsppTaps(C.Structure):
    _fields_ = [("p", C.c_void_p), ("xi", C.c_void_p), ("gamma", C.c_void_p), ("q", C.c_void_p), ("G", C.c_void_p),
                ("w_mvdr", C.c_void_p), ("w_pmwf", C.c_void_p), ("Phi_vv_inv_last", C.c_void_p)]


class McsppCdrParams(C.Structure):
    _fields_ = [("n_fft", C.c_int32), ("n_streams", C.c_int32), ("n_mics", C.c_int32), ("n_frames", C.c_int32),
                ("frm_cnt", C.c_int32), ("ell", C.c_int32), ("mcra_L", C.c_int32), ("cdr_only", C.c_int32),
                ("band_lo_bin", C.c_int32), ("band_hi_bin", C.c_int32), ("init_frames", C.c_int32),
                ("fallback_loaded_frames", C.c_int32),
                ("alpha", C.c_double), ("alpha_d", C.c_double), ("alpha_cdr", C.c_double),
                ("load_min", C.c_double), ("load_max", C.c_double), ("snr_min", C.c_double), ("snr_max", C.c_double),
                ("pmwf_beta", C.c_double), ("q_init", C.c_double),
                ("mcra_alpha_d", C.c_double), ("mcra_alpha_s", C.c_double), ("mcra_delta_s", C.c_double),
                ("mcra_alpha_p", C.c_double), ("mcra_p_min", C.c_double), ("mcra_p_max", C.c_double)]


class McsppCdrTaps(C.Structure):
    _fields_ = [("p", C.c_void_p), ("xi", C.c_void_p), ("gamma", C.c_void_p), ("q", C.c_void_p), ("cdr", C.c_void_p),
                ("w", C.c_void_p)]


class GscParams(C.Structure):
    _fields_ = [("n_fft", C.c_int32), ("n_streams", C.c_int32), ("n_mics", C.c_int32), ("n_frames", C.c_int32),
                ("frm_cnt", C.c_int32), ("method", C.c_int32), ("init_frames", C.c_int32), ("reserved", C.c_int32),
                ("alpha", C.c_double), ("alpha_d", C.c_double), ("diag_eps", C.c_double), ("psi_0", C.c_double),
                ("q_min", C.c_double), ("q_max", C.c_double), ("p_min", C.c_double), ("p_max", C.c_double),
                ("snr_min", C.c_double), ("snr_max", C.c_double), ("Gmin", C.c_double), ("mu", C.c_double)]


class GscTaps(C.Structure):
    _fields_ = [("p", C.c_void_p), ("G", C.c_void_p), ("xi", C.c_void_p), ("gamma", C.c_void_p), ("q", C.c_void_p)]


class SubbandNlmsParams(C.Structure):
    _fields_ = [("n_bins", C.c_int32), ("n_streams", C.c_int32), ("n_filters", C.c_int32), ("n_frames", C.c_int32),
                ("n_ch", C.c_int32), ("filter_len", C.c_int32), ("one_minus_p", C.c_int32), ("plain_lms", C.c_int32),
                ("mu", C.c_double), ("alpha", C.c_double), ("eps", C.c_double)]


class FdafParams(C.Structure):
    _fields_ = [("frame_len", C.c_int32), ("n_streams", C.c_int32), ("n_ch", C.c_int32), ("n_samples", C.c_int32),
                ("fir_truncate", C.c_int32), ("non_causal", C.c_int32), ("one_minus_p", C.c_int32), ("reserved", C.c_int32),
                ("mu", C.c_double), ("alpha", C.c_double)]


class AmvdrParams(C.Structure):
    _fields_ = [("n_fft", C.c_int32), ("n_streams", C.c_int32), ("n_mics", C.c_int32), ("n_frames", C.c_int32),
                ("frm_cnt", C.c_int32), ("ell", C.c_int32), ("mcra_L", C.c_int32), ("method", C.c_int32),
                ("alpha_y", C.c_double), ("alpha_v", C.c_double), ("diag", C.c_double), ("vad_thr", C.c_double),
                ("mcra_alpha_d", C.c_double), ("mcra_alpha_s", C.c_double), ("mcra_delta_s", C.c_double),
                ("mcra_alpha_p", C.c_double), ("mcra_p_min", C.c_double), ("mcra_p_max", C.c_double)]


class FdgscParams(C.Structure):
    _fields_ = [("frame_len", C.c_int32), ("n_streams", C.c_int32), ("n_mics", C.c_int32), ("n_samples", C.c_int32),
                ("filter_len", C.c_int32), ("frm_cnt", C.c_int32), ("ell", C.c_int32), ("mcra_L", C.c_int32),
                ("fp64", C.c_int32), ("dc_notch", C.c_int32), ("reserved", C.c_int32), ("reserved2", C.c_int32),
                ("mu_bm", C.c_double), ("mu_aic", C.c_double), ("alpha", C.c_double), ("notch_radius", C.c_double),
                ("maxnorm", C.c_double), ("delta", C.c_double),
                ("mcra_alpha_d", C.c_double), ("mcra_alpha_s", C.c_double), ("mcra_delta_s", C.c_double),
                ("mcra_alpha_p", C.c_double), ("mcra_p_min", C.c_double), ("mcra_p_max", C.c_double)]


class OmlsaMultiParams(C.Structure):
    _fields_ = [("n_bins", C.c_int32), ("n_streams", C.c_int32), ("n_frames", C.c_int32), ("n_mics", C.c_int32),
                ("first_frame", C.c_int32), ("frm_cnt", C.c_int32), ("ell", C.c_int32), ("mcra_L", C.c_int32),
                ("cal_weights", C.c_int32), ("u_const", C.c_int32),
                ("alpha_d", C.c_double), ("alpha_s", C.c_double), ("alpha_xi", C.c_double), ("beta", C.c_double),
                ("Gmin", C.c_double), ("q_min", C.c_double), ("q_max", C.c_double),
                ("mcra_alpha_d", C.c_double), ("mcra_alpha_s", C.c_double), ("mcra_delta_s", C.c_double),
                ("mcra_alpha_p", C.c_double), ("mcra_p_min", C.c_double), ("mcra_p_max", C.c_double)]


class ChainParams(C.Structure):
    _fields_ = [("est", McsppParams), ("hop", C.c_int32), ("n_samples", C.c_int32), ("fft_fp64", C.c_int32),
                ("apply_gain", C.c_int32), ("scale", C.c_double)]


_lib = None


def _declare(lib):
    vp, i32, dbl = C.c_void_p, C.c_int, C.c_double
    lib.ds_version.restype = i32
    lib.ds_last_error.restype = C.c_char_p
    lib.ds_init.restype = i32
    lib.ds_device_info.argtypes = [C.POINTER(i32)] * 3
    lib.ds_stft_num_frames.argtypes = [C.POINTER(StftParams)]
    lib.ds_stft_run.argtypes = [C.POINTER(StftParams), vp, vp, vp, vp, vp]
    lib.ds_istft_run.argtypes = [C.POINTER(IstftParams), vp, vp, vp, vp, vp]
    lib.ds_fixedbf_state_bytes.argtypes = [C.POINTER(FixedBfParams)]
    lib.ds_fixedbf_state_bytes.restype = C.c_size_t
    lib.ds_fixedbf_run.argtypes = [C.POINTER(FixedBfParams), vp, vp, vp, vp, vp, vp]
    lib.ds_mvdr_weight_run.argtypes = [i32, i32, vp, vp, vp, vp]
    lib.ds_pmwf_weight_run.argtypes = [i32, i32, vp, vp, vp, dbl, vp, vp]
    lib.ds_apply_weights_run.argtypes = [i32, i32, i32, i32, vp, i32, vp, vp, vp]
    lib.ds_omlsa_gain_run.argtypes = [i32, i32, vp, vp, dbl, vp, vp, vp]
    lib.ds_mcra_state_bytes.argtypes = [C.POINTER(McraParams)]
    lib.ds_mcra_state_bytes.restype = C.c_size_t
    lib.ds_mcra_run.argtypes = [C.POINTER(McraParams), vp, vp, vp, vp, vp]
    lib.ds_mcra_advance.argtypes = [C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    lib.ds_mcra_advance.restype = None
    lib.ds_mcspp_default_params.argtypes = [C.POINTER(McsppParams), i32, i32, i32, i32]
    lib.ds_mcspp_default_params.restype = None
    lib.ds_mcspp_state_bytes.argtypes = [C.POINTER(McsppParams)]
    lib.ds_mcspp_state_bytes.restype = C.c_size_t
    lib.ds_mcspp_run.argtypes = [C.POINTER(McsppParams), vp, vp, vp, i32, vp, i32, C.POINTER(McsppTaps), vp]
    lib.ds_mcspp_export.argtypes = [C.POINTER(McsppParams), vp, i32, vp, vp]
    lib.ds_amvdr_default_params.argtypes = [C.POINTER(AmvdrParams), i32, i32, i32, i32]
    lib.ds_amvdr_default_params.restype = None
    lib.ds_amvdr_state_bytes.argtypes = [C.POINTER(AmvdrParams)]
    lib.ds_amvdr_state_bytes.restype = C.c_size_t
    lib.ds_amvdr_run.argtypes = [C.POINTER(AmvdrParams), vp, vp, vp, vp, vp, vp, vp]
    lib.ds_amvdr_export.argtypes = [C.POINTER(AmvdrParams), vp, i32, vp, vp]
    lib.ds_fdgsc_default_params.argtypes = [C.POINTER(FdgscParams), i32, i32, i32, i32]
    lib.ds_fdgsc_default_params.restype = None
    lib.ds_fdgsc_state_bytes.argtypes = [C.POINTER(FdgscParams)]
    lib.ds_fdgsc_state_bytes.restype = C.c_size_t
    lib.ds_fdgsc_run.argtypes = [C.POINTER(FdgscParams), vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.ds_fdgsc_workspace_bytes.argtypes = [C.POINTER(FdgscParams)]
    lib.ds_fdgsc_workspace_bytes.restype = C.c_size_t
    lib.ds_fdgsc_run_ws.argtypes = [C.POINTER(FdgscParams), vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.ds_fir_run.argtypes = [i32, i32, i32, i32, vp, vp, vp, vp, vp, vp]
    lib.ds_omlsa_multi_default_params.argtypes = [C.POINTER(OmlsaMultiParams), i32, i32, i32, i32]
    lib.ds_omlsa_multi_default_params.restype = None
    lib.ds_omlsa_multi_state_bytes.argtypes = [C.POINTER(OmlsaMultiParams)]
    lib.ds_omlsa_multi_state_bytes.restype = C.c_size_t
    lib.ds_omlsa_multi_run.argtypes = [C.POINTER(OmlsaMultiParams), vp, vp, vp, vp, vp, vp, vp]
    lib.ds_zelinski_state_bytes.argtypes = [i32, i32, i32]
    lib.ds_zelinski_state_bytes.restype = C.c_size_t
    lib.ds_zelinski_run.argtypes = [i32, i32, i32, i32, dbl, dbl, vp, vp, vp, vp, vp]
    lib.ds_pcm16_to_float_run.argtypes = [C.c_size_t, vp, vp, vp]
    lib.ds_float_to_pcm16_run.argtypes = [C.c_size_t, vp, vp, vp]
    lib.ds_phat_run.argtypes = [i32, i32, i32, i32, vp, vp, vp]
    lib.ds_srp_run.argtypes = [i32, i32, i32, i32, dbl, i32, vp, vp, vp, vp, i32, vp]
    lib.ds_srp_workspace_bytes.argtypes = [i32, i32, i32, i32]
    lib.ds_srp_workspace_bytes.restype = C.c_size_t
    lib.ds_fdgsc_notch_run.argtypes = [C.POINTER(FdgscParams), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.ds_fdgsc_notch_run.restype = C.c_int
    lib.ds_gsc_default_params.argtypes = [C.POINTER(GscParams), C.c_int, C.c_int, C.c_int, C.c_int]
    lib.ds_gsc_default_params.restype = None
    lib.ds_gsc_state_bytes.argtypes = [C.POINTER(GscParams)]
    lib.ds_gsc_state_bytes.restype = C.c_size_t
    lib.ds_gsc_run.argtypes = [C.POINTER(GscParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                               C.POINTER(GscTaps), C.c_void_p]
    lib.ds_gsc_run.restype = C.c_int
    lib.ds_subband_nlms_state_bytes.argtypes = [C.POINTER(SubbandNlmsParams)]
    lib.ds_subband_nlms_state_bytes.restype = C.c_size_t
    lib.ds_subband_nlms_run.argtypes = [C.POINTER(SubbandNlmsParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p]
    lib.ds_subband_nlms_run.restype = C.c_int
    lib.ds_dcnotch_run.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ds_dcnotch_run.restype = C.c_int
    lib.ds_channel_mean_run.argtypes = [C.c_int, C.c_int, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ds_channel_mean_run.restype = C.c_int
    lib.ds_subband_rls_state_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
    lib.ds_subband_rls_state_bytes.restype = C.c_size_t
    lib.ds_subband_rls_run.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ds_subband_rls_run.restype = C.c_int
    lib.ds_fdaf_state_bytes.argtypes = [C.c_int, C.c_int]
    lib.ds_fdaf_state_bytes.restype = C.c_size_t
    lib.ds_fdaf_run.argtypes = [C.POINTER(FdafParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ds_fdaf_run.restype = C.c_int
    lib.ds_adjacent_diff_run.argtypes = [C.c_int, C.c_int, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ds_adjacent_diff_run.restype = C.c_int
    lib.ds_power_run.argtypes = [C.c_longlong, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.ds_power_run.restype = C.c_int
    lib.ds_spectral_gain_run.argtypes = [C.c_longlong, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.ds_spectral_gain_run.restype = C.c_int
    lib.ds_mcspp_cdr_default_params.argtypes = [C.POINTER(McsppCdrParams), C.c_int, C.c_int, C.c_int, C.c_int]
    lib.ds_mcspp_cdr_default_params.restype = None
    lib.ds_mcspp_cdr_state_bytes.argtypes = [C.POINTER(McsppCdrParams)]
    lib.ds_mcspp_cdr_state_bytes.restype = C.c_size_t
    lib.ds_mcspp_cdr_workspace_bytes.argtypes = [C.POINTER(McsppCdrParams)]
    lib.ds_mcspp_cdr_workspace_bytes.restype = C.c_size_t
    lib.ds_mcspp_cdr_run.argtypes = [C.POINTER(McsppCdrParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                     C.c_void_p, C.POINTER(McsppCdrTaps), C.c_void_p]
    lib.ds_mcspp_cdr_run.restype = C.c_int
    lib.ds_mcspp_cdr_export.argtypes = [C.POINTER(McsppCdrParams), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.ds_mcspp_cdr_export.restype = C.c_int
    lib.ds_masked_cov_run.argtypes = [i32, i32, i32, i32, i32, i32, vp, i32, vp, dbl, vp, vp, vp]
    lib.ds_steering_run.argtypes = [C.c_longlong, i32, vp, vp, vp]
    lib.ds_gev_run.argtypes = [C.c_longlong, i32, vp, vp, vp, vp]
    lib.ds_phase_correction_run.argtypes = [i32, i32, i32, vp, vp]
    lib.ds_ban_run.argtypes = [C.c_longlong, i32, vp, vp, dbl, vp, vp]
    lib.ds_mvdr_from_cov_run.argtypes = [C.c_longlong, i32, vp, vp, vp, vp]
    lib.ds_apply_stream_weights_run.argtypes = [i32, i32, i32, i32, vp, i32, vp, vp, vp]
    lib.ds_idoa_rtf_state_bytes.argtypes = [i32, i32, i32]
    lib.ds_idoa_rtf_state_bytes.restype = C.c_size_t
    lib.ds_idoa_spp_state_bytes.argtypes = [i32, i32, i32]
    lib.ds_idoa_spp_state_bytes.restype = C.c_size_t
    lib.ds_idoa_rtf_run.argtypes = [i32, i32, i32, i32, dbl, vp, i32, vp, vp, vp]
    lib.ds_idoa_spp_run.argtypes = [i32, i32, i32, i32, i32, vp, i32, vp, vp, vp, vp, vp, i32, vp, vp]
    lib.ds_chain_state_bytes.argtypes = [C.POINTER(ChainParams)]
    lib.ds_chain_state_bytes.restype = C.c_size_t
    lib.ds_chain_workspace_bytes.argtypes = [C.POINTER(ChainParams)]
    lib.ds_chain_workspace_bytes.restype = C.c_size_t
    lib.ds_chain_run.argtypes = [C.POINTER(ChainParams), vp, vp, vp, vp, vp, vp, vp]
    lib.ds_chain_run_profiled.argtypes = [C.POINTER(ChainParams), vp, vp, vp, vp, vp, vp, vp, C.POINTER(C.c_float)]
    lib.ds_chain_run_io.argtypes = [C.POINTER(ChainParams), vp, vp, vp, vp, vp, i32, vp, i32, vp]
    lib.ds_stft_pcm16_run.argtypes = [C.POINTER(StftParams), vp, vp, vp, vp, vp]
    lib.ds_istft_pcm16_run.argtypes = [C.POINTER(IstftParams), vp, vp, vp, vp, vp]
    lib.ds_multibeam_tc_layout.argtypes = [i32, i32, i32, i32, i32, C.POINTER(C.c_int32), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    lib.ds_multibeam_tc_run.argtypes = [i32, i32, i32, i32, i32, vp, vp, vp, vp, vp]
    lib.ds_istft_frames_inner_run.argtypes = [C.POINTER(IstftParams), vp, vp, vp, C.c_longlong, vp, vp]
    lib.ds_wpe_state_bytes.argtypes = [i32, i32, i32, i32, i32]
    lib.ds_wpe_state_bytes.restype = C.c_size_t
    lib.ds_wpe_run.argtypes = [i32, i32, i32, i32, i32, i32, dbl, dbl, vp, vp, i32, vp, vp]
    lib.ds_memcpy2d_async.argtypes = [vp, C.c_size_t, vp, C.c_size_t, C.c_size_t, C.c_size_t, i32, vp]
    lib.ds_fp64_peak_run.argtypes = [i32, vp, vp]
    lib.ds_fp64_peak_run.restype = C.c_double
    lib.ds_double_to_pcm16_run.argtypes = [C.c_size_t, vp, vp, vp]


def lib():
    """Load the shared library (raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DsError("libds_b200.so not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `python distantspeech_b200/_build.py` (there is no CPU fallback)")
        _lib = C.CDLL(LIB_PATH)
        _declare(_lib)
    return _lib


def check(rc: int, what: str = ""):
    if rc != DS_OK:
        msg = lib().ds_last_error().decode("utf-8", "replace")
        raise DsError("%s failed (%d): %s" % (what or "ds call", rc, msg))


# ---------------------------------------------------------------------------
# torch hand-off helpers
# ---------------------------------------------------------------------------
_torch = None


def torch():
    global _torch
    if _torch is None:
        import torch as _t
        _torch = _t
    return _torch


def require_cuda():
    t = torch()
    if not t.cuda.is_available():
        raise DsError("distantspeech_b200 needs a CUDA device (sm_100a); no CPU fallback exists")
    return t


_inited = set()


def ensure_init():
    t = require_cuda()
    dev = t.cuda.current_device()
    if dev not in _inited:
        check(lib().ds_init(), "ds_init")
        _inited.add(dev)


def stream_ptr():
    return C.c_void_p(torch().cuda.current_stream().cuda_stream)


def ptr(tensor):
    return C.c_void_p(tensor.data_ptr()) if tensor is not None else C.c_void_p(0)


def to_device(a, dtype):
    """numpy / torch -> contiguous CUDA tensor of `dtype` (torch dtype)."""
    t = require_cuda()
    if isinstance(a, t.Tensor):
        return a.to(device="cuda", dtype=dtype).contiguous()
    return t.as_tensor(np.ascontiguousarray(a)).to(device="cuda", dtype=dtype).contiguous()


_window_cache = {}


def device_window(window: np.ndarray, n_fft: int):
    """float64 window, centre-padded to n_fft (librosa.util.pad_center), cached on device."""
    t = require_cuda()
    w = np.asarray(window, dtype=np.float64)
    if w.shape[0] != n_fft:
        lpad = (n_fft - w.shape[0]) // 2
        w = np.pad(w, (lpad, n_fft - w.shape[0] - lpad))
    key = (t.cuda.current_device(), n_fft, w.tobytes())
    if key not in _window_cache:
        while len(_window_cache) >= 64:
            _window_cache.pop(next(iter(_window_cache)))      # evict the oldest entry only
        _window_cache[key] = t.as_tensor(w).to("cuda")
    return _window_cache[key]


def spectral_power(Xd, via_abs=False):
    """|X|^2 of a complex64/complex128 CUDA tensor on the device (ds_power_run) -> float64 tensor of the same shape."""
    t = require_cuda()
    Xd = Xd.contiguous()
    out = t.empty(Xd.shape, dtype=t.float64, device=Xd.device)
    if Xd.numel():
        check(lib().ds_power_run(Xd.numel(), ptr(Xd), int(Xd.dtype == t.complex128), int(bool(via_abs)), ptr(out), stream_ptr()),
              "ds_power_run")
    return out
