"""McSpp -- drop-in for ``DistantSpeech/noise_estimation/mcspp.py`` (McSpp :46, estimation :248) with its
coherent-to-diffuse-ratio prior ``McCDR`` (``noise_estimation/mccdr.py`` :25, estimation :164).

Per frame: q = 1 - McCDR.estimation(y); diagonal loading from the 500 Hz..2 kHz average of q;
Phi_yy recursion; complex 4x4 inverse of herm(Phi_vv) + loading I with the reference's fallback
where xi < 0; xi, gamma, posterior p; noise-covariance update; PMWF weights with beta = 10.
Everything runs in CUDA (``ds_mcspp_cdr_run``); the float64 state lives on the device.

As shipped the reference only works with 4 channels: ``McSpp.__init__`` builds ``McCDR(nfft)`` with its default
``channels=4`` (mcspp.py:54) whose PSD tracker then raises IndexError for more microphones.  Here 4 ... 8 channels
run ("lift the M <= 4 limit"): above 4 the behaviour is the reference's with ``McCDR(nfft,
channels=M)`` handed in (the pin: oracle/ref_harness.make_mcspp, tests/golden/mcspp_cdr_m68.npz).  With 7 or 8
channels the patched reference still raises LinAlgError in frames 5 .. M-2 (the xi < 0 fallback inverse of
mcspp.py:224-227 drops its loading after 5 frames, when Phi_yy is still rank deficient); here the loading is kept until
Phi_yy has full rank (``fallback_loaded_frames = max(5, M - 1)``: the reference's behaviour for M <= 6).  Fewer than 4
channels raise ValueError (the CDR's microphone pair (1, 2) is read at the 4-channel pair index and is 0/0 below).
Extensions: a leading stream axis, ``estimation_frames`` (many frames per launch), ``repeat=True`` on the device.
"""
import ctypes as C

import numpy as np

from .. import _lib as L
from ..beamformer.MicArray import MicArray
from ..beamformer.gen_noise_msc import gen_noise_msc


class _CdrMcra(object):
    """The MCRA tracker McCDR owns (mccdr.py:55-56); state is part of the device blob."""

    def __init__(self, owner):
        self._o = owner
        self.L = 65
        self.alpha_d, self.alpha_s, self.delta_s, self.alpha_p = 0.95, 0.8, 5, 0.2
        self.p_max, self.p_min = 0.999, 1e-3
        self.ell = 1
        self.frm_cnt = 0

    def _field(self, i):
        blk = self._o._export(7)
        if blk is None:
            return np.zeros(self._o.half_bin)
        v = blk[:, i, :]
        return v[0] if v.shape[0] == 1 else v

    S = property(lambda self: self._field(0))
    Smin = property(lambda self: self._field(1))
    Stmp = property(lambda self: self._field(2))
    p = property(lambda self: self._field(3))
    lambda_d = property(lambda self: self._field(4))


class _CdrCore(object):
    """Device state + launches shared by McSpp and McCDR."""

    def __init__(self, nfft, channels):
        if not 4 <= channels <= 8:
            raise ValueError("McSpp / McCDR run with 4..8 channels: pair (1, 2) of the CDR is undefined below 4 in the "
                             "reference (mcspp.py:54 builds McCDR with channels=4); the kernels are compiled up to 8")
        self.nfft, self.channels = int(nfft), int(channels)
        self.half_bin = int(self.nfft / 2 + 1)
        self.alpha = self.alpha_d = 0.92
        self.mcra = _CdrMcra(self)
        self.MicArray = MicArray(arrayType="circular", r=0.032, M=self.channels)  # mccdr.py:58 (channels = 4 as shipped)
        self.Fvv = gen_noise_msc(mic=self.MicArray, nfft=self.nfft)                # diffuse model of that array
        self._Fn_dev = None
        self._state = None
        self._S = None
        self.frm_cnt = 0

    def _params(self, S, T, cdr_only=0):
        p = L.McsppCdrParams()
        L.lib().ds_mcspp_cdr_default_params(C.byref(p), self.nfft, S, self.channels, T)
        m = self.mcra
        p.frm_cnt, p.ell, p.mcra_L, p.cdr_only = int(m.frm_cnt), int(m.ell), int(m.L), int(cdr_only)
        p.alpha, p.alpha_d = float(self.alpha), float(self.alpha_d)
        p.fallback_loaded_frames = max(5, self.channels - 1)
        p.mcra_alpha_d, p.mcra_alpha_s, p.mcra_delta_s = float(m.alpha_d), float(m.alpha_s), float(m.delta_s)
        p.mcra_alpha_p, p.mcra_p_min, p.mcra_p_max = float(m.alpha_p), float(m.p_min), float(m.p_max)
        return p

    def _ensure(self, S):
        t = L.require_cuda()
        L.ensure_init()
        if self._state is None or self._S != S:
            nbytes = L.lib().ds_mcspp_cdr_state_bytes(C.byref(self._params(S, 1)))
            self._state = t.zeros(nbytes, dtype=t.uint8, device="cuda")
            self._S = S
        if self._Fn_dev is None:
            self._Fn_dev = t.as_tensor(np.ascontiguousarray(self.Fvv[:, 1, 2], dtype=np.float64)).to("cuda")

    def _export(self, field):
        if self._state is None:
            return None
        t = L.require_cuda()
        S, K, M = self._S, self.half_bin, self.channels
        if field <= 3:
            out = t.empty((S, K, M, M), dtype=t.complex128, device="cuda")
        else:
            rows = {4: 2 * M, 5: 4, 6: M + M * (M - 1), 7: 6}[field]
            out = t.empty((S, rows, K), dtype=t.float64, device="cuda")
        L.check(L.lib().ds_mcspp_cdr_export(C.byref(self._params(S, 1)), L.ptr(self._state), field, L.ptr(out),
                                            L.stream_ptr()), "ds_mcspp_cdr_export")
        return out.cpu().numpy()

    @staticmethod
    def _to_stmk(y):
        """[K, M] | [S, K, M] one frame, complex -> [S, 1, M, K] CUDA."""
        t = L.require_cuda()
        yd = y.to("cuda") if isinstance(y, t.Tensor) else t.as_tensor(
            np.ascontiguousarray(np.asarray(y, dtype=np.complex128))).to("cuda")
        if yd.dtype not in (t.complex64, t.complex128):
            yd = yd.to(t.complex128)
        if yd.dim() == 2:
            yd = yd[None]
        return yd.permute(0, 2, 1)[:, None, :, :].contiguous()

    def _run(self, Xd, cdr_only=False, want_Y=False, want_w=True, repeat=False):
        t = L.require_cuda()
        S, T, M, K = Xd.shape
        if M != self.channels or K != self.half_bin:
            raise ValueError("expected [.., %d bins, %d channels]" % (self.half_bin, self.channels))
        self._ensure(S)
        prm = self._params(S, T, cdr_only=int(bool(cdr_only)) | (2 if repeat else 0))
        ws = t.empty(L.lib().ds_mcspp_cdr_workspace_bytes(C.byref(prm)), dtype=t.uint8, device="cuda")
        f64 = dict(dtype=t.float64, device="cuda")
        out = {"cdr": t.empty((S, T, K), **f64)}
        if not cdr_only:
            for k in ("p", "xi", "gamma", "q"):
                out[k] = t.empty((S, T, K), **f64)
            if want_w:
                out["w"] = t.empty((S, T, M, K), dtype=t.complex128, device="cuda")
            if want_Y:
                out["Y"] = t.empty((S, T, K), dtype=t.complex64, device="cuda")
        addr = (lambda k: out[k].data_ptr() if k in out else None)
        taps = L.McsppCdrTaps(addr("p"), addr("xi"), addr("gamma"), addr("q"), addr("cdr"), addr("w"))
        L.check(L.lib().ds_mcspp_cdr_run(C.byref(prm), L.ptr(self._state), L.ptr(ws), L.ptr(self._Fn_dev), L.ptr(Xd),
                                         int(Xd.dtype == t.complex128), L.ptr(out.get("Y")), C.byref(taps),
                                         L.stream_ptr()), "ds_mcspp_cdr_run")
        f, e = C.c_int32(self.mcra.frm_cnt), C.c_int32(self.mcra.ell)
        L.lib().ds_mcra_advance(int(self.mcra.L), T, C.byref(f), C.byref(e))
        self.mcra.frm_cnt, self.mcra.ell = f.value, e.value
        self.frm_cnt += T
        return out

    def reset(self):
        self._state = None
        self.mcra.frm_cnt, self.mcra.ell = 0, 1
        self.frm_cnt = 0


def _sq(v):
    return v[0] if v.shape[0] == 1 else v


class McSpp(_CdrCore):
    def __init__(self, nfft=256, channels=4, mic_array=None) -> None:
        super().__init__(nfft, channels)
        K, M = self.half_bin, self.channels
        self.q = np.ones(K) * 0.6
        self.p = np.zeros(K)
        self.xi = np.zeros(K)
        self.xi_last = np.zeros(K)
        self.gamma = np.zeros(K)
        self.w = np.zeros((K, M), dtype=complex)
        self.mccdr = McCDR.__new__(McCDR)          # view on the same device state, like self.mccdr in the reference
        self.mccdr.__dict__ = self.__dict__
        if mic_array is not None:
            self.mic_array = mic_array
            self.steer_vector = mic_array.steering_vector(look_direction=30)       # mcspp.py:67-69

    def _mat(self, field):
        v = self._export(field)
        if v is None:
            return np.zeros((self.half_bin, self.channels, self.channels), dtype=complex)
        return _sq(v)

    Phi_yy = property(lambda self: self._mat(0))
    Phi_vv = property(lambda self: self._mat(1))
    Phi_vv_inv = property(lambda self: self._mat(2))
    Phi_xx = property(lambda self: self._mat(3))

    def _publish(self, out):
        last = {k: _sq(out[k][:, -1].cpu().numpy()) for k in ("p", "xi", "gamma", "q")}
        self.p, self.xi, self.gamma, self.q = last["p"], last["xi"], last["gamma"], last["q"]
        self.xi_last = self.xi.copy()
        if "w" in out:
            self.w = _sq(out["w"][:, -1].permute(0, 2, 1).cpu().numpy())            # [K, M]

    def estimation(self, y, diag_value=1e-4, repeat=False):
        """One frame: y [K, M] complex (or [S, K, M]) -> posterior SPP p [K]  (mcspp.py:248-305).
        ``diag_value`` is ignored like in the reference (overwritten at :269); ``repeat=True`` runs estimation_core a
        second time on the updated noise covariance (:282-284)."""
        out = self._run(self._to_stmk(y), repeat=repeat)
        self._publish(out)
        return self.p

    def estimation_frames(self, D, want_Y=False, repeat=False):
        """Extension: D [K, T, M] (or [S, K, T, M]) complex -> dict of per-frame arrays p, xi, gamma, q, cdr
        [K, T], w [K, T, M] and (want_Y) the PMWF output spectrum Y = w^H y [K, T]."""
        t = L.require_cuda()
        Dd = D.to("cuda") if isinstance(D, t.Tensor) else t.as_tensor(
            np.ascontiguousarray(np.asarray(D, dtype=np.complex128))).to("cuda")
        batched = Dd.dim() == 4
        if not batched:
            Dd = Dd[None]
        out = self._run(Dd.permute(0, 2, 3, 1).contiguous(), want_Y=want_Y, repeat=repeat)
        self._publish(out)
        res = {k: out[k].permute(0, 2, 1) for k in ("p", "xi", "gamma", "q", "cdr")}
        res["w"] = out["w"].permute(0, 3, 1, 2)
        if want_Y:
            res["Y"] = out["Y"].permute(0, 2, 1)
        if not batched:
            res = {k: v[0] for k, v in res.items()}
        return {k: v.cpu().numpy() for k, v in res.items()}

    def compute_q(self, y, q_max=0.99, q_min=0.01):
        raise NotImplementedError("fused into estimation(); the prior of the last frame is the attribute q")

    def compute_pmwf_weight(self, xi, Rxx, Rvv_inv, Gmin=0.0631, beta=1):
        from ..beamformer.beamformer import compute_pmwf_weight
        self.w = compute_pmwf_weight(xi, Rxx, Rvv_inv, Gmin=Gmin, beta=beta)


class McCDR(_CdrCore):
    """``McCDR(nfft).estimation(y) -> sqrt(CDR^2 * p_mcra)`` (mccdr.py:164-177)."""

    def __init__(self, nfft=256, channels=4) -> None:
        super().__init__(nfft, channels)
        self.alpha_d = 0.95
        self.Gamma = np.zeros(self.half_bin)

    def estimation(self, y, theta=135):
        out = self._run(self._to_stmk(y), cdr_only=True)
        self.Gamma = _sq(out["cdr"][:, -1].cpu().numpy())
        return self.Gamma

    def estimate_ddr(self, y, unbias=True):
        raise NotImplementedError("fused into estimation(); the squared, clipped CDR of the last frame is exported "
                                  "with the state (ds_mcspp_cdr_export field 5)")
