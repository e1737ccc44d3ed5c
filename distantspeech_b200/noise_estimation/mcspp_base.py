"""Multichannel speech-presence-probability tracker -- drop-in for
``DistantSpeech/noise_estimation/mcspp_base.py`` (McSppBase :29, estimation :262,
compute_omlsa_weight :140, compute_pmwf_weight :220, update_noise_psd :299).

``estimation(y[K, M])`` processes one frame like the reference and keeps every
public attribute (Phi_yy, Phi_vv, Phi_vv_inv, Phi_xx, xi, gamma, q, p, w, G) in
sync; the float64 state lives on the device.  Extensions: a leading stream axis
and ``estimation_frames`` (many frames per launch).
"""
import ctypes as C

import numpy as np

from .. import _lib as L


class _InnerMcra(object):
    """View of the MCRA tracker embedded in the McSppBase state (mcspp_base.py:76-77)."""

    def __init__(self, owner):
        self._o = owner
        self.L = 15
        self.alpha_d, self.alpha_s, self.delta_s, self.alpha_p = 0.95, 0.8, 5, 0.2
        self.p_max, self.p_min = 0.999, 1e-3
        self.ell = 1
        self.frm_cnt = 0

    def _field(self, i):
        blk = self._o._export(2)
        if blk is None:
            return np.zeros(self._o.half_bin)
        v = blk[:, i, :]
        return v[0] if v.shape[0] == 1 else v

    S = property(lambda self: self._field(0))
    Smin = property(lambda self: self._field(1))
    Stmp = property(lambda self: self._field(2))
    p = property(lambda self: self._field(3))
    lambda_d = property(lambda self: self._field(4))


class McSppBase(object):
    def __init__(self, nfft=256, channels=4) -> None:
        self.channels = channels
        self.nfft = nfft
        self.half_bin = int(self.nfft / 2 + 1)
        self.alpha_d = 0.92
        self.alpha = 0.92
        self.alpha_s = 0.8
        self.delta_s = 5
        self.alpha_p = 0.2
        self.L = 125
        self.diagonal_eps = np.eye(self.channels) * 1e-6
        self.mcra = _InnerMcra(self)
        self.frm_cnt = 0
        K, M = self.half_bin, channels
        self.q = np.ones(K) * 0.6
        self.p = np.zeros(K)
        self.G_H1 = np.zeros(K)
        self.G = np.zeros(K)
        self.xi = np.zeros(K)
        self.gamma = np.zeros(K)
        self.w = np.zeros((K, M), dtype=complex)
        self.Phi_vv_inv = np.zeros((K, M, M), dtype=complex)
        self._state = None
        self._S = None
        self._prev_vv = None
        self._last_G_args = None

    # ---- device state ------------------------------------------------------------
    def _params(self, S, T, full=1):
        p = L.McsppParams()
        L.lib().ds_mcspp_default_params(C.byref(p), self.nfft, S, self.channels, T)
        p.frm_cnt, p.ell, p.mcra_L, p.full_state = int(self.mcra.frm_cnt), int(self.mcra.ell), int(self.mcra.L), full
        p.alpha, p.alpha_d = float(self.alpha), float(self.alpha_d)
        p.diag_eps = float(self.diagonal_eps[0, 0])
        m = self.mcra
        p.mcra_alpha_d, p.mcra_alpha_s, p.mcra_delta_s = float(m.alpha_d), float(m.alpha_s), float(m.delta_s)
        p.mcra_alpha_p, p.mcra_p_min, p.mcra_p_max = float(m.alpha_p), float(m.p_min), float(m.p_max)
        return p

    def _ensure(self, S):
        t = L.require_cuda()
        if self._state is None or self._S != S:
            if self._state is not None:
                self.mcra.frm_cnt, self.mcra.ell = 0, 1      # re-zeroed state = new streams: the inner MCRA's counters restart too
            nbytes = L.lib().ds_mcspp_state_bytes(C.byref(self._params(S, 1)))
            self._state = t.zeros(nbytes, dtype=t.uint8, device="cuda")
            self._S = S

    def _export(self, field, as_numpy=True):
        if self._state is None:
            return None
        t = L.require_cuda()
        S, K, M = self._S, self.half_bin, self.channels
        if field == 2:
            out = t.empty((S, 5, K), dtype=t.float64, device="cuda")
        else:
            out = t.empty((S, K, M, M), dtype=t.complex128, device="cuda")
        L.check(L.lib().ds_mcspp_export(C.byref(self._params(S, 1)), L.ptr(self._state), field, L.ptr(out),
                                        L.stream_ptr()), "ds_mcspp_export")
        return out.cpu().numpy() if as_numpy else out

    def _mat(self, field):
        v = self._export(field)
        if v is None:
            return np.zeros((self.half_bin, self.channels, self.channels), dtype=complex)
        return v[0] if v.shape[0] == 1 else v

    Phi_yy = property(lambda self: self._mat(0))
    Phi_vv = property(lambda self: self._mat(1))

    @property
    def Phi_xx(self):
        """Phi_yy (after the last frame) - Phi_vv (before its update)  (mcspp_base.py:274)."""
        if self._state is None or self._prev_vv is None:
            return np.zeros((self.half_bin, self.channels, self.channels), dtype=complex)
        v = (self._export(0, as_numpy=False) - self._prev_vv).cpu().numpy()
        return v[0] if v.shape[0] == 1 else v

    # ---- estimation ----------------------------------------------------------------
    def _run(self, Xd, a0=None, want_Y=False, apply_gain=True, keep_prev_vv=True):
        """Xd [S, T, M, K] complex64/complex128 CUDA.  Returns dict of device tensors."""
        t = L.require_cuda()
        S, T, M, K = Xd.shape
        assert M == self.channels and K == self.half_bin
        self._ensure(S)
        if keep_prev_vv:
            self._prev_vv = self._export(1, as_numpy=False) if T == 1 else None
        prm = self._params(S, T, full=1)
        f64 = dict(dtype=t.float64, device="cuda")
        out = {k: t.empty((S, T, K), **f64) for k in ("p", "xi", "gamma", "q", "G")}
        out["w_pmwf"] = t.empty((S, T, M, K), dtype=t.complex128, device="cuda")
        out["Ainv"] = t.empty((S, K, M, M), **f64)
        a0d = None
        if a0 is not None:
            a0d = L.to_device(np.ascontiguousarray(np.asarray(a0, dtype=np.complex128).T), t.complex128)   # [M, K]
            out["w_mvdr"] = t.empty((S, T, M, K), dtype=t.complex128, device="cuda")
            if want_Y:
                out["Y"] = t.empty((S, T, K), dtype=t.complex64, device="cuda")
        taps = L.McsppTaps(out["p"].data_ptr(), out["xi"].data_ptr(), out["gamma"].data_ptr(), out["q"].data_ptr(),
                           out["G"].data_ptr(), out["w_mvdr"].data_ptr() if "w_mvdr" in out else None,
                           out["w_pmwf"].data_ptr(), out["Ainv"].data_ptr())
        L.check(L.lib().ds_mcspp_run(C.byref(prm), L.ptr(self._state), L.ptr(a0d), L.ptr(Xd),
                                     int(Xd.dtype == t.complex128), L.ptr(out.get("Y")), int(apply_gain),
                                     C.byref(taps), L.stream_ptr()), "ds_mcspp_run")
        f, e = C.c_int32(self.mcra.frm_cnt), C.c_int32(self.mcra.ell)
        L.lib().ds_mcra_advance(int(self.mcra.L), T, C.byref(f), C.byref(e))
        self.mcra.frm_cnt, self.mcra.ell = f.value, e.value
        self.frm_cnt += T
        return out

    def _publish(self, out):
        """Mirror the last frame of a run into the reference's public attributes."""
        sq = (lambda v: v[0] if v.shape[0] == 1 else v)
        last = {k: out[k][:, -1].cpu().numpy() for k in ("p", "xi", "gamma", "q", "G")}
        self.p, self.xi, self.gamma, self.q = sq(last["p"]), sq(last["xi"]), sq(last["gamma"]), sq(last["q"])
        self._G_fused = sq(last["G"])
        self._last_G_args = (self.xi, self.p)
        self.w = sq(out["w_pmwf"][:, -1].permute(0, 2, 1).cpu().numpy())            # [K, M]
        self.Phi_vv_inv = sq(out["Ainv"].cpu().numpy().astype(complex))

    def estimation(self, y):
        """One frame: y [K, M] complex (or [S, K, M]) -> posterior SPP p [K]  (mcspp_base.py:262-297)."""
        t = L.require_cuda()
        if isinstance(y, t.Tensor):
            yd = y.to("cuda")
        else:
            yd = t.as_tensor(np.ascontiguousarray(np.asarray(y, dtype=np.complex128))).to("cuda")
        if yd.dtype not in (t.complex64, t.complex128):
            yd = yd.to(t.complex128)
        if yd.dim() == 2:
            yd = yd[None]
        Xd = yd.permute(0, 2, 1)[:, None, :, :].contiguous()                       # [S, 1, M, K]
        out = self._run(Xd)
        self._publish(out)
        return self.p

    def estimation_frames(self, D, a0=None, apply_gain=True):
        """Extension: D [K, T, M] (or [S, K, T, M]) complex -> dict of per-frame arrays
        p, xi, gamma, q, G [K, T]; with ``a0`` [K, M] also w_mvdr [K, T, M] and the
        beamformed (and gained) spectrum Y [K, T]."""
        t = L.require_cuda()
        if isinstance(D, t.Tensor):
            Dd = D.to("cuda")
        else:
            Dd = t.as_tensor(np.ascontiguousarray(np.asarray(D, dtype=np.complex128))).to("cuda")
        batched = Dd.dim() == 4
        if not batched:
            Dd = Dd[None]
        Xd = Dd.permute(0, 2, 3, 1).contiguous()                                    # [S, T, M, K]
        out = self._run(Xd, a0=a0, want_Y=a0 is not None, apply_gain=apply_gain, keep_prev_vv=False)
        self._publish(out)
        res = {}
        for k in ("p", "xi", "gamma", "q", "G"):
            res[k] = out[k].permute(0, 2, 1)                                        # [S, K, T]
        res["w_pmwf"] = out["w_pmwf"].permute(0, 3, 1, 2)                           # [S, K, T, M]
        if a0 is not None:
            res["w_mvdr"] = out["w_mvdr"].permute(0, 3, 1, 2)
            res["Y"] = out["Y"].permute(0, 2, 1)
        if not batched:
            res = {k: v[0] for k, v in res.items()}
        return {k: v.cpu().numpy() for k, v in res.items()}

    # ---- the remaining reference methods ---------------------------------------------
    def estimate_noisy_psd(self, y, alpha=0.92):
        raise NotImplementedError("fused into estimation(); Phi_yy is available as an attribute")

    def compute_omlsa_weight(self, xi, p, Gmin=0.0631):
        """G = clip((xi/(1+xi))^p Gmin^(1-p), Gmin, 1), G[:2] = 0  (mcspp_base.py:140-155)."""
        t = L.require_cuda()
        xid = L.to_device(np.asarray(xi, dtype=np.float64), t.float64).reshape(-1, self.half_bin)
        pd = L.to_device(np.asarray(p, dtype=np.float64), t.float64).reshape(-1, self.half_bin)
        G = t.empty_like(xid)
        GH1 = t.empty_like(xid)
        L.check(L.lib().ds_omlsa_gain_run(xid.shape[0], self.half_bin, L.ptr(xid), L.ptr(pd), float(Gmin), L.ptr(G),
                                          L.ptr(GH1), L.stream_ptr()), "ds_omlsa_gain_run")
        shp = np.shape(xi)
        self.G = G.cpu().numpy().reshape(shp)
        self.G_H1 = GH1.cpu().numpy().reshape(shp)

    def compute_pmwf_weight(self, xi, Rxx, Rvv_inv, Gmin=0.0631, beta=1):
        from ..beamformer.beamformer import compute_pmwf_weight
        self.w = compute_pmwf_weight(xi, Rxx, Rvv_inv, Gmin=Gmin, beta=beta)

    def reset(self):
        self._state = None
        self._prev_vv = None
        self.mcra.frm_cnt, self.mcra.ell = 0, 1
        self.frm_cnt = 0
