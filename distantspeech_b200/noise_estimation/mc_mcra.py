"""McMcra -- drop-in for ``DistantSpeech/noise_estimation/mc_mcra.py`` (McMcra :25, estimation :179).

Multichannel speech-presence tracker on the REAL parts of the PSD matrices with the local threshold prior of
``compute_q_local`` (:89-103) and an OMLSA-style gain (:152-156); it is the presence detector and postfilter of
the frequency-domain GSC (``beamformer/GSC.py``).  State and arithmetic live on the device (``ds_gsc_run`` without
propagation vectors).  Matrix attributes keep the reference's ``[M, M, K]`` layout.
Extensions: a leading stream axis and ``estimation_frames``.
"""
import ctypes as C

import numpy as np

from .. import _lib as L


class McMcra(object):
    def __init__(self, nfft=256, channels=4) -> None:
        if not 2 <= channels <= 8:
            raise ValueError("McMcra on the device is compiled for 2..8 channels")
        self.channels = self.M = channels
        self.nfft = nfft
        self.half_bin = int(self.nfft / 2 + 1)
        self.alpha_d = 0.95
        self.alpha = 0.92
        self.q_max, self.q_min = 0.99, 0            # attributes of the reference; compute_q_local uses .99 / .01
        self.psi_0 = 100
        self.psi_tilde_0 = 100
        K = self.half_bin
        self.q = np.ones(K) * 0.6
        self.q_local = np.ones(K) * 0.999
        self.p = np.zeros(K)
        self.G = np.zeros(K)
        self.xi = np.zeros(K)
        self.gamma = np.zeros(K)
        self.frm_cnt = 0
        self._state = None
        self._S = None

    # ---- device state ------------------------------------------------------------------
    def _params(self, S, T):
        p = L.GscParams()
        L.lib().ds_gsc_default_params(C.byref(p), self.nfft, S, self.channels, T)
        p.frm_cnt = int(self.frm_cnt)
        p.alpha, p.alpha_d, p.psi_0 = float(self.alpha), float(self.alpha_d), float(self.psi_0)
        return p

    def _ensure(self, S):
        t = L.require_cuda()
        L.ensure_init()
        if self._state is None or self._S != S:
            self._state = t.zeros(L.lib().ds_gsc_state_bytes(C.byref(self._params(S, 1))), dtype=t.uint8, device="cuda")
            self._S = S
            self.frm_cnt = 0

    def _blob(self):
        t = L.require_cuda()
        return self._state.view(t.float64).view(self._S, -1, self.half_bin)        # [S, NE, K]

    def _matrix(self, which):
        """Phi_yy (0) / Phi_vv (1) unpacked to the reference layout [M, M, K] (or [S, M, M, K])."""
        M, K = self.channels, self.half_bin
        if self._state is None:
            return np.zeros((M, M, K))
        NP = M * (M + 1) // 2
        tri = self._blob()[:, which * NP:(which + 1) * NP, :].cpu().numpy()        # [S, NP, K]
        out = np.zeros((self._S, M, M, K))
        e = 0
        for i in range(M):
            for j in range(i, M):
                out[:, i, j] = tri[:, e]
                out[:, j, i] = tri[:, e]
                e += 1
        return out[0] if self._S == 1 else out

    Phi_yy = property(lambda self: self._matrix(0))
    Phi_vv = property(lambda self: self._matrix(1))
    Phi_xx = property(lambda self: self._matrix(0) - self._matrix(1))

    def _run(self, Xd, a_dev=None, want_Y=False, method=2):
        """Xd [S, T, M, K] complex CUDA -> dict of device tensors (p, G, xi, gamma, q [S, T, K]; Y [S, T, K])."""
        t = L.require_cuda()
        S, T, M, K = Xd.shape
        if M != self.channels or K != self.half_bin:
            raise ValueError("expected [.., %d bins, %d channels]" % (self.half_bin, self.channels))
        self._ensure(S)
        prm = self._params(S, T)
        prm.method = int(method)
        out = {k: t.empty((S, T, K), dtype=t.float64, device="cuda") for k in ("p", "G", "xi", "gamma", "q")}
        if want_Y:
            out["Y"] = t.empty((S, T, K), dtype=t.complex64, device="cuda")
        taps = L.GscTaps(*[out[k].data_ptr() for k in ("p", "G", "xi", "gamma", "q")])
        L.check(L.lib().ds_gsc_run(C.byref(prm), L.ptr(self._state), L.ptr(a_dev), L.ptr(Xd), int(Xd.dtype == t.complex128),
                                   L.ptr(out.get("Y")), C.byref(taps), L.stream_ptr()), "ds_gsc_run")
        self.frm_cnt += T
        sq = (lambda v: v[0] if v.shape[0] == 1 else v)
        last = {k: sq(out[k][:, -1].cpu().numpy()) for k in ("p", "G", "xi", "gamma", "q")}
        self.p, self.G, self.xi, self.gamma, self.q = last["p"], last["G"], last["xi"], last["gamma"], last["q"]
        self.q_local = self.q
        return out

    def estimation(self, y):
        """One frame: y [K, M] complex (or [S, K, M]); updates p, q, xi, gamma, G, Phi_yy, Phi_vv (mc_mcra.py:179-221)."""
        t = L.require_cuda()
        yd = y.to("cuda") if isinstance(y, t.Tensor) else t.as_tensor(
            np.ascontiguousarray(np.asarray(y, dtype=np.complex128))).to("cuda")
        if yd.dtype not in (t.complex64, t.complex128):
            yd = yd.to(t.complex128)
        if yd.dim() == 2:
            yd = yd[None]
        self._run(yd.permute(0, 2, 1)[:, None, :, :].contiguous())

    def estimation_frames(self, D):
        """Extension: D [K, T, M] (or [S, K, T, M]) complex -> dict of per-frame arrays p, G, xi, gamma, q [K, T]."""
        t = L.require_cuda()
        Dd = D.to("cuda") if isinstance(D, t.Tensor) else t.as_tensor(
            np.ascontiguousarray(np.asarray(D, dtype=np.complex128))).to("cuda")
        batched = Dd.dim() == 4
        if not batched:
            Dd = Dd[None]
        out = self._run(Dd.permute(0, 2, 3, 1).contiguous())
        res = {k: v.permute(0, 2, 1) for k, v in out.items()}
        if not batched:
            res = {k: v[0] for k, v in res.items()}
        return {k: v.cpu().numpy() for k, v in res.items()}

    def compute_weight(self, xi, Gmin=0.0631):
        raise NotImplementedError("fused into estimation(); the gain of the last frame is the attribute G")

    def reset(self):
        self._state = None
        self.frm_cnt = 0
