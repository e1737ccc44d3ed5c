"""Multichannel OMLSA postfilter (Cohen / Gannot / Berdugo 2003) -- drop-in for
``DistantSpeech/noise_estimation/omlsa_multi.py`` (NsOmlsaMulti :28, estimation :73).
The M MCRA trackers and all per-bin state live on the device (ds_omlsa_multi_run)."""
import ctypes as C

import numpy as np

from .. import _lib as L
from .NoiseEstimationBase import NoiseEstimationBase


class NsOmlsaMulti(NoiseEstimationBase):
    def __init__(self, nfft=256, M=4, cal_weights=False) -> None:
        super(NsOmlsaMulti, self).__init__(nfft=nfft)
        self.Gmin = np.power(10, (-12 / 10))
        self.q_min = 1e-6
        self.q_max = 0.9999998
        self.alpha_d = 0.85
        self.first_frame = 1
        self.M = M
        self.win = np.array([0.25, 0.5, 0.25])
        self.alpha_s = 0.8
        self.cal_weights = cal_weights
        self.mcra_L = 15
        self.G = np.ones(self.half_bin)
        self.p = np.zeros(self.half_bin)
        self.lambda_d = np.zeros(self.half_bin)
        self._state = None
        self._S = None

    def _params(self, S, T):
        p = L.OmlsaMultiParams()
        L.lib().ds_omlsa_multi_default_params(C.byref(p), self.half_bin, S, T, self.M)
        p.first_frame, p.frm_cnt, p.ell, p.mcra_L = int(self.first_frame), int(self.frm_cnt), int(self.ell), int(self.mcra_L)
        p.cal_weights = int(bool(self.cal_weights))
        p.alpha_d, p.alpha_s, p.Gmin = float(self.alpha_d), float(self.alpha_s), float(self.Gmin)
        p.q_min, p.q_max = float(self.q_min), float(self.q_max)
        return p

    def _run(self, yd, ud, u_const=False):
        """yd [S, T, K], ud [S, T, M-1, K] (u_const: [S, M-1, K], shared by all frames) float64 CUDA
        -> (G, lambda_d, p) [S, T, K]."""
        t = L.require_cuda()
        S, T, K = yd.shape
        assert K == self.half_bin
        if self._state is None or self._S != S:
            self._state = t.zeros(L.lib().ds_omlsa_multi_state_bytes(C.byref(self._params(S, T))), dtype=t.uint8, device="cuda")
            # reference initial values (:31-52): G_H1 = G = gamma = 1
            st = self._state.view(t.float64).view(S, -1, K)
            o = 5 * self.M + self.M
            st[:, o + 1] = 1.0   # gamma
            st[:, o + 2] = 1.0   # G_H1
            st[:, o + 4] = 1.0   # G
            st[:, o + 5] = 1.0   # q_hat
            st[:, o + 6] = 1.0   # xi_hat
            self._S = S
        prm = self._params(S, T)
        prm.u_const = int(bool(u_const))
        G = t.empty((S, T, K), dtype=t.float64, device="cuda")
        lam = t.empty_like(G)
        p = t.empty_like(G)
        L.check(L.lib().ds_omlsa_multi_run(C.byref(prm), L.ptr(self._state), L.ptr(yd), L.ptr(ud), L.ptr(G), L.ptr(lam),
                                           L.ptr(p), L.stream_ptr()), "ds_omlsa_multi_run")
        f, e = C.c_int32(self.frm_cnt), C.c_int32(self.ell)
        L.lib().ds_mcra_advance(int(self.mcra_L), T, C.byref(f), C.byref(e))
        self.frm_cnt, self.ell = f.value, e.value
        self.first_frame = 0
        return G, lam, p

    def estimation(self, y, u):
        """y [K] beamformer-output power, u [K, M-1] reference powers -> lambda_d [K]
        (``None`` on the very first frame, like the reference :87-93)."""
        t = L.require_cuda()
        assert len(y) == self.half_bin
        was_first = self.first_frame == 1
        yd = t.as_tensor(np.ascontiguousarray(y, dtype=np.float64)).to("cuda").reshape(1, 1, -1)
        ud = t.as_tensor(np.ascontiguousarray(np.asarray(u, dtype=np.float64).T)).to("cuda")[None, None]
        G, lam, p = self._run(yd, ud.contiguous())
        self.G, self.lambda_d, self.p = G[0, 0].cpu().numpy(), lam[0, 0].cpu().numpy(), p[0, 0].cpu().numpy()
        return None if was_first else self.lambda_d

    def estimation_frames(self, Y, U):
        """Extension: Y [T, K], U [T, K, M-1] -> dict(G, lambda_d, p) each [T, K]."""
        t = L.require_cuda()
        yd = t.as_tensor(np.ascontiguousarray(Y, dtype=np.float64)).to("cuda")[None]
        ud = t.as_tensor(np.ascontiguousarray(np.asarray(U, dtype=np.float64).transpose(0, 2, 1))).to("cuda")[None]
        G, lam, p = self._run(yd, ud.contiguous())
        self.G, self.lambda_d, self.p = G[0, -1].cpu().numpy(), lam[0, -1].cpu().numpy(), p[0, -1].cpu().numpy()
        return {"G": G[0].cpu().numpy(), "lambda_d": lam[0].cpu().numpy(), "p": p[0].cpu().numpy()}
