"""MCRA noise-PSD tracker -- drop-in for
``DistantSpeech/noise_estimation/mcra.py`` (NoiseEstimationMCRA :20, estimation :27).

State (S, Smin, Stmp, p, lambda_d) lives on the device in float64; every
``estimation`` call launches ds_mcra_run.  ``estimation(Y[K])`` follows the
reference; ``estimation_frames(Y[T, K])`` (extension) runs T frames in one
launch, and a leading stream axis ``[S, ...]`` batches independent streams.
"""
import ctypes as C

import numpy as np

from .. import _lib as L
from .NoiseEstimationBase import NoiseEstimationBase

_FIELDS = {"S": 0, "Smin": 1, "Stmp": 2, "p": 3, "lambda_d": 4}


class NoiseEstimationMCRA(NoiseEstimationBase):
    def __init__(self, nfft=256, p_max=0.999, p_min=1e-3) -> None:
        super(NoiseEstimationMCRA, self).__init__(nfft=nfft)
        self.p_max = p_max
        self.p_min = p_min
        self.L = 15
        self._state = None        # [S, 5, K] float64 CUDA
        self._nstreams = None

    # ---- device state ------------------------------------------------------
    def _ensure(self, S):
        t = L.require_cuda()
        if self._state is None or self._nstreams != S:
            if self._state is not None:
                self.frm_cnt, self.ell = 0, 1      # a different batch is a new set of streams: counters restart with the state
            self._state = t.zeros((S, 5, self.half_bin), dtype=t.float64, device="cuda")
            self._nstreams = S

    def _get(self, name):
        if self._state is None:
            return np.zeros(self.half_bin)
        v = self._state[:, _FIELDS[name], :].cpu().numpy()
        return v[0] if self._nstreams == 1 else v

    def _set(self, name, value):
        t = L.require_cuda()
        v = np.asarray(value, dtype=np.float64)
        self._ensure(1 if v.ndim == 1 else v.shape[0])
        self._state[:, _FIELDS[name], :] = t.as_tensor(v.reshape(self._nstreams, self.half_bin)).to("cuda")

    S = property(lambda self: self._get("S"), lambda self, v: self._set("S", v))
    Smin = property(lambda self: self._get("Smin"), lambda self, v: self._set("Smin", v))
    Stmp = property(lambda self: self._get("Stmp"), lambda self, v: self._set("Stmp", v))
    p = property(lambda self: self._get("p"), lambda self, v: self._set("p", v))
    lambda_d = property(lambda self: self._get("lambda_d"), lambda self, v: self._set("lambda_d", v))

    @property
    def alpha_tilde(self):
        return self.alpha_d + (1 - self.alpha_d) * self.p

    # ---- kernels -------------------------------------------------------------
    def _run(self, Yd, want_p=False):
        """Yd [S, T, K] float64 CUDA -> lambda_d [S, T, K] (and p)."""
        t = L.require_cuda()
        S, T, K = Yd.shape
        assert K == self.half_bin, 'len(Y):{} != half_bin:{}'.format(K, self.half_bin)
        self._ensure(S)
        prm = L.McraParams(K, S, T, int(self.L), int(self.frm_cnt), int(self.ell), 0, 0, float(self.alpha_d),
                           float(self.alpha_s), float(self.delta_s), float(self.alpha_p), float(self.p_min),
                           float(self.p_max))
        lam = t.empty((S, T, K), dtype=t.float64, device="cuda")
        pout = t.empty((S, T, K), dtype=t.float64, device="cuda") if want_p else None
        L.check(L.lib().ds_mcra_run(C.byref(prm), L.ptr(self._state), L.ptr(Yd), L.ptr(lam), L.ptr(pout),
                                    L.stream_ptr()), "ds_mcra_run")
        f, e = C.c_int32(self.frm_cnt), C.c_int32(self.ell)
        L.lib().ds_mcra_advance(int(self.L), T, C.byref(f), C.byref(e))
        self.frm_cnt, self.ell = f.value, e.value
        return lam, pout

    @staticmethod
    def _power(Y):
        t = L.require_cuda()
        if isinstance(Y, t.Tensor):
            Yd = Y.to("cuda")
            if Yd.dtype == t.complex128:
                return L.spectral_power(Yd, via_abs=True)
            return Yd.to(t.float64)
        Y = np.asarray(Y)
        if Y.dtype == 'complex':                 # complex128 only, like the reference (:29-30): np.abs(Y) ** 2
            return L.spectral_power(t.as_tensor(np.ascontiguousarray(Y)).to("cuda"), via_abs=True)
        return t.as_tensor(np.ascontiguousarray(Y, dtype=np.float64)).to("cuda")

    def estimation(self, Y):
        """One frame: Y [K] power (or complex128) spectrum -> lambda_d [K]; 2-D input -> column 0
        (mcra.py:32-33).  Batched extension: pass ``estimation_batch``."""
        t = L.require_cuda()
        as_torch = isinstance(Y, t.Tensor)
        Yd = self._power(Y)
        if Yd.dim() > 1:
            Yd = Yd[:, 0]
        lam, _ = self._run(Yd.reshape(1, 1, -1).contiguous())
        return lam[0, 0] if as_torch else lam[0, 0].cpu().numpy()

    def estimation_frames(self, Y, return_p=False):
        """Extension: Y [T, K] (or [S, T, K]) -> lambda_d of every frame (and p)."""
        t = L.require_cuda()
        as_torch = isinstance(Y, t.Tensor)
        Yd = self._power(Y)
        batched = Yd.dim() == 3
        if not batched:
            Yd = Yd[None]
        lam, p = self._run(Yd.contiguous(), want_p=return_p)
        if not batched:
            lam = lam[0]
            p = p[0] if p is not None else None
        if not as_torch:
            lam = lam.cpu().numpy()
            p = p.cpu().numpy() if p is not None else None
        return (lam, p) if return_p else lam
