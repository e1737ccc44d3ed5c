"""Shared core of the STFT-domain NLMS filters -- ``DistantSpeech/adaptivefilter/SubbandAF.py`` (SubbandAF :12),
``SubbandLMS.py`` (update :28-84) and ``SubbandLmsMc.py`` (update :144-191).  The per-bin recursion runs in
``ds_subband_nlms_run``; the analysis / synthesis transforms around it are the device STFT / ISTFT."""
import ctypes as C

import numpy as np

from .. import _lib as L
from ..transform.transform import _sqrt_hann, stft_device, istft_device


class SubbandAF(object):
    def __init__(self, filter_len=2, num_bands=512, mu=0.1, normalization=True, alpha=0.9, m=2, hop_length=None,
                 input_td=False, channel=1):
        self.filter_len = filter_len
        self.num_bands = num_bands
        self.half_band = int(num_bands / 2) + 1
        self.M = channel
        self.mu_value = mu
        self.mu = np.ones((self.half_band, 1)) * mu
        self.norm = normalization
        self.alpha = alpha
        self.hop_length = int(num_bands / 2) if hop_length is None else hop_length
        self.window = _sqrt_hann(num_bands)
        self.return_td = False
        self._state = None
        self._S = None

    # ---- device state: [S][F=1][NE][K] float64 ---------------------------------------------
    def _params(self, S, T, one_minus_p=0, eps=1e-4):
        p = L.SubbandNlmsParams(self.half_band, S, 1, T, self.M, self.filter_len, int(one_minus_p), int(not self.norm),
                                float(self.mu_value), float(self.alpha), float(eps))
        return p

    def _ensure(self, S):
        t = L.require_cuda()
        L.ensure_init()
        if self._state is None or self._S != S:
            nb = L.lib().ds_subband_nlms_state_bytes(C.byref(self._params(S, 1)))
            self._state = t.zeros(nb, dtype=t.uint8, device="cuda")
            ov = self.num_bands - self.hop_length
            self._hist_x = t.zeros((S, self.M, ov), dtype=t.float32, device="cuda")
            self._hist_d = t.zeros((S, 1, ov), dtype=t.float32, device="cuda")
            self._tail = t.zeros((S, 1, ov), dtype=t.float32, device="cuda")
            self._S = S

    def _blob(self):
        t = L.require_cuda()
        return self._state.view(t.float64).view(self._S, -1, self.half_band)

    @property
    def W(self):
        """Filter weights [half_band, filter_len] (one channel) or [half_band, filter_len, channel]."""
        LC = self.filter_len * self.M
        if self._state is None:
            w = np.zeros((1, LC, self.half_band), dtype=complex)
        else:
            b = self._blob()[:, :2 * LC, :].cpu().numpy()
            w = b[:, :LC] + 1j * b[:, LC:]
        w = w.transpose(0, 2, 1).reshape(w.shape[0], self.half_band, self.filter_len, self.M)
        w = w[0] if w.shape[0] == 1 else w
        return w[..., 0] if self.M == 1 else w

    @property
    def P(self):
        if self._state is None:
            return np.zeros(self.half_band)
        v = self._blob()[:, 4 * self.filter_len * self.M, :].cpu().numpy()
        return v[0] if v.shape[0] == 1 else v

    def _run_spec(self, X, D, p, one_minus_p=False, eps=1e-4):
        """X [S, T, C, K] c64, D [S, T, 1, K] c64, p [S, T, K] f64 or None (CUDA) -> Err [S, T, 1, K] c128."""
        t = L.require_cuda()
        S, T, Cn, K = X.shape
        self._ensure(S)
        Err = t.empty((S, T, 1, K), dtype=t.complex128, device="cuda")
        prm = self._params(S, T, one_minus_p, eps)
        L.check(L.lib().ds_subband_nlms_run(C.byref(prm), L.ptr(self._state), L.ptr(X.contiguous()), L.ptr(D.contiguous()),
                                            L.ptr(p), L.ptr(Err), L.stream_ptr()), "ds_subband_nlms_run")
        return Err

    def _update_td(self, x_n, d_n, alpha, p):
        """Time-domain blocks in, time-domain error block out (update_input_data :53-59 with float input)."""
        t = L.require_cuda()
        x = np.asarray(x_n, dtype=np.float64).reshape(len(x_n), -1)
        d = np.asarray(d_n, dtype=np.float64).reshape(-1, 1)
        if x.shape[1] != self.M:
            raise ValueError("expected %d input channel(s)" % self.M)
        self._ensure(1)
        win = L.device_window(self.window, self.num_bands)
        xd = t.as_tensor(np.ascontiguousarray(x.T, dtype=np.float32)).to("cuda")[None]
        dd = t.as_tensor(np.ascontiguousarray(d.T, dtype=np.float32)).to("cuda")[None]
        X = stft_device(xd, self.num_bands, self.hop_length, win, L.DS_STFT_STREAMING, history=self._hist_x)
        D = stft_device(dd, self.num_bands, self.hop_length, win, L.DS_STFT_STREAMING, history=self._hist_d)
        T = X.shape[1]
        pd = None
        if p is not None and not isinstance(p, float):
            pv = np.asarray(p, dtype=np.float64)
            pv = pv[:, 0] if pv.ndim == 2 else pv
            assert pv.shape[0] == self.half_band
            pd = t.as_tensor(np.array(np.broadcast_to(pv[None, None, :], (1, T, self.half_band)))).to("cuda")
        elif isinstance(p, float):
            pd = t.full((1, T, self.half_band), p, dtype=t.float64, device="cuda")
        Err = self._run_spec(X, D, pd, eps=alpha)
        y = istft_device(Err, self.num_bands, self.hop_length, win, L.DS_STFT_STREAMING, tail=self._tail,
                         scale=self.hop_length / float(np.sum(self.window ** 2)))
        self.return_td = True
        return y[0, 0].double().cpu().numpy(), self.W
