"""``DistantSpeech/adaptivefilter/SubbandRLS.py`` (SubbandRLS :12, update :44-71): per-bin RLS filter with
``filter_len`` frame taps; recursion in ``ds_subband_rls_run``, analysis / synthesis by the device STFT / ISTFT."""
import numpy as np

from .. import _lib as L
from ..transform.transform import _sqrt_hann, stft_device, istft_device


class SubbandRLS(object):
    def __init__(self, filter_len=2, num_bands=512, forgetting_factor=0.998, mu=0.5, normalization=True, alpha=0.9, m=2,
                 hop_length=None, input_td=False):
        if not 1 <= filter_len <= 4:
            raise ValueError("SubbandRLS on the device is compiled for filter_len 1..4")
        self.filter_len, self.num_bands = filter_len, num_bands
        self.half_band = int(num_bands / 2) + 1
        self.mu_value = mu
        self.forgetting_factor = forgetting_factor
        self.forgetting_factor_inv = 1.0 / forgetting_factor
        self.hop_length = int(num_bands / 2) if hop_length is None else hop_length
        self.window = _sqrt_hann(num_bands)
        self.return_td = False
        self._state = None

    def _ensure(self):
        t = L.require_cuda()
        L.ensure_init()
        if self._state is None:
            Lf, K = self.filter_len, self.half_band
            st = t.zeros((1, 4 * Lf + 2 * Lf * Lf, K), dtype=t.float64, device="cuda")
            for i in range(Lf):
                st[:, 4 * Lf + i * Lf + i, :] = 1.0 / 1e-3                  # P = I / 1e-3 (:38-40)
            self._state = st
            ov = self.num_bands - self.hop_length
            self._hist_x = t.zeros((1, 1, ov), dtype=t.float32, device="cuda")
            self._hist_d = t.zeros((1, 1, ov), dtype=t.float32, device="cuda")
            self._tail = t.zeros((1, 1, ov), dtype=t.float32, device="cuda")

    @property
    def W(self):
        Lf = self.filter_len
        if self._state is None:
            return np.zeros((self.half_band, Lf), dtype=complex)
        b = self._state[0, :2 * Lf, :].cpu().numpy()
        return (b[:Lf] + 1j * b[Lf:]).T

    @property
    def P(self):
        Lf = self.filter_len
        if self._state is None:
            return np.tile(np.eye(Lf, dtype=complex) / 1e-3, (self.half_band, 1, 1))
        b = self._state[0, 4 * Lf:, :].cpu().numpy()
        return (b[:Lf * Lf] + 1j * b[Lf * Lf:]).T.reshape(self.half_band, Lf, Lf)

    def update(self, x_n, d_n, alpha=1e-4, p=None):
        """x_n, d_n [samples] float blocks (a multiple of hop_length) -> (err block, W); ``alpha`` / ``p`` are unused
        by the reference's RLS update too."""
        t = L.require_cuda()
        self._ensure()
        win = L.device_window(self.window, self.num_bands)
        as_dev = (lambda v: t.as_tensor(np.ascontiguousarray(np.asarray(v, dtype=np.float32).reshape(1, 1, -1))).to("cuda"))
        X = stft_device(as_dev(x_n), self.num_bands, self.hop_length, win, L.DS_STFT_STREAMING, history=self._hist_x)
        D = stft_device(as_dev(d_n), self.num_bands, self.hop_length, win, L.DS_STFT_STREAMING, history=self._hist_d)
        T = X.shape[1]
        Err = t.empty((1, T, 1, self.half_band), dtype=t.complex128, device="cuda")
        L.check(L.lib().ds_subband_rls_run(1, self.half_band, T, self.filter_len, float(self.mu_value), float(self.forgetting_factor),
                                           L.ptr(self._state), L.ptr(X), L.ptr(D), L.ptr(Err), L.stream_ptr()),
                "ds_subband_rls_run")
        y = istft_device(Err, self.num_bands, self.hop_length, win, L.DS_STFT_STREAMING, tail=self._tail,
                         scale=self.hop_length / float(np.sum(self.window ** 2)))
        self.return_td = True
        return y[0, 0].double().cpu().numpy(), self.W
