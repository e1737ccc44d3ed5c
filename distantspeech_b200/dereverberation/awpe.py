"""Adaptive (RLS) WPE dereverberation -- drop-in for ``DistantSpeech/dereverberation/awpe.py`` (Wpe :28, update :128).

Per frequency bin a multichannel linear predictor estimates the late reverberation of the current frame from
``filter_len`` frames that lie ``delay`` frames in the past; the prediction error of every channel is the dereverberated
spectrum.  The recursion (awpe.py:152-187) runs in ``ds_wpe_run``, analysis / synthesis are the device STFT / ISTFT.

The reference's ``update`` cannot run as shipped: it calls ``self.check_input_data``, which no class defines (:150), and
ends in ``return output, self.W`` with ``output`` unassigned (:188-191).  Here ``update(x_n)`` returns what those lines
evidently mean, ``(synthesis(err[:, 0]), W)``; ``process`` (extension) returns every channel and batches streams.  The
filter bank is the streaming sqrt-Hann ``Transform`` (the reference's ``Subband`` Nyquist(M) bank is outside the hot
path); the arithmetic is pinned against the reference's own ``update`` body
(oracle/ref_harness.make_wpe, tests/golden/wpe.npz).
"""
import numpy as np

from .. import _lib as L
from ..transform.transform import _sqrt_hann, stft_device, istft_device


class Wpe(object):
    def __init__(self, channels=2, filter_len=2, num_bands=512, forgetting_factor=0.998, delay=4, mu=0.5,
                 normalization=True, alpha=0.9, m=2, hop_length=None, input_td=False):
        if channels * filter_len > 16 or channels > 8:
            raise ValueError("Wpe on the device is compiled for channels <= 8 and channels * filter_len <= 16")
        self.channels, self.filter_len, self.num_bands = int(channels), int(filter_len), int(num_bands)
        self.half_band = int(num_bands / 2) + 1
        self.forgetting_factor = forgetting_factor
        self.forgetting_factor_inv = 1.0 / forgetting_factor
        self.D = int(delay)
        self.hop_length = int(num_bands / 2) if hop_length is None else int(hop_length)
        self.window = _sqrt_hann(self.num_bands)
        self.return_td = False
        self._state = None
        self._S = None

    # ---- device state: [S][NE][K] float64 ------------------------------------------------------------
    def _ensure(self, S):
        t = L.require_cuda()
        L.ensure_init()
        if self._state is None or self._S != S:
            C_, Lf, K = self.channels, self.filter_len, self.half_band
            CL = C_ * Lf
            nb = L.lib().ds_wpe_state_bytes(S, K, C_, Lf, self.D)
            st = t.zeros(nb // 8, dtype=t.float64, device="cuda").view(S, -1, K)
            for i in range(CL):
                st[:, 2 * C_ * CL + i * CL + i, :] = 1e-3                    # P = 1e-3 I (awpe.py:66-71)
            self._state = st
            ov = self.num_bands - self.hop_length
            self._hist = t.zeros((S, C_, ov), dtype=t.float32, device="cuda")
            self._tail = t.zeros((S, C_, ov), dtype=t.float32, device="cuda")
            self._S = S

    def _view(self, lo, n):
        v = self._state[:, lo:lo + n, :].cpu().numpy()
        return v

    @property
    def W(self):
        """[half_band, channels, channels * filter_len] complex (leading stream axis when batched)."""
        C_, CL, K = self.channels, self.channels * self.filter_len, self.half_band
        if self._state is None:
            return np.zeros((K, C_, CL), dtype=complex)
        w = (self._view(0, C_ * CL) + 1j * self._view(C_ * CL, C_ * CL)).transpose(0, 2, 1).reshape(-1, K, C_, CL)
        return w[0] if w.shape[0] == 1 else w

    @property
    def P(self):
        C_, CL, K = self.channels, self.channels * self.filter_len, self.half_band
        if self._state is None:
            return np.tile(np.eye(CL, dtype=complex) * 1e-3, (K, 1, 1))
        o = 2 * C_ * CL
        p = (self._view(o, CL * CL) + 1j * self._view(o + CL * CL, CL * CL)).transpose(0, 2, 1).reshape(-1, K, CL, CL)
        return p[0] if p.shape[0] == 1 else p

    @property
    def var(self):
        if self._state is None:
            return np.zeros((self.half_band, 1))
        v = self._state[:, -1, :].cpu().numpy()[..., None]
        return v[0] if v.shape[0] == 1 else v

    # ---- processing ------------------------------------------------------------------------------------
    def process_device(self, xs):
        """xs [S, C, N] float32 CUDA (N a multiple of hop_length) -> dereverberated [S, C, N] float32 CUDA."""
        t = L.require_cuda()
        S, C_, N = xs.shape
        if C_ != self.channels or N % self.hop_length != 0 or N < self.hop_length:
            raise ValueError("expected [S, %d, N] with N a positive multiple of hop_length=%d" % (self.channels, self.hop_length))
        self._ensure(S)
        win = L.device_window(self.window, self.num_bands)
        X = stft_device(xs.contiguous(), self.num_bands, self.hop_length, win, L.DS_STFT_STREAMING, history=self._hist)
        T = X.shape[1]
        Err = t.empty((S, T, C_, self.half_band), dtype=t.complex128, device="cuda")
        L.check(L.lib().ds_wpe_run(S, self.half_band, T, C_, self.filter_len, self.D, float(self.forgetting_factor), 0.98,
                                   L.ptr(self._state), L.ptr(X), 0, L.ptr(Err), L.stream_ptr()), "ds_wpe_run")
        return istft_device(Err, self.num_bands, self.hop_length, win, L.DS_STFT_STREAMING, tail=self._tail,
                            scale=self.hop_length / float(np.sum(self.window ** 2)))

    def process(self, x):
        """Extension: x [N, C] (or [S, N, C]) -> dereverberated y of the same shape (every channel's prediction error)."""
        t = L.require_cuda()
        as_torch = isinstance(x, t.Tensor)
        xd = L.to_device(x, t.float32)
        batched = xd.dim() == 3
        if not batched:
            xd = xd[None]
        y = self.process_device(xd.permute(0, 2, 1).contiguous()).permute(0, 2, 1)
        if not batched:
            y = y[0]
        return y if as_torch else y.double().cpu().numpy()

    def update(self, x_n, alpha=1e-4, p=None):
        """x_n [samples, ch] float block (a multiple of hop_length) -> (dereverberated block of channel 0, W)."""
        y = self.process(np.asarray(x_n, dtype=np.float64))
        self.return_td = True
        return y[:, 0], self.W
