"""STFT / ISTFT / streaming Transform -- drop-in for
``DistantSpeech/transform/transform.py`` (stft :10, istft :237, Transform :407).

Same signatures, argument meaning and error behaviour as the reference for
1-D / 2-D inputs; the only extension is an optional leading batch axis of
independent streams.  Inputs may be NumPy arrays (results come back as NumPy,
like the reference) or CUDA torch tensors (results stay on the device).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib as L


def _sqrt_hann(n_fft: int) -> np.ndarray:
    # sqrt(get_window('hann', n_fft, fftbins=True))   (transform.py:418-419)
    n = np.arange(n_fft)
    return np.sqrt(0.5 - 0.5 * np.cos(2.0 * np.pi * n / n_fft))


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _check_window(window):
    # reference quirk 1: the string / None branches are dead (transform.py:189-193)
    if isinstance(window, str):
        raise AttributeError("window must be an ndarray: the reference's string branch is dead code "
                             "(transform.py:189-196, 'str' object has no attribute 'shape')")
    if window is None:
        raise ValueError("window must be an ndarray (transform.py:191-193 fails for window=None)")
    return np.asarray(window, dtype=np.float64)


def stft_device(x_dev, n_fft, hop, window_dev, mode, history=None, fft_fp64=False, out_c128=False):
    """x_dev [S, C, N] float32 CUDA -> X [S, T, C, K] complex64 (or complex128) CUDA."""
    t = L.require_cuda()
    L.ensure_init()
    S, Cn, N = x_dev.shape
    p = L.StftParams(n_fft, hop, S, Cn, N, mode, int(fft_fp64), int(out_c128))
    T = L.lib().ds_stft_num_frames(C.byref(p))
    if T < 0:
        raise L.DsError("bad stft parameters")
    K = n_fft // 2 + 1
    X = t.empty((S, T, Cn, K), dtype=t.complex128 if out_c128 else t.complex64, device="cuda")
    L.check(L.lib().ds_stft_run(C.byref(p), L.ptr(window_dev), L.ptr(history), L.ptr(x_dev), L.ptr(X), L.stream_ptr()),
            "ds_stft_run")
    return X


def istft_device(Y_dev, n_fft, hop, window_dev, mode, tail=None, scale=1.0, fft_fp64=False):
    """Y_dev [S, T, C, K] complex64/complex128 CUDA -> y [S, C, n_out] float32 CUDA."""
    t = L.require_cuda()
    L.ensure_init()
    S, T, Cn, K = Y_dev.shape
    assert K == n_fft // 2 + 1
    in_c128 = Y_dev.dtype == t.complex128
    n_out = hop * T if mode == L.DS_STFT_STREAMING else n_fft + hop * (T - 1)
    y = t.empty((S, Cn, n_out), dtype=t.float32, device="cuda")
    p = L.IstftParams(n_fft, hop, S, Cn, T, mode, int(fft_fp64), int(in_c128), float(scale))
    L.check(L.lib().ds_istft_run(C.byref(p), L.ptr(window_dev), L.ptr(tail), L.ptr(Y_dev), L.ptr(y), L.stream_ptr()),
            "ds_istft_run")
    return y


def stft(y, n_fft=2048, hop_length=None, win_length=None, window="hann", center=True,
         dtype=np.complex64, pad_mode="reflect", precision=None):
    """Short-time Fourier transform (reference transform.py:10-221).

    y [n] (or [S, n]) real -> D [1 + n_fft/2, T] (or [S, K, T]) of ``dtype``.
    ``precision``: "fp32" (default for complex64) or "fp64" FFT arithmetic.
    """
    if win_length is None:
        win_length = n_fft
    if hop_length is None:
        hop_length = int(win_length // 4)
    w = _check_window(window)
    t = L.require_cuda()
    if center and pad_mode != "reflect":
        raise L.DsError("only pad_mode='reflect' is implemented on the device")
    as_torch = _is_torch(y)
    yd = L.to_device(y, t.float32)
    batched = yd.dim() == 2
    if not batched:
        yd = yd[None]
    if yd.dim() != 2:
        raise ValueError("stft expects a 1-D signal (or [streams, n])")
    want128 = np.dtype(dtype) == np.complex128
    fp64 = (precision == "fp64") or (precision is None and want128)
    X = stft_device(yd[:, None, :].contiguous(), n_fft, hop_length, L.device_window(w, n_fft),
                    L.DS_STFT_CENTER if center else L.DS_STFT_PLAIN, fft_fp64=fp64, out_c128=want128)
    D = X[:, :, 0, :].permute(0, 2, 1)            # [S, K, T]
    if not batched:
        D = D[0]
    if as_torch:
        return D
    return np.asfortranarray(D.cpu().numpy().astype(dtype, copy=False))


def istft(stft_matrix, hop_length=None, win_length=None, window="hann", center=True,
          dtype=np.float32, length=None, precision=None):
    """Inverse STFT (reference transform.py:237-404): float32 overlap-add, no
    window-sum normalisation.  stft_matrix [K, T] (or [S, K, T]) -> y [n]."""
    t = L.require_cuda()
    as_torch = _is_torch(stft_matrix)
    if as_torch:
        Yd = stft_matrix.to("cuda")
    else:
        Yd = t.as_tensor(np.ascontiguousarray(stft_matrix)).to("cuda")
    batched = Yd.dim() == 3
    if not batched:
        Yd = Yd[None]
    n_fft = 2 * (Yd.shape[1] - 1)
    if win_length is None:
        win_length = n_fft
    if hop_length is None:
        hop_length = int(win_length // 4)
    w = _check_window(window)
    n_frames = Yd.shape[2]
    if length:
        padded = length + int(n_fft) if center else length
        n_frames = min(n_frames, int(np.ceil(padded / hop_length)))
    if Yd.dtype not in (t.complex64, t.complex128):
        Yd = Yd.to(t.complex64)
    fp64 = (precision == "fp64") or (precision is None and Yd.dtype == t.complex128)
    Yl = Yd[:, :, :n_frames].permute(0, 2, 1)[:, :, None, :].contiguous()     # [S, T, 1, K]
    y = istft_device(Yl, n_fft, hop_length, L.device_window(w, n_fft), L.DS_STFT_PLAIN, fft_fp64=fp64)[:, 0, :]
    if length is None:
        if center:
            y = y[:, n_fft // 2: y.shape[1] - n_fft // 2]
    else:
        start = n_fft // 2 if center else 0
        y = y[:, start:]
        if y.shape[1] > length:
            y = y[:, :length]
        elif y.shape[1] < length:
            y = t.nn.functional.pad(y, (0, length - y.shape[1]))
    if not batched:
        y = y[0]
    if as_torch:
        return y
    return y.cpu().numpy().astype(dtype, copy=False)


class Transform(object):
    """Streaming multichannel STFT/ISTFT (reference transform.py:407-496).

    ``stft(x[N, M] | [N])`` -> ``[K, T, M]`` complex128; ``istft(Y[K] | [K, C] |
    [K, T, C])`` -> ``[hop*T(, C)]`` float64 (float32-rounded values, squeezed).
    With a leading batch axis: ``x[S, N, M]`` -> ``[S, K, T, M]`` and
    ``Y[S, K, T, C]`` -> ``[S, hop*T, C]``.  The chunk length must be a multiple
    of ``hop_length`` for a gap-free stream (reference quirk 4, :441).  ``window`` must have exactly
    ``n_fft`` taps (every reference call site passes none or a full-length window); anything else raises ValueError.
    """

    def __init__(self, channel=1, n_fft=256, hop_length=128, window=None, precision="fp32"):
        self.channel = channel
        self.n_fft = n_fft
        self.frame_length = n_fft
        self.hop_length = hop_length
        self.first_frame = 1
        self.window = _sqrt_hann(n_fft) if window is None else np.asarray(window, dtype=np.float64)
        if self.window.ndim != 1 or self.window.shape[0] != n_fft:
            # the device history / tail hold n_fft - hop samples per channel; a shorter window would make the
            # reference keep win_len - hop samples instead (transform.py:427,438) -- not supported, say so loudly
            raise ValueError("Transform: window must have n_fft = %d taps, got shape %s" % (n_fft, self.window.shape))
        self.half_bin = int(n_fft / 2 + 1)
        self.win_len = self.window.shape[0]
        self.overlap = self.win_len - hop_length
        self.W0 = np.sum(self.window ** 2)
        self.precision = precision
        self._hist = None     # [S, M, overlap] float32 CUDA
        self._tail = None     # [S, C, overlap] float32 CUDA
        self._S = None

    # reference attributes (views of the device state)
    @property
    def previous_input(self):
        if self._hist is None:
            return np.zeros((self.overlap, self.channel))
        h = self._hist.permute(0, 2, 1).double().cpu().numpy()
        return h[0] if self._S == 1 else h

    @property
    def previous_output(self):
        if self._tail is None:
            return np.zeros((self.overlap, self.channel))
        h = self._tail.permute(0, 2, 1).double().cpu().numpy()
        return h[0] if self._S == 1 else h

    def _state(self, S):
        t = L.require_cuda()
        if self._hist is None or self._S != S:
            self._S = S
            self._hist = t.zeros((S, self.channel, max(self.overlap, 1)), dtype=t.float32, device="cuda")
            self._tail = t.zeros((S, self.channel, max(self.overlap, 1)), dtype=t.float32, device="cuda")

    def stft(self, x):
        t = L.require_cuda()
        as_torch = _is_torch(x)
        xd = L.to_device(x, t.float32)
        if xd.dim() == 1:
            xd = xd[:, None]
        batched = xd.dim() == 3
        if not batched:
            xd = xd[None]
        S, N, M = xd.shape
        if M != self.channel:
            raise ValueError("input has %d channels, Transform was built for %d" % (M, self.channel))
        self._state(S)
        if self.overlap + N < self.n_fft:
            raise ValueError("chunk shorter than one hop: the reference fails in util.frame here")
        xs = xd.permute(0, 2, 1).contiguous()                         # [S, M, N]
        X = stft_device(xs, self.n_fft, self.hop_length, L.device_window(self.window, self.n_fft),
                        L.DS_STFT_STREAMING, history=self._hist, fft_fp64=(self.precision == "fp64"))
        Y = X.permute(0, 3, 1, 2)                                     # [S, K, T, M]
        if not batched:
            Y = Y[0]
        if as_torch:
            return Y
        return Y.cpu().numpy().astype(np.complex128)

    def istft(self, Y):
        t = L.require_cuda()
        as_torch = _is_torch(Y)
        Yd = Y.to("cuda") if as_torch else t.as_tensor(np.ascontiguousarray(Y)).to("cuda")
        batched = Yd.dim() == 4
        if not batched:
            if Yd.dim() == 1:                      # single frame, single channel (:461)
                Yd = Yd[:, None, None]
            if Yd.dim() == 2:                      # single frame x channels (:463)
                Yd = Yd[:, None, :]
            Yd = Yd[None]
        S, K, T, Cn = Yd.shape
        assert Cn <= self.channel, 'n_channels:{} != self.channel:{}'.format(Cn, self.channel)
        self._state(S)
        if Yd.dtype not in (t.complex64, t.complex128):
            Yd = Yd.to(t.complex64)
        Yl = Yd.permute(0, 2, 3, 1).contiguous()                      # [S, T, C, K]
        tail = self._tail[:, :Cn, :].contiguous() if Cn != self.channel else self._tail
        y = istft_device(Yl, self.n_fft, self.hop_length, L.device_window(self.window, self.n_fft),
                         L.DS_STFT_STREAMING, tail=tail, scale=self.hop_length / self.W0,
                         fft_fp64=(self.precision == "fp64"))
        if Cn != self.channel:
            self._tail[:, :Cn, :] = tail
        out = y.permute(0, 2, 1)                                      # [S, hop*T, C]
        if not batched:
            out = out[0]
        if as_torch:
            return out.squeeze()
        return out.double().cpu().numpy().squeeze()

    def magphase(self, D, power=1):
        mag = np.abs(D)
        mag **= power
        phase = np.exp(1.0j * np.angle(D))
        return mag, phase

    def analysis(self, x):
        return self.stft(x)

    def synthesis(self, Y):
        return self.istft(Y)
