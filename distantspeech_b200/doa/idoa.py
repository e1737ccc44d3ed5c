"""Spatial speech-presence probability from the instantaneous-DOA feature -- drop-in for
``DistantSpeech/doa/idoa.py`` (Idoa :20, estimate :92, process :175).

``estimate(X[K, T, M], theta=None)`` returns ``p[K, T, n_theta]`` like the reference;
``process(x[N, M], theta=None, default_direction=90)`` returns the signal of microphone 0 scaled per
frame by ``max(mean(p[64:128, :, direction]), 0.01)``.  Directions never interact in the reference
(beta_n, the statistics and p are all per direction), so ``process`` tracks only the direction it
uses -- the other 359 columns of the reference's map do not reach its output.  Extension: a leading
stream axis on ``x``.  All arithmetic runs in csrc/idoa.cu; the recursive state (RTF estimate per bin,
four statistics per bin and direction) lives on the device between calls, like the reference's
attributes.  Not built: ``pre_emphsis=True`` and the realtime PyAudio wrapper (``IdoaRealtime``).
"""
import numpy as np

from .. import _lib as L
from ..beamformer.MicArray import MicArray
from ..transform.transform import Transform, stft_device, istft_device


class Idoa(object):
    def __init__(self, mic_array: MicArray) -> None:
        self.mic_array = mic_array
        self.transform = Transform(channel=mic_array.M, n_fft=mic_array.n_fft, hop_length=int(mic_array.n_fft / 2))
        self.half_bin = self.transform.half_bin
        self.idoa_dim = mic_array.M - 1
        self.n_theta = 180 if mic_array.arrayType == "linear" else 360          # idoa.py:41-44
        K = self.half_bin
        self.Psi = np.zeros((K, self.idoa_dim, self.n_theta), dtype=complex)   # predefined free-field RTF, :73-76
        for th in range(self.n_theta):
            sv = mic_array.steering_vector(look_direction=th)
            self.Psi[:, :, th] = sv[:, 1:] / sv[:, 0:1]
        self.beta = 7.6
        self._psi_dev = None
        self._rtf_state = None       # [S][1 + 2(M-1)][K]
        self._spp_state = None       # [S][n_slots][4][K]
        self._slots = None           # direction index of every state slot
        self._S = None

    # ---- device state ------------------------------------------------------------------
    def _ensure(self, S, slots):
        t = L.require_cuda()
        L.ensure_init()
        K, M = self.half_bin, self.mic_array.M
        if self._psi_dev is None:
            self._psi_dev = t.as_tensor(np.ascontiguousarray(self.Psi.transpose(2, 1, 0))).to("cuda")   # [n_theta, M-1, K]
        slots = [int(v) for v in slots]
        if self._rtf_state is None or self._S != S:
            self._rtf_state = t.zeros(L.lib().ds_idoa_rtf_state_bytes(S, M, K), dtype=t.uint8, device="cuda")
            self._spp_state = None
            self._S = S
        if self._spp_state is None or self._slots != slots:
            if self._spp_state is not None:
                raise ValueError("this Idoa object already tracks directions %s; use a fresh object for %s"
                                 % (self._slots[:4], slots[:4]))
            self._spp_state = t.zeros(L.lib().ds_idoa_spp_state_bytes(S, len(slots), K), dtype=t.uint8, device="cuda")
            self._slots = slots
            self._slots_dev = t.as_tensor(np.asarray(slots, dtype=np.int32)).to("cuda")

    def _attr(self, idx):
        """mu_Delta / mu_Delta_h0 / var_Delta_h0 / p as [K, n_slots] (first stream), like the reference's attributes."""
        if self._spp_state is None:
            return None
        t = L.require_cuda()
        v = self._spp_state.view(t.float64).reshape(self._S, len(self._slots), 4, self.half_bin)[0, :, idx, :]
        return (v + (0.1 if idx == 2 else 0.0)).t().cpu().numpy()

    mu_Delta = property(lambda self: self._attr(0))
    mu_Delta_h0 = property(lambda self: self._attr(1))
    var_Delta_h0 = property(lambda self: self._attr(2))
    p = property(lambda self: self._attr(3))

    def _run(self, Xd, slots, only_theta, want_p, want_Y):
        """Xd [S, T, M, K] complex64/complex128 CUDA -> (p [S, T, n_slots, K] or None, Y [S, T, K] c128 or None)."""
        t = L.require_cuda()
        S, T, M, K = Xd.shape
        if M != self.mic_array.M or K != self.half_bin:
            raise ValueError("expected a spectrum [.., %d mics, %d bins]" % (self.mic_array.M, self.half_bin))
        self._ensure(S, slots)
        c128 = int(Xd.dtype == t.complex128)
        B = t.empty((S, T, 2 * (M - 1) + 1, K), dtype=t.float64, device="cuda")
        L.check(L.lib().ds_idoa_rtf_run(S, T, M, K, 0.02, L.ptr(Xd), c128, L.ptr(self._rtf_state), L.ptr(B), L.stream_ptr()),
                "ds_idoa_rtf_run")
        p = t.empty((S, T, len(self._slots), K), dtype=t.float64, device="cuda") if want_p else None
        Y = t.empty((S, T, K), dtype=t.complex128, device="cuda") if want_Y else None
        L.check(L.lib().ds_idoa_spp_run(S, T, M, K, len(self._slots), L.ptr(self._slots_dev), int(only_theta),
                                        L.ptr(self._psi_dev), L.ptr(B), L.ptr(self._spp_state), L.ptr(p), L.ptr(Xd), c128,
                                        L.ptr(Y), L.stream_ptr()), "ds_idoa_spp_run")
        return p, Y

    # ---- reference API -------------------------------------------------------------------
    def estimate(self, X, theta=None):
        """X [half_bin, n_frames, channels] complex -> p [half_bin, n_frames, n_theta]  (idoa.py:92-173).
        With ``theta`` given only that column sees the data (the others evolve with Delta = 0, as in the reference)."""
        t = L.require_cuda()
        as_torch = isinstance(X, t.Tensor)
        Xd = X.to("cuda") if as_torch else t.as_tensor(np.ascontiguousarray(X)).to("cuda")
        if Xd.dtype not in (t.complex64, t.complex128):
            Xd = Xd.to(t.complex128)
        assert Xd.shape[0] == self.half_bin
        Xl = Xd.permute(1, 2, 0).contiguous()[None]                               # [1, T, M, K]
        p, _ = self._run(Xl, range(self.n_theta), -1 if theta is None else int(theta), True, False)
        p = p[0].permute(2, 0, 1)                                                 # [K, T, n_theta]
        return p if as_torch else p.cpu().numpy()

    def process(self, x, theta=None, default_direction=90, pre_emphsis=False):
        """x [samples, channels] (or [S, samples, channels]) -> enhanced signal [samples] (or [S, samples])
        (idoa.py:175-209)."""
        if pre_emphsis:
            raise NotImplementedError("pre_emphsis=True is not built")
        t = L.require_cuda()
        as_torch = isinstance(x, t.Tensor)
        xd = L.to_device(x, t.float32)
        batched = xd.dim() == 3
        if not batched:
            xd = xd[None]
        S, N, M = xd.shape
        tf = self.transform
        if M != tf.channel:
            raise ValueError("input has %d channels, Idoa was built for %d" % (M, tf.channel))
        tf._state(S)
        wdev = L.device_window(tf.window, tf.n_fft)
        X = stft_device(xd.permute(0, 2, 1).contiguous(), tf.n_fft, tf.hop_length, wdev, L.DS_STFT_STREAMING,
                        history=tf._hist)                                          # [S, T, M, K]
        d = int(theta) if theta is not None else int(default_direction)
        _, Y = self._run(X, [d], -1, False, True)
        tail = tf._tail[:, :1, :].contiguous()                                     # the reference reuses its M-channel Transform
        y = istft_device(Y[:, :, None, :], tf.n_fft, tf.hop_length, wdev, L.DS_STFT_STREAMING, tail=tail,
                         scale=tf.hop_length / tf.W0)
        tf._tail[:, :1, :] = tail
        y = y[:, 0, :]
        if not batched:
            y = y[0]
        return y if as_torch else y.double().cpu().numpy()
