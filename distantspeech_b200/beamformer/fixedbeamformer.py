"""Fixed delay-and-sum / superdirective beamformer -- drop-in for
``DistantSpeech/beamformer/fixedbeamformer.py`` (FixedBeamformer :96,
compute_weights :109, process_freframe :147, process :167).

``process`` runs the fused STFT -> weight -> ISTFT kernel (ds_fixedbf_run); the
spectrum never leaves the SM.  Batched use: ``process(x[S, N, M])`` -> ``[S, N]``
and ``process_multibeam`` evaluates several look directions in one pass.
"""
import ctypes as C

import numpy as np

from .. import _lib as L
from .beamformer import beamformer, apply_weights


class FixedBeamformer(beamformer):
    def __init__(self, MicArray, frameLen=256, hop=None, nfft=None, c=343, fs=16000, r=0.032):
        beamformer.__init__(self, MicArray, frame_len=frameLen, hop=hop, nfft=nfft, c=c, fs=fs)
        self.angle = (197, 0)
        self.AlgorithmList = ["src", "DS", "MVDR"]
        self.AlgorithmIndex = 0
        self.W = self.compute_weights(look_angle=self.angle)
        self._state = None
        self._state_key = None

    def compute_weights(self, look_angle=[90, 0], weightType="SD", diag_value=1e-3):
        """DS: a0 / M;  SD: mvdr(a0, inv(Gamma + diag I)), default "SD" (:109-145).
        Deliberate deviation: the reference builds Gamma with gen_noise_msc's default
        nfft=256 (:140) and therefore crashes for any other nfft; here Gamma always
        has this beamformer's nfft (identical result at nfft=256)."""
        return beamformer.compute_weights(self, look_angle=look_angle, weightType=weightType, diag_value=diag_value)

    def process_freframe(self, X_n):
        """Yf[k] = sum_m conj(W[k, m]) X_n[k, m]   (:147-165)."""
        return apply_weights(self.W, X_n)

    def _run(self, x, W):
        """x [S, N, M] (numpy / torch) , W [B, K, M] complex -> y [S, B, N] float32 CUDA."""
        t = L.require_cuda()
        L.ensure_init()
        xd = L.to_device(x, t.float32)
        S, N, M = xd.shape
        if N % self.hop != 0:
            raise AssertionError('output:{}, x:{}'.format((N // self.hop) * self.hop, tuple(xd.shape)))
        xs = xd.permute(0, 2, 1).contiguous()                                   # [S, M, N]
        Wd = L.to_device(np.asarray(W, dtype=np.complex64), t.complex64)
        B = Wd.shape[0]
        tf = self.transform
        p = L.FixedBfParams(self.nfft, self.hop, S, M, N, B, 0, 0, float(self.hop / tf.W0))
        key = (S, M, B, self.nfft, self.hop)
        if self._state is None or self._state_key != key:
            nbytes = L.lib().ds_fixedbf_state_bytes(C.byref(p))
            self._state = t.zeros(nbytes, dtype=t.uint8, device="cuda")
            self._state_key = key
        y = t.empty((S, B, N), dtype=t.float32, device="cuda")
        L.check(L.lib().ds_fixedbf_run(C.byref(p), L.ptr(L.device_window(tf.window, self.nfft)), L.ptr(Wd),
                                       L.ptr(self._state), L.ptr(xs), L.ptr(y), L.stream_ptr()), "ds_fixedbf_run")
        return y

    def reset(self):
        """Drop the streaming state (input history / output tail)."""
        self._state = None
        self._tc_key = None

    def process(self, x, angle=(0, 0), weightType=None):
        """x [samples, channel] (or [S, samples, channel]) -> [samples] (or [S, samples]).

        Like the reference (:183-188) the weights are recomputed from ``angle`` on every
        call with compute_weights' default type ("SD"): ``list(angle) != self.angle`` is
        always true there.  ``weightType`` is an extension to pick "DS"."""
        t = L.require_cuda()
        assert x.shape[-1] >= 2
        angle = list(angle)
        if weightType is None:
            self.W = self.compute_weights(look_angle=angle)
        else:
            self.W = self.compute_weights(look_angle=angle, weightType=weightType)
        as_torch = isinstance(x, t.Tensor)
        batched = len(x.shape) == 3
        xin = x if batched else x[None]
        y = self._run(xin, self.W[None])[:, 0, :]
        if not batched:
            y = y[0]
        if as_torch:
            return y
        return y.double().cpu().numpy().squeeze()

    def _run_tensor(self, x, W):
        """The multi-beam path on the tensor cores: device STFT -> ds_multibeam_tc_run (per-bin [beams x 6M] . [6M x frames]
        GEMMs on tcgen05, tf32 head + tail operands) -> per-beam ISTFT of the frames-innermost spectrum.
        x [S, N, M], W [B, K, M] complex -> y [S, B, N] float32 CUDA.  Streaming state (input history, per-beam output
        tail) is kept between calls like the fused kernel's."""
        from ..transform.transform import stft_device
        t = L.require_cuda()
        L.ensure_init()
        xd = L.to_device(x, t.float32)
        S, N, M = xd.shape
        if N % self.hop != 0:
            raise AssertionError('output:{}, x:{}'.format((N // self.hop) * self.hop, tuple(xd.shape)))
        xs = xd.permute(0, 2, 1).contiguous()                                   # [S, M, N]
        Wd = L.to_device(np.asarray(W, dtype=np.complex64), t.complex64).contiguous()
        B, K = Wd.shape[0], self.half_bin
        tf = self.transform
        ov = self.nfft - self.hop
        key = (S, M, B, self.nfft, self.hop)
        if getattr(self, "_tc_key", None) != key:
            self._tc_key = key
            self._tc_hist = t.zeros((S, M, max(ov, 1)), dtype=t.float32, device="cuda")
            self._tc_tail = t.zeros((S, B, max(ov, 1)), dtype=t.float32, device="cuda")
        win = L.device_window(tf.window, self.nfft)
        X = stft_device(xs, self.nfft, self.hop, win, L.DS_STFT_STREAMING, history=self._tc_hist)     # [S, T, M, K] c64
        T = X.shape[1]
        pitch, wsb, yb = C.c_int32(), C.c_size_t(), C.c_size_t()
        L.check(L.lib().ds_multibeam_tc_layout(S, T, M, K, B, C.byref(pitch), C.byref(wsb), C.byref(yb)), "ds_multibeam_tc_layout")
        ws = t.empty(wsb.value, dtype=t.uint8, device="cuda")
        Y = t.empty((S, B, K, pitch.value), dtype=t.complex64, device="cuda")
        L.check(L.lib().ds_multibeam_tc_run(S, T, M, K, B, L.ptr(Wd), L.ptr(X), L.ptr(ws), L.ptr(Y), L.stream_ptr()),
                "ds_multibeam_tc_run")
        ip = L.IstftParams(self.nfft, self.hop, S, B, T, L.DS_STFT_STREAMING, 0, 0, float(self.hop / tf.W0))
        y = t.empty((S, B, T * self.hop), dtype=t.float32, device="cuda")
        L.check(L.lib().ds_istft_frames_inner_run(C.byref(ip), L.ptr(win), L.ptr(self._tc_tail), L.ptr(Y), pitch.value, L.ptr(y),
                                                  L.stream_ptr()), "ds_istft_frames_inner_run")
        return y

    def process_multibeam(self, x, angles, weightType="SD", engine="auto"):
        """Extension: x [S, N, M], ``angles`` list of (az, el) -> [S, B, N], all beams in one pass.
        ``engine``: "tensor" = per-bin GEMMs on the tensor cores (tcgen05; 4, 8 or 16 microphones), "simt" = the fused
        CUDA-core kernels, "auto" = tensor cores from 16 beams up (below that the fused kernel, whose spectrum never
        leaves the SM, is faster)."""
        t = L.require_cuda()
        wkey = (tuple(tuple(float(v) for v in a) for a in angles), weightType)
        if getattr(self, "_mb_wkey", None) != wkey:                          # one-off design per set of look directions (host)
            self._mb_W = np.stack([self.compute_weights(look_angle=list(a), weightType=weightType) for a in angles])
            self._mb_wkey = wkey
        W = self._mb_W
        as_torch = isinstance(x, t.Tensor)
        if engine not in ("auto", "tensor", "simt"):
            raise ValueError("engine must be 'auto', 'tensor' or 'simt'")
        use_tc = engine == "tensor" or (engine == "auto" and len(angles) >= 16 and self.M in (4, 8, 16))
        if use_tc and self.M not in (4, 8, 16):
            raise ValueError("the tensor-core multi-beam path is compiled for 4, 8 or 16 microphones")
        y = self._run_tensor(x, W) if use_tc else self._run(x, W)
        return y if as_torch else y.double().cpu().numpy()
