"""TDGSC -- drop-in for ``DistantSpeech/beamformer/TDGSC.py`` (TDGSC :21, process :110-174).

Per block of ``frameLen`` samples: DC notch, time alignment, mean fixed beam, MCRA (L = 65) on the spectrum of the
fixed beam, pairwise-difference blocking matrix, and the constrained FDAF canceller ``FastFreqLms.update`` (non-causal,
taps truncated by 30, gated per bin by 1 - p).  The stages feed forward, so each runs over the whole call as one
launch (FIR, channel mean, adjacent difference, STFT, MCRA, ``ds_fdaf_run``), all on the device.
``postfilter=True`` (:157-170) runs NsOmlsaMulti on the spectra of the output and of the blocking outputs and
resynthesises the gained output through the streaming transform -- also feed-forward (the canceller never sees the
post-filtered signal), so it is four more launches over the whole call.  Compiled for frameLen = 256.
Extension: a leading stream axis ``x [S, N, M]``.
"""
import ctypes as C

import numpy as np

from .. import _lib as L
from ..noise_estimation.mcra import NoiseEstimationMCRA
from ..noise_estimation.omlsa_multi import NsOmlsaMulti
from ..transform.transform import _sqrt_hann, stft_device, istft_device
from .FDGSC import TimeAlignment
from .MicArray import MicArray
from .beamformer import beamformer


class _CancellerView(object):
    """``aic_filter`` of the reference (a FastFreqLms): read-only view of the canceller state in the device blob."""

    def __init__(self, owner):
        self._o = owner
        self.filter_len = owner.frameLen
        self.n_fft = 2 * owner.frameLen
        self.n_channels = owner.M - 1

    @property
    def W(self):
        return self._o._canceller_W()

    @property
    def w(self):
        return np.fft.irfft(self.W, n=self.n_fft, axis=-2)[..., : self.filter_len, :]


class TDGSC(beamformer):
    def __init__(self, mic_array: MicArray, frameLen=256, angle=[197, 0]):
        if frameLen != 256:
            raise L.DsError("the CUDA TDGSC is compiled for frameLen = 256 (reference default)")
        if not 2 <= mic_array.M <= 8:
            raise ValueError("TDGSC on the device takes 2..8 microphones")
        beamformer.__init__(self, mic_array, frame_len=frameLen)
        self.mic_array = mic_array
        self.angle = np.array(angle) / 180 * np.pi if isinstance(angle, list) else angle
        self.time_alignment = TimeAlignment(mic_array, angle=self.angle)
        self.mcra = NoiseEstimationMCRA(nfft=frameLen * 2)
        self.mcra.L = 65
        self.spp = self.mcra
        self.omlsa_multi = NsOmlsaMulti(nfft=frameLen * 2, cal_weights=True, M=self.M)
        self.mu, self.alpha, self.fir_truncate = 0.01, 0.9, 30
        self._gated, self._non_causal = True, True        # GSC.process1 runs the same chain ungated and causal
        self._st = None
        self.aic_filter = _CancellerView(self)

    def reset(self):
        self._st = None
        self.mcra = NoiseEstimationMCRA(nfft=self.frameLen * 2)
        self.mcra.L = 65
        self.spp = self.mcra
        self.omlsa_multi = NsOmlsaMulti(nfft=self.frameLen * 2, cal_weights=True, M=self.M)

    def _ensure(self, S):
        t = L.require_cuda()
        L.ensure_init()
        if self._st is None or self._st["S"] != S:
            M, Lf = self.M, self.frameLen
            z = (lambda *shape, dt=t.float32: t.zeros(shape, dtype=dt, device="cuda"))
            self.reset()
            self._st = dict(S=S, notch=z(S, M, 2, dt=t.float64), fir=z(S, M, self.time_alignment.delay_filter_len - 1, dt=t.float64),
                            h_fbf=z(S, 1, Lf),
                            h_pf_y=z(S, 1, Lf), h_pf_u=z(S, M - 1, Lf), t_pf=z(S, 1, Lf),      # postfilter transforms (:40-41)
                            fdaf=t.zeros(L.lib().ds_fdaf_state_bytes(S, M - 1), dtype=t.uint8, device="cuda"))
        return self._st

    def _canceller_W(self):
        """Canceller weights [257, M-1] complex (``aic_filter.W`` of the reference)."""
        if self._st is None:
            return np.zeros((self.frameLen + 1, self.M - 1), dtype=complex)
        t = L.require_cuda()
        Cn, K = self.M - 1, self.frameLen + 1
        per = Cn * K * 2 + K + Cn * self.frameLen + self.frameLen // 2
        blob = self._st["fdaf"].view(t.float32).view(self._st["S"], per)[:, :Cn * K * 2].reshape(-1, Cn, K, 2).cpu().numpy()
        w = (blob[..., 0] + 1j * blob[..., 1]).transpose(0, 2, 1)
        return w[0] if w.shape[0] == 1 else w

    def process(self, x, postfilter=False):
        """x [samples, chs] (or [S, samples, chs]) -> (output, p [257, n_blocks], output_bm [samples, chs-1]);
        the caller's array ends up DC-notched (:130-131)."""
        t = L.require_cuda()
        as_torch = isinstance(x, t.Tensor)
        xd = L.to_device(x, t.float32)
        batched = xd.dim() == 3
        if not batched:
            xd = xd[None]
        S, N, M = xd.shape
        if M != self.M:
            raise ValueError("expected %d channels, got %d" % (self.M, M))
        st = self._ensure(S)
        Lf, K = self.frameLen, self.frameLen + 1
        lib, sp = L.lib(), L.stream_ptr()
        xs = xd.permute(0, 2, 1).contiguous()                                          # [S, M, N]
        L.check(lib.ds_dcnotch_run(S, M, N, 0.98, L.ptr(st["notch"]), L.ptr(xs), sp), "ds_dcnotch_run")
        notched = xs.permute(0, 2, 1)
        if as_torch:
            (x if batched else x[None])[...] = notched.to(x.dtype)
        elif isinstance(x, np.ndarray) and x.flags.writeable:
            (x if batched else x[None])[...] = notched.cpu().numpy().astype(x.dtype)
        Nb = (N // Lf) * Lf
        T = Nb // Lf
        f64 = dict(dtype=t.float64, device="cuda")
        out = t.zeros((S, N), **f64)
        out_bm = t.zeros((S, N, M - 1), **f64)
        p_out = t.zeros((S, K, T), **f64)
        if T > 0:
            xin = xs[:, :, :Nb].double().contiguous()
            aligned = t.empty_like(xin)
            scratch = t.empty_like(xin)
            h = t.as_tensor(np.ascontiguousarray(self.time_alignment.delay_filter.T)).to("cuda")
            L.check(lib.ds_fir_run(S, M, Nb, self.time_alignment.delay_filter_len, L.ptr(h), L.ptr(st["fir"]), L.ptr(xin),
                                   L.ptr(aligned), L.ptr(scratch), sp), "ds_fir_run")
            fbf = t.empty((S, Nb), **f64)
            L.check(lib.ds_channel_mean_run(S, M, Nb, L.ptr(aligned), L.ptr(fbf), sp), "ds_channel_mean_run")
            bm = t.empty((S, M - 1, Nb), dtype=t.float32, device="cuda")
            L.check(lib.ds_adjacent_diff_run(S, M, Nb, L.ptr(aligned), L.ptr(bm), sp), "ds_adjacent_diff_run")
            fbf32 = fbf.float()
            p = None
            if self._gated:
                win = L.device_window(_sqrt_hann(2 * Lf), 2 * Lf)
                D = stft_device(fbf32[:, None, :].contiguous(), 2 * Lf, Lf, win, L.DS_STFT_STREAMING, history=st["h_fbf"])   # [S, T, 1, K]
                pw = L.spectral_power(D[:, :, 0, :].to(t.complex128), via_abs=True)                                       # [S, T, K]
                _, p = self.mcra._run(pw, want_p=True)                                                                    # [S, T, K]
                p = p.contiguous()
            e = t.empty((S, Nb), dtype=t.float32, device="cuda")
            prm = L.FdafParams(Lf, S, M - 1, Nb, int(self.fir_truncate), int(self._non_causal), int(self._gated), 0,
                               float(self.mu), float(self.alpha))
            L.check(lib.ds_fdaf_run(C.byref(prm), L.ptr(st["fdaf"]), L.ptr(bm), L.ptr(fbf32.contiguous()), L.ptr(p), L.ptr(e), sp),
                    "ds_fdaf_run")
            if postfilter:
                # Y = transform_fbf.stft(output_n), U = transform_bm.stft(bm_output), NsOmlsaMulti on their powers,
                # Y *= sqrt(G), output_n = transform_fbf.istft(Y)                                            (:157-170)
                win = L.device_window(_sqrt_hann(2 * Lf), 2 * Lf)
                Yd = stft_device(e[:, None, :].contiguous(), 2 * Lf, Lf, win, L.DS_STFT_STREAMING, history=st["h_pf_y"])   # [S, T, 1, K]
                Ud = stft_device(bm.contiguous(), 2 * Lf, Lf, win, L.DS_STFT_STREAMING, history=st["h_pf_u"])              # [S, T, M-1, K]
                G, lam, pp = self.omlsa_multi._run(L.spectral_power(Yd).view(S, T, K), L.spectral_power(Ud))
                last = [v[0, -1] if S == 1 else v[:, -1] for v in (G, lam, pp)]
                self.omlsa_multi.G, self.omlsa_multi.lambda_d, self.omlsa_multi.p = (v.cpu().numpy() for v in last)
                Yg = t.empty((S, T, 1, K), dtype=t.complex128, device="cuda")
                L.check(lib.ds_spectral_gain_run(S * T * K, L.ptr(Yd), 0, L.ptr(G), 1, L.ptr(Yg), sp), "ds_spectral_gain_run")
                e = istft_device(Yg, 2 * Lf, Lf, win, L.DS_STFT_STREAMING, tail=st["t_pf"],
                                 scale=Lf / float(np.sum(_sqrt_hann(2 * Lf) ** 2)))[:, 0, :]
            out[:, :Nb] = e.double()
            out_bm[:, :Nb] = bm.permute(0, 2, 1).double()
            if p is not None:
                p_out = p.permute(0, 2, 1)
        outs = [out, p_out, out_bm]
        if not batched:
            outs = [o[0] for o in outs]
        if not as_torch:
            outs = [o.cpu().numpy() for o in outs]
        return tuple(outs)
