"""SubbandGSC -- drop-in for ``DistantSpeech/beamformer/SubbandGSC.py`` (SubbandGSC :67, process :170-262).

Per block of ``frameLen`` samples: DC notch, time alignment, McSpp speech presence on the aligned spectrum, M blocking
filters (``SubbandLMS``: 2 frame taps per bin, input = fixed beam, desired = aligned channel, gated by p), fixed beam
delayed by one block, multichannel canceller (``SubbandLmsMc``, gated by 1 - p).  The stages feed forward only, so
each runs over the whole call as one batched launch: FIR, channel mean, STFTs, the McSpp kernels, the NLMS kernel for
the M blocking filters, ISTFT, STFT, the NLMS kernel for the canceller, ISTFT -- all on the device.

The reference module does not import as it stands (it asks FDGSC.py for a ``DelayObj`` that is not there, :23) and it
owns a ``McSpp``, so like that class it only runs with 4 microphones.  ``postfilter=True`` computes a gain the
reference then discards (:248, the synthesis is commented out), so it changes nothing and is accepted and ignored.
Extension: a leading stream axis ``x [S, M, N]``.
"""
import ctypes as C

import numpy as np

from .. import _lib as L
from ..adaptivefilter.SubbandLMS import SubbandLMS
from ..adaptivefilter.SubbandLmsMc import SubbandLmsMc
from ..noise_estimation.mcspp import McSpp
from ..transform.transform import _sqrt_hann, stft_device, istft_device
from .FDGSC import TimeAlignment
from .MicArray import MicArray
from .beamformer import beamformer


class SubbandGSC(beamformer):
    def __init__(self, mic_array: MicArray, frameLen=256, angle=[197, 0]):
        beamformer.__init__(self, mic_array, frame_len=frameLen)
        self.angle = np.array(angle) / 180 * np.pi if isinstance(angle, list) else angle
        self.time_alignment = TimeAlignment(mic_array, angle=self.angle)
        self.gamma = mic_array.gamma
        self.spp = McSpp(nfft=frameLen * 2, channels=self.M)          # raises for M != 4, like the reference fails
        self.bm = [SubbandLMS(filter_len=2, num_bands=frameLen * 2, mu=1e-1) for _ in range(self.M)]   # views, see _bm_W
        self.aic_filter = SubbandLmsMc(filter_len=2, num_bands=frameLen * 2, channel=self.M, mu=0.01, alpha=0.8)
        self._st = None

    def reset(self):
        self._st = None
        self.spp.reset()
        self.aic_filter._state = None

    def _ensure(self, S):
        t = L.require_cuda()
        L.ensure_init()
        if self._st is not None and self._st["S"] == S:
            return self._st
        M, Lf, K = self.M, self.frameLen, self.frameLen + 1
        FL = self.time_alignment.delay_filter_len
        z = (lambda *shape, dt=t.float32: t.zeros(shape, dtype=dt, device="cuda"))
        bm_prm = L.SubbandNlmsParams(K, S, M, 1, 1, 2, 0, 0, 1e-1, 0.9, 1e-4)
        self.spp.reset()
        self.aic_filter._state = None
        self._st = dict(S=S, notch=z(S, M, 2, dt=t.float64), fir=z(S, M, FL - 1, dt=t.float64),
                        h_al=z(S, M, Lf), h_fbf=z(S, 1, Lf), h_bm=z(S, M, Lf), h_fd=z(S, 1, Lf),
                        t_bm=z(S, M, Lf), t_out=z(S, 1, Lf), fbf_last=z(S, Lf, dt=t.float64),
                        bm_state=t.zeros(L.lib().ds_subband_nlms_state_bytes(C.byref(bm_prm)), dtype=t.uint8, device="cuda"))
        return self._st

    def process(self, x, postfilter=False):
        """x [n_chs, n_samples] (or [S, n_chs, n_samples]) -> (output, fix_output, bm_output, p, aligned_output)
        like SubbandGSC.py:262; the caller's array ends up DC-notched (:176-177)."""
        t = L.require_cuda()
        as_torch = isinstance(x, t.Tensor)
        xd = L.to_device(x, t.float32)
        batched = xd.dim() == 3
        if not batched:
            xd = xd[None]
        xd = xd.contiguous().clone()
        S, M, N = xd.shape
        if M != self.M:
            raise ValueError("expected %d channels, got %d" % (self.M, M))
        st = self._ensure(S)
        Lf, K = self.frameLen, self.frameLen + 1
        lib, sp = L.lib(), L.stream_ptr()
        # DC notch over the whole signal, in place (:176-177)
        L.check(lib.ds_dcnotch_run(S, M, N, 0.98, L.ptr(st["notch"]), L.ptr(xd), sp), "ds_dcnotch_run")
        if as_torch:
            (x if batched else x[None])[...] = xd.to(x.dtype)
        elif isinstance(x, np.ndarray) and x.flags.writeable:
            (x if batched else x[None])[...] = xd.cpu().numpy().astype(x.dtype)
        Nb = (N // Lf) * Lf
        T = Nb // Lf
        f64 = dict(dtype=t.float64, device="cuda")
        out = t.zeros((S, N), **f64)
        fix = t.zeros((S, N), **f64)
        bm_full = t.zeros((S, N, M), **f64)
        al_full = t.zeros((S, N, M), **f64)
        p_out = t.zeros((S, K, T), **f64)
        if T > 0:
            win = L.device_window(_sqrt_hann(2 * Lf), 2 * Lf)
            scale = Lf / float(np.sum(_sqrt_hann(2 * Lf) ** 2))
            # time alignment (float64 FIR, like the reference) and fixed beam
            xin = xd[:, :, :Nb].double().contiguous()
            aligned = t.empty_like(xin)
            scratch = t.empty_like(xin)
            h = t.as_tensor(np.ascontiguousarray(self.time_alignment.delay_filter.T)).to("cuda")
            L.check(lib.ds_fir_run(S, M, Nb, self.time_alignment.delay_filter_len, L.ptr(h), L.ptr(st["fir"]), L.ptr(xin),
                                   L.ptr(aligned), L.ptr(scratch), sp), "ds_fir_run")
            fbf = t.empty((S, Nb), **f64)
            L.check(lib.ds_channel_mean_run(S, M, Nb, L.ptr(aligned), L.ptr(fbf), sp), "ds_channel_mean_run")
            # spectra of the aligned block and the fixed beam; McSpp presence probability per frame
            D_al = stft_device(aligned.float(), 2 * Lf, Lf, win, L.DS_STFT_STREAMING, history=st["h_al"])      # [S, T, M, K]
            p = self.spp._run(D_al, want_w=False)["p"]                                                          # [S, T, K]
            X_f = stft_device(fbf.float()[:, None, :].contiguous(), 2 * Lf, Lf, win, L.DS_STFT_STREAMING, history=st["h_fbf"])
            # M blocking filters: input = fixed beam, desired = aligned channel m, gate p (:220-226)
            E_bm = t.empty((S, T, M, K), dtype=t.complex128, device="cuda")
            prm = L.SubbandNlmsParams(K, S, M, T, 1, 2, 0, 0, 1e-1, 0.9, 1e-4)
            L.check(lib.ds_subband_nlms_run(C.byref(prm), L.ptr(st["bm_state"]), L.ptr(X_f), L.ptr(D_al), L.ptr(p), L.ptr(E_bm), sp),
                    "ds_subband_nlms_run")
            bm = istft_device(E_bm, 2 * Lf, Lf, win, L.DS_STFT_STREAMING, tail=st["t_bm"], scale=scale)          # [S, M, Nb] f32
            # fixed beam delayed by one block (:229), canceller over the blocking outputs gated by 1 - p (:233-237)
            fbf_d = t.cat([st["fbf_last"], fbf[:, :Nb - Lf]], dim=1)
            st["fbf_last"] = fbf[:, Nb - Lf:].clone()
            X_bm = stft_device(bm, 2 * Lf, Lf, win, L.DS_STFT_STREAMING, history=st["h_bm"])
            D_f = stft_device(fbf_d.float()[:, None, :].contiguous(), 2 * Lf, Lf, win, L.DS_STFT_STREAMING, history=st["h_fd"])
            E = self.aic_filter._run_spec(X_bm, D_f, p, one_minus_p=True)
            y = istft_device(E, 2 * Lf, Lf, win, L.DS_STFT_STREAMING, tail=st["t_out"], scale=scale)[:, 0, :]
            out[:, :Nb] = y.double()
            fix[:, :Nb] = fbf_d
            bm_full[:, :Nb] = bm.permute(0, 2, 1).double()
            al_full[:, :Nb] = aligned.permute(0, 2, 1)
            p_out = p.permute(0, 2, 1)
        outs = [out, fix, bm_full, p_out, al_full]
        if not batched:
            outs = [o[0] for o in outs]
        if not as_torch:
            outs = [o.cpu().numpy() for o in outs]
        return tuple(outs)

    def _bm_W(self, m):
        """Weights of blocking filter m, [half_band, 2] (what ``self.bm[m].W`` is in the reference)."""
        if self._st is None:
            return np.zeros((self.frameLen + 1, 2), dtype=complex)
        t = L.require_cuda()
        b = self._st["bm_state"].view(t.float64).view(self._st["S"], self.M, 9, self.frameLen + 1)[0, m].cpu().numpy()
        return (b[0:2] + 1j * b[2:4]).T
