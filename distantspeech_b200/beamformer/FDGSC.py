"""Overlap-save frequency-domain GSC -- drop-in for ``DistantSpeech/beamformer/FDGSC.py``
(FDGSC :37, process :201) with its building blocks ``TimeAlignment``
(fixedbeamformer.py:51-93), ``AdaptiveBlockingMatrixFilter`` (gsc_bm.py),
``AdaptiveInterferenceCancellation`` (gsc_aic.py), ``FilterDcNotch16`` (feature.py:32-49)
and the delay lines.  Two device implementations with identical results and state: ``impl="pipeline"`` (default;
ds_fdgsc_run_ws: feed-forward alignment / spectra / detector kernels, one warp per blocking filter, one CTA per
canceller, intermediate signals through a scratch workspace) and ``impl="fused"`` (ds_fdgsc_run: one CTA per stream,
every filter state resident in shared memory for the whole utterance, no workspace).

``process(x[N, M])`` returns the reference's 9-tuple; ``x`` is overwritten with the
DC-notched signal like the reference does (FDGSC.py:213).  A leading stream axis
``x[S, N, M]`` batches independent streams.
"""
import ctypes as C

import numpy as np

from .. import _lib as L
from ..transform.multirate import fractional_delay_filter_bank
from ..noise_estimation.omlsa_multi import NsOmlsaMulti
from ..transform.transform import _sqrt_hann, stft_device, istft_device
from .MicArray import MicArray


class TimeAlignment(object):
    """Fractional-delay alignment towards ``angle`` (fixedbeamformer.py:51-93)."""

    def __init__(self, mic_array: MicArray, angle=[197, 0], frame_len=256, hop=None, nfft=None, r=0.032, fs=16000):
        self.M = mic_array.M
        self.angle = np.array(angle) / 180 * np.pi if isinstance(angle, list) else angle
        self.tau = mic_array.compute_tau(self.angle)
        self.tau = -(self.tau - np.max(self.tau))
        delay_samples = np.array(self.tau)[:, 0] * mic_array.fs
        self.delay_filter = fractional_delay_filter_bank(delay_samples)           # [filter_len, M]
        self.delay_filter_len = self.delay_filter.shape[0]
        self.fir_cache = np.zeros((self.delay_filter_len - 1, self.M))
        self._cache_dev = None

    def process(self, x):
        """x [samples, chs] -> time-aligned [samples, chs] (streaming FIR, ds_fir_run)."""
        t = L.require_cuda()
        xd = t.as_tensor(np.ascontiguousarray(np.asarray(x, dtype=np.float64).T)).to("cuda")[None]     # [1, M, N]
        _, M, N = xd.shape
        if self._cache_dev is None:
            self._cache_dev = t.as_tensor(np.ascontiguousarray(self.fir_cache.T)).to("cuda")[None].contiguous()
        h = t.as_tensor(np.ascontiguousarray(self.delay_filter.T)).to("cuda")
        y = t.empty_like(xd)
        scratch = t.empty_like(xd)
        L.check(L.lib().ds_fir_run(1, M, N, self.delay_filter_len, L.ptr(h), L.ptr(self._cache_dev), L.ptr(xd), L.ptr(y),
                                   L.ptr(scratch), L.stream_ptr()), "ds_fir_run")
        self.fir_cache = self._cache_dev[0].t().cpu().numpy()
        return y[0].t().cpu().numpy()


class _FilterView(object):
    """Read-only view of one adaptive filter (``W`` [257, chs] complex, ``w`` [256, chs]) of the fused kernel."""

    def __init__(self, W):
        self.W = W
        self.n_fft = 2 * (W.shape[0] - 1)
        self.filter_len = self.n_fft // 2
        self.w = np.fft.irfft(W, n=self.n_fft, axis=0)[: self.filter_len, :]


class _McraView(object):
    def __init__(self):
        self.L = 60
        self.alpha_d, self.alpha_s, self.delta_s, self.alpha_p = 0.95, 0.8, 5, 0.2
        self.p_max, self.p_min = 0.999, 1e-3
        self.ell, self.frm_cnt = 1, 0
        self.half_bin = 257
        self.p = np.zeros(257)


class FDGSC(object):
    def __init__(self, mic_array: MicArray, frameLen=256, angle=[197, 0], precision="fp32", impl="pipeline"):
        if frameLen != 256:
            raise L.DsError("the CUDA FDGSC is compiled for frameLen = 256 (reference default)")
        self.MicArray = mic_array
        self.M = mic_array.M
        self.frameLen = frameLen
        self.hop = frameLen // 2
        self.nfft = frameLen * 2
        self.fs = mic_array.fs
        self.angle = np.array(angle) / 180 * np.pi if isinstance(angle, list) else angle
        self.time_alignment = TimeAlignment(mic_array, angle=self.angle)
        self.gamma = mic_array.gamma
        self.spp = _McraView()
        self.precision = precision
        if impl not in ("pipeline", "fused"):
            raise ValueError("impl must be 'pipeline' (kernels cut along the data dependences, needs a workspace) or "
                             "'fused' (one CTA per stream, no workspace)")
        self.impl = impl
        self._ws = None
        self.mu_bm, self.mu_aic, self.alpha = 0.1, 0.1, 0.9
        self.phi, self.psi = None, None          # ccafbounds: computed by the reference ctor but never used
        self._state = None
        self._pf = None
        self._diag = None
        self._S = None
        self.bm = None
        self.aic_filter = None

    def _params(self, S, N):
        p = L.FdgscParams()
        L.lib().ds_fdgsc_default_params(C.byref(p), S, self.M, N, self.time_alignment.delay_filter_len)
        m = self.spp
        p.frm_cnt, p.ell, p.mcra_L = int(m.frm_cnt), int(m.ell), int(m.L)
        p.fp64 = int(self.precision == "fp64")
        p.mu_bm, p.mu_aic, p.alpha = float(self.mu_bm), float(self.mu_aic), float(self.alpha)
        return p

    def _launch(self, prm, h, xrun, yrun, bm_out, fix_out, p_out):
        """ds_fdgsc_run_ws (pipeline of kernels, scratch workspace kept between calls) or ds_fdgsc_run (fused kernel)."""
        t = L.require_cuda()
        win = L.device_window(_sqrt_hann(512), 512)
        if self.impl == "fused":
            L.check(L.lib().ds_fdgsc_run(C.byref(prm), L.ptr(h), L.ptr(win), L.ptr(self._state), L.ptr(xrun), L.ptr(yrun),
                                         L.ptr(bm_out), L.ptr(fix_out), L.ptr(p_out), L.stream_ptr()), "ds_fdgsc_run")
            return
        need = L.lib().ds_fdgsc_workspace_bytes(C.byref(prm))
        if self._ws is None or self._ws.numel() < need or self._ws.device.index != t.cuda.current_device():
            self._ws = None
            self._ws = t.empty(need, dtype=t.uint8, device="cuda")
        L.check(L.lib().ds_fdgsc_run_ws(C.byref(prm), L.ptr(h), L.ptr(win), L.ptr(self._state), L.ptr(self._ws), L.ptr(xrun),
                                        L.ptr(yrun), L.ptr(bm_out), L.ptr(fix_out), L.ptr(p_out), L.stream_ptr()), "ds_fdgsc_run_ws")

    def reset_state(self):
        """Zero the recursive state in place (a fresh utterance for the same batch size, no reallocation)."""
        if self._state is not None:
            self._state.zero_()
        self._pf = None
        self._diag = None
        self.spp.frm_cnt, self.spp.ell = 0, 1

    def reset(self):
        self._state = None
        self._pf = None
        self._diag = None
        self.spp.frm_cnt, self.spp.ell = 0, 1

    def _postfilter(self, y_aic, fix_out, bm_out):
        """postfilter=True exactly as FDGSC.py:283-295 behaves (all on the device):
          * the beam spectrum comes from ``transform_fbf``, the Transform that :270 has just fed the delayed
            fixed-beamformer block -> analysed frame n = window * [fixed_output_delayed_n | aic_output_n];
          * ``transform_bm.stft(bm_output[:, :-1])`` sees the WHOLE output buffer every block and only its
            frame 0 is used -> the reference powers are those of [tail of the previous call's buffer | block 0],
            the same for every block of a call;
          * gain sqrt(G) of NsOmlsaMulti (G = 1 on the very first frame), then transform_fbf.istft.
        y_aic, fix_out [S, Nb] float32, bm_out [S, M, Nb] float32 -> [S, Nb] float32."""
        t = L.require_cuda()
        S, Nb = y_aic.shape
        Lf, M, K = self.frameLen, self.M, self.frameLen + 1
        nblk = Nb // Lf
        win = L.device_window(_sqrt_hann(2 * Lf), 2 * Lf)
        if getattr(self, "_pf", None) is None or self._pf["S"] != S:
            self._pf = dict(S=S, fix_last=t.zeros((S, Lf), dtype=t.float32, device="cuda"),
                            bm_hist=t.zeros((S, M - 1, Lf), dtype=t.float32, device="cuda"),
                            tail=t.zeros((S, 1, Lf), dtype=t.float32, device="cuda"),
                            omlsa=NsOmlsaMulti(nfft=2 * Lf, cal_weights=True, M=M))
        pf = self._pf
        # frames [fixed_output_delayed_n | aic_output_n]: a hop = n_fft transform of the interleaved blocks
        fixdel = t.cat([pf["fix_last"], fix_out[:, :Nb - Lf]], dim=1)
        pf["fix_last"] = fix_out[:, Nb - Lf:Nb].clone()
        z = t.stack([fixdel.view(S, nblk, Lf), y_aic.view(S, nblk, Lf)], dim=2).reshape(S, 1, 2 * Nb).contiguous()
        Y = stft_device(z, 2 * Lf, 2 * Lf, win, L.DS_STFT_PLAIN)                                    # [S, nblk, 1, K] c64
        # reference powers: frame 0 of the buffer-wide transform
        hist = pf["bm_hist"].clone()
        U0 = stft_device(bm_out[:, :M - 1, :Lf].contiguous(), 2 * Lf, Lf, win, L.DS_STFT_STREAMING, history=hist)
        pf["bm_hist"] = bm_out[:, :M - 1, Nb - Lf:Nb].contiguous()
        Ypow = L.spectral_power(Y).view(S, nblk, K)                       # np.real(Y * conj(Y))   (:288)
        Upow = L.spectral_power(U0).view(S, M - 1, K)
        G, _, _ = pf["omlsa"]._run(Ypow, Upow, u_const=True)
        Yg = t.empty((S, nblk, 1, K), dtype=t.complex128, device="cuda")
        L.check(L.lib().ds_spectral_gain_run(Ypow.numel(), L.ptr(Y), 0, L.ptr(G), 1, L.ptr(Yg), L.stream_ptr()),
                "ds_spectral_gain_run")
        out = istft_device(Yg, 2 * Lf, Lf, win, L.DS_STFT_STREAMING, tail=pf["tail"], scale=Lf / float(np.sum(_sqrt_hann(2 * Lf) ** 2)),
                           fft_fp64=self.precision == "fp64")
        return out[:, 0, :]

    def _alignment_diagnostics(self, x_notched, fix_out):
        """fix_output_delayed, aligned_output, aligned_output_delayed of the return tuple (FDGSC.py:256,267,299-302).
        The fused kernel keeps them on chip; here they are rebuilt from its inputs/outputs with the streaming FIR
        (ds_fir_run, float64 like the reference) and two carried delay lines (frameLen and frameLen/2 samples).
        x_notched [S, M, Nb] float32, fix_out [S, Nb] float32 -> three float64 CUDA tensors."""
        t = L.require_cuda()
        S, M, Nb = x_notched.shape
        Lf, FL = self.frameLen, self.time_alignment.delay_filter_len
        d = self._diag
        if d is None or d["S"] != S:
            d = self._diag = dict(S=S, cache=t.zeros((S, M, FL - 1), dtype=t.float64, device="cuda"),
                                  al=t.zeros((S, M, Lf // 2), dtype=t.float64, device="cuda"),
                                  fix=t.zeros((S, Lf), dtype=t.float64, device="cuda"))
        h = t.as_tensor(np.ascontiguousarray(self.time_alignment.delay_filter.T)).to("cuda")
        xin = x_notched.double().contiguous()
        aligned = t.empty_like(xin)
        scratch = t.empty_like(xin)
        L.check(L.lib().ds_fir_run(S, M, Nb, FL, L.ptr(h), L.ptr(d["cache"]), L.ptr(xin), L.ptr(aligned), L.ptr(scratch),
                                   L.stream_ptr()), "ds_fir_run")
        al_d = t.cat([d["al"], aligned[:, :, :Nb - Lf // 2]], dim=2)
        d["al"] = aligned[:, :, Nb - Lf // 2:].clone()
        fx = fix_out.double()
        fix_d = t.cat([d["fix"], fx[:, :Nb - Lf]], dim=1)
        d["fix"] = fx[:, Nb - Lf:].clone()
        return fix_d, aligned.permute(0, 2, 1), al_d.permute(0, 2, 1)

    def process_device(self, xs, out=None, dc_notch=True):
        """Device entry point of the batched path: xs [S, M, N] float32 CUDA (mic-major, N a multiple of frameLen;
        overwritten with the DC-notched signal like FDGSC.py:213) -> output [S, N] float32 CUDA.  Only ``output`` of
        the reference's tuple is produced (no diagnostics streams are written); the recursive state carries over
        between calls exactly like ``process``."""
        t = L.require_cuda()
        L.ensure_init()
        if not isinstance(xs, t.Tensor) or not xs.is_cuda or xs.dim() != 3 or xs.dtype != t.float32 or not xs.is_contiguous():
            raise ValueError("xs must be a contiguous float32 CUDA tensor [S, M, N]")
        S, M, N = xs.shape
        if M != self.M or N < self.frameLen or N % self.frameLen != 0:
            raise ValueError("expected [S, %d, N] with N a positive multiple of frameLen=%d" % (self.M, self.frameLen))
        if self._state is None or self._S != S:
            self._state = t.zeros(L.lib().ds_fdgsc_state_bytes(C.byref(self._params(S, N))), dtype=t.uint8, device="cuda")
            self._S = S
            self.spp.frm_cnt, self.spp.ell = 0, 1
        prm = self._params(S, N)
        prm.dc_notch = int(bool(dc_notch))
        y = out if out is not None else t.empty((S, N), dtype=t.float32, device="cuda")
        key = (t.cuda.current_device(),)
        if getattr(self, "_h_dev", None) is None or self._h_dev[0] != key:
            self._h_dev = (key, t.as_tensor(np.ascontiguousarray(self.time_alignment.delay_filter.T)).to("cuda"))     # [M, FL]
        self._launch(prm, self._h_dev[1], xs, y, None, None, None)
        f, e = C.c_int32(self.spp.frm_cnt), C.c_int32(self.spp.ell)
        L.lib().ds_mcra_advance(int(self.spp.L), N // self.frameLen, C.byref(f), C.byref(e))
        self.spp.frm_cnt, self.spp.ell = f.value, e.value
        return y

    def process(self, x, postfilter=False, dc_notch=True, diagnostics=None):
        """x [n_samples, n_chs] (or [S, n_samples, n_chs]) -> (output, p, fix_output, fix_output_delayed,
        bm_output, aligned_output, aligned_output_delayed, bm, aic_filter) like FDGSC.py:307-317.
        ``diagnostics``: materialise fix_output_delayed / aligned_output / aligned_output_delayed (default: yes
        for the reference's single-stream call, no -- None in the tuple -- for the batched extension, where they
        would cost three float64 copies of the input)."""
        t = L.require_cuda()
        L.ensure_init()
        as_torch = isinstance(x, t.Tensor)
        xd = L.to_device(x, t.float32)
        batched = xd.dim() == 3
        if not batched:
            xd = xd[None]
        S, N, M = xd.shape
        if M != self.M:
            raise ValueError("expected %d channels, got %d" % (self.M, M))
        Nb = (N // self.frameLen) * self.frameLen
        xs = xd.permute(0, 2, 1).contiguous()                                    # [S, M, N]
        if self._state is None or self._S != S:
            self._state = t.zeros(L.lib().ds_fdgsc_state_bytes(C.byref(self._params(S, Nb))), dtype=t.uint8, device="cuda")
            self._S = S
            self.spp.frm_cnt, self.spp.ell = 0, 1
        prm = self._params(S, Nb)
        prm.dc_notch = int(bool(dc_notch))
        nblk = Nb // self.frameLen
        xrun = xs if Nb == N else xs[:, :, :Nb].contiguous()
        y = t.zeros((S, N), dtype=t.float32, device="cuda")
        yrun = y if Nb == N else t.empty((S, Nb), dtype=t.float32, device="cuda")
        bm_out = t.zeros((S, M, Nb), dtype=t.float32, device="cuda")
        fix_out = t.zeros((S, Nb), dtype=t.float32, device="cuda")
        p_out = t.empty((S, nblk, 257), dtype=t.float64, device="cuda")
        h = t.as_tensor(np.ascontiguousarray(self.time_alignment.delay_filter.T)).to("cuda")       # [M, FL]
        self._launch(prm, h, xrun, yrun, bm_out, fix_out, p_out)
        f, e = C.c_int32(self.spp.frm_cnt), C.c_int32(self.spp.ell)
        L.lib().ds_mcra_advance(int(self.spp.L), nblk, C.byref(f), C.byref(e))
        self.spp.frm_cnt, self.spp.ell = f.value, e.value
        if postfilter:
            yrun = self._postfilter(yrun, fix_out, bm_out)
            if Nb == N:
                y = yrun
        if Nb != N:
            y[:, :Nb] = yrun
        # quirk 11: the caller's array now holds the DC-notched signal -- all of it, also the samples beyond the
        # last full block, whose only other effect is on the notch memories the next call starts from
        if dc_notch:
            notched = xrun.permute(0, 2, 1)
            if Nb != N:
                tail = xs[:, :, Nb:].contiguous()
                L.check(L.lib().ds_fdgsc_notch_run(C.byref(prm), L.ptr(self._state), L.ptr(tail), N - Nb, L.stream_ptr()),
                        "ds_fdgsc_notch_run")
                notched = t.cat([notched, tail.permute(0, 2, 1)], dim=1)
            if as_torch:
                (x if batched else x[None])[:, :, :] = notched.to(x.dtype)
            elif isinstance(x, np.ndarray) and x.flags.writeable:
                xv = x if batched else x[None]
                xv[:, :, :] = notched.cpu().numpy().astype(x.dtype)
        # filter views from the state blob
        K = 257
        st = self._state.view(t.float64).view(S, -1)
        nW = 2 * M * K
        Wbm = t.view_as_complex(st[:, :nW].reshape(S, M, K, 2).contiguous()).cpu().numpy()
        Waic = t.view_as_complex(st[:, nW:2 * nW].reshape(S, M, K, 2).contiguous()).cpu().numpy()
        self.bm = [_FilterView(Wbm[0, m][:, None]) for m in range(M)]
        self.aic_filter = _FilterView(Waic[0].T)
        p = p_out.permute(0, 2, 1)
        self.spp.p = p[0, :, -1].cpu().numpy()
        diag = [None, None, None]
        if (not batched) if diagnostics is None else diagnostics:
            pad = (lambda v: v if Nb == N else t.nn.functional.pad(v, (0, 0, 0, N - Nb) if v.dim() == 3 else (0, N - Nb)))
            diag = [pad(v) for v in self._alignment_diagnostics(xrun, fix_out)]
        outs = [y, p, fix_out, diag[0], bm_out.permute(0, 2, 1), diag[1], diag[2]]
        if not batched:
            outs = [o[0] if o is not None else None for o in outs]
        if not as_torch:
            outs = [o.double().cpu().numpy() if o is not None else None for o in outs]
        return (outs[0], outs[1], outs[2], outs[3], outs[4], outs[5], outs[6], self.bm, self.aic_filter)
