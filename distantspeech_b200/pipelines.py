"""Batched pipelines over many independent microphone-array streams.

``MvdrMcsppChain`` is the config-4 composition pinned in SURVEY.md 8c -- the
reference has no single function for it; the closest call sites are
example/mcsppbase.ipynb cell 3, example/mvdr.ipynb cell 4 and GSC.py:225,286:

    D = Transform(n_fft, hop, channel=M).stft(x)
    est = McSppBase(nfft, channels=M); a0 = beamformer(...).compute_steering_vector_from_doa(look)
    for n: est.estimation(D[:, n, :]); w = compute_mvdr_weight(a0, est.Phi_vv_inv)
           est.compute_omlsa_weight(est.xi, est.p); Y[:, n] = (w^H D[:, n, :]) * est.G
    y = Transform(n_fft, hop, channel=1).istft(Y)

All three stages run in CUDA (ds_chain_run).  Streams are independent, so a job
shards over GPUs by splitting the stream axis; there is no collective.
"""
import ctypes as C

import numpy as np

from . import _lib as L
from .beamformer.beamformer import beamformer
from .transform.transform import _sqrt_hann


class MvdrMcsppChain(object):
    def __init__(self, mic_array, look_angle=(0, 0), n_fft=512, hop=256, apply_gain=True, full_state=False,
                 fft_precision="fp32"):
        self.mic_array = mic_array
        self.M = mic_array.M
        self.n_fft, self.hop = int(n_fft), int(hop)
        self.K = self.n_fft // 2 + 1
        self.apply_gain = bool(apply_gain)
        self.full_state = bool(full_state)
        self.fft_fp64 = fft_precision == "fp64"
        self.window = _sqrt_hann(self.n_fft)
        self.W0 = float(np.sum(self.window ** 2))
        bf = beamformer(mic_array, frame_len=self.n_fft, hop=self.hop, nfft=self.n_fft)
        self.a0 = bf.compute_steering_vector_from_doa(look_angle)          # [K, M]
        self.frm_cnt, self.ell, self.mcra_L = 0, 1, 15
        self._state = None
        self._ws = None
        self._key = None
        self._a0_dev = None
        self._hp = None

    # ------------------------------------------------------------------
    def _params(self, S, N):
        p = L.ChainParams()
        L.lib().ds_mcspp_default_params(C.byref(p.est), self.n_fft, S, self.M, N // self.hop)
        p.est.frm_cnt, p.est.ell, p.est.mcra_L = self.frm_cnt, self.ell, self.mcra_L
        p.est.full_state = int(self.full_state)
        p.hop, p.n_samples, p.fft_fp64, p.apply_gain = self.hop, N, int(self.fft_fp64), int(self.apply_gain)
        p.scale = self.hop / self.W0
        return p

    def reset(self):
        self._state = None
        self.frm_cnt, self.ell = 0, 1

    def _prepare(self, S, N):
        t = L.require_cuda()
        L.ensure_init()
        p = self._params(S, N)
        if self._state is None or self._key is None or self._key[0] != S:
            self._state = t.zeros(L.lib().ds_chain_state_bytes(C.byref(p)), dtype=t.uint8, device="cuda")
            self.frm_cnt, self.ell = 0, 1
            p = self._params(S, N)
        need = L.lib().ds_chain_workspace_bytes(C.byref(p))
        if self._ws is None or self._ws.numel() < need or self._ws.device.index != t.cuda.current_device():
            self._ws = None
            self._ws = t.empty(need, dtype=t.uint8, device="cuda")
        self._key = (S, N)
        if self._a0_dev is None or self._a0_dev.device.index != t.cuda.current_device():
            self._a0_dev = t.as_tensor(np.ascontiguousarray(self.a0.T)).to("cuda")     # [M, K] complex128
        return p

    def _check_input(self, x_dev, out):
        """Shared validation of the device entry points: x_dev [S, M, N] contiguous float32 or int16 CUDA tensor."""
        t = L.require_cuda()
        if not isinstance(x_dev, t.Tensor) or x_dev.dim() != 3 or not x_dev.is_cuda:
            raise ValueError("x_dev must be a CUDA tensor [S, %d, N]" % self.M)
        S, M, N = x_dev.shape
        if M != self.M or N % self.hop != 0 or N < self.hop:
            raise ValueError("expected [S, %d, N] with N a positive multiple of hop=%d" % (self.M, self.hop))
        if x_dev.dtype not in (t.float32, t.int16) or not x_dev.is_contiguous():
            raise ValueError("x_dev must be a contiguous float32 (or int16 PCM) CUDA tensor")
        if out is not None:
            if (not out.is_cuda or tuple(out.shape) != (S, N) or out.dtype not in (t.float32, t.int16)
                    or not out.is_contiguous()):
                raise ValueError("out must be a contiguous float32 (or int16 PCM) CUDA tensor [S, N]")
        return S, M, N

    def _advance(self, N):
        f, e = C.c_int32(self.frm_cnt), C.c_int32(self.ell)
        L.lib().ds_mcra_advance(self.mcra_L, N // self.hop, C.byref(f), C.byref(e))
        self.frm_cnt, self.ell = f.value, e.value

    def process_device(self, x_dev, out=None):
        """x_dev [S, M, N] float32 CUDA (mic-major) -> y [S, N] float32 CUDA.  No copies, no sync.
        int16 PCM on either side (x_dev and / or ``out`` of dtype int16) runs load_audio's ``/ 32767`` and
        save_audio's ``* 32767 -> int16`` (beamformer/utils.py:182-196) inside the analysis / synthesis kernels."""
        t = L.require_cuda()
        S, M, N = self._check_input(x_dev, out)
        p = self._prepare(S, N)
        y = out if out is not None else t.empty((S, N), dtype=t.float32, device="cuda")
        L.check(L.lib().ds_chain_run_io(C.byref(p), L.ptr(L.device_window(self.window, self.n_fft)), L.ptr(self._a0_dev),
                                        L.ptr(self._state), L.ptr(self._ws), L.ptr(x_dev), int(x_dev.dtype == t.int16),
                                        L.ptr(y), int(y.dtype == t.int16), L.stream_ptr()), "ds_chain_run_io")
        self._advance(N)
        return y

    def process_device_profiled(self, x_dev, out=None):
        """Like process_device (float32 only) but synchronises and returns (y, [ms_analysis, ms_perbin, ms_synthesis])
        measured with CUDA events on the launching stream (bench/roofline use only)."""
        t = L.require_cuda()
        S, M, N = self._check_input(x_dev, out)
        if x_dev.dtype != t.float32 or (out is not None and out.dtype != t.float32):
            raise ValueError("process_device_profiled takes float32 tensors")
        p = self._prepare(S, N)
        y = out if out is not None else t.empty((S, N), dtype=t.float32, device="cuda")
        ms = (C.c_float * 3)()
        L.check(L.lib().ds_chain_run_profiled(C.byref(p), L.ptr(L.device_window(self.window, self.n_fft)),
                                              L.ptr(self._a0_dev), L.ptr(self._state), L.ptr(self._ws), L.ptr(x_dev),
                                              L.ptr(y), L.stream_ptr(), ms), "ds_chain_run_profiled")
        self._advance(N)
        return y, [float(v) for v in ms]

    def process(self, x):
        """Reference-shaped call: x [N, M] (or [S, N, M]) NumPy / torch -> y [N] (or [S, N]).
        NumPy in -> NumPy (float64 holding float32 values, like Transform.istft) out."""
        t = L.require_cuda()
        as_torch = isinstance(x, t.Tensor)
        xd = L.to_device(x, t.float32)
        batched = xd.dim() == 3
        if not batched:
            xd = xd[None]
        y = self.process_device(xd.permute(0, 2, 1).contiguous())
        if not batched:
            y = y[0]
        return y if as_torch else y.double().cpu().numpy()

    # ------------------------------------------------------------------
    def _host_pipeline(self, S, M, n_slice, in_dtype, out_dtype):
        """Streams, events, staging buffers and the chain object that owns the batch's recursive state for process_host,
        created once per (batch, slice, dtypes) and kept for the life of the object."""
        t = L.require_cuda()
        key = (S, M, n_slice, in_dtype, out_dtype, t.cuda.current_device())
        hp = getattr(self, "_hp", None)
        if hp is not None and hp["key"] == key:
            return hp
        sub = MvdrMcsppChain.__new__(MvdrMcsppChain)        # the batch's own recursive state, separate from process()'s
        sub.__dict__.update(self.__dict__)
        sub._state = sub._ws = sub._key = sub._hp = None
        hp = {"key": key, "s_in": t.cuda.Stream(), "s_out": t.cuda.Stream(),
              "xbuf": [t.empty((S, M, n_slice), dtype=in_dtype, device="cuda") for _ in range(2)],
              "ybuf": [t.empty((S, n_slice), dtype=out_dtype, device="cuda") for _ in range(2)],
              "ev_in": [t.cuda.Event() for _ in range(2)], "ev_done": [t.cuda.Event() for _ in range(2)],
              "ev_out": [t.cuda.Event() for _ in range(2)], "sub": sub}
        self._hp = hp
        return hp

    @staticmethod
    def _slices(T, slice_frames):
        """Frame ranges of the time slices: equal slices, then a short last one so that the pipeline drains quickly."""
        out, lo = [], 0
        while lo < T:
            n = min(slice_frames, T - lo)
            if T - lo - n == 0 and n > max(4, slice_frames // 4) and len(out) > 0:
                n = n - max(4, slice_frames // 4)          # split the final slice: big part + short tail
            out.append((lo, lo + n))
            lo += n
        return out

    def process_host(self, x_host, y_host=None, slice_frames=16):
        """End-to-end call with HOST buffers: x_host [S, M, N] float32 or int16 PCM (the reference's on-disk
        format; scaled exactly like load_audio, float32(pcm) / 32767, utils.py:184-185) -> y_host [S, N] float32,
        or int16 PCM when an int16 ``y_host`` is handed in (save_audio's (audio * 32767).astype(int16), utils.py:193).
        Pinned buffers for speed.  Both conversions run inside the analysis / synthesis kernels: int16 samples are
        what crosses PCIe and what the kernels read and write.

        The batch is cut along TIME: slice j of every stream ([S, M, slice] -- one strided DMA request) is copied in
        while the kernels work on slice j - 1 with the recursive state carried on the device (chunked streaming is
        bit-identical to one call) and slice j - 2 is copied out, on three CUDA streams.  Every kernel launch sees the
        whole batch (full occupancy) and the pipeline drains in the time of one short last slice, so the call runs at
        the pace of the host-to-device copy.  Each call is a fresh utterance (state reset)."""
        t = L.require_cuda()
        if x_host.dim() != 3 or x_host.dtype not in (t.float32, t.int16) or not x_host.is_contiguous():
            raise ValueError("x_host must be a contiguous [S, M, N] float32 or int16 tensor")
        S, M, N = x_host.shape
        if M != self.M or N % self.hop != 0 or N < self.hop:
            raise ValueError("expected [S, %d, N] with N a positive multiple of hop=%d" % (self.M, self.hop))
        if y_host is None:
            y_host = t.empty((S, N), dtype=t.float32, pin_memory=True)
        if tuple(y_host.shape) != (S, N) or y_host.dtype not in (t.float32, t.int16) or not y_host.is_contiguous():
            raise ValueError("y_host must be a contiguous [S, N] float32 or int16 tensor")
        slices = self._slices(N // self.hop, max(1, int(slice_frames)))
        n_max = max(b - a for a, b in slices) * self.hop
        hp = self._host_pipeline(S, M, n_max, x_host.dtype, y_host.dtype)
        cur = t.cuda.current_stream()
        s_in, s_out, xbuf, ybuf = hp["s_in"], hp["s_out"], hp["xbuf"], hp["ybuf"]
        ev_in, ev_done, ev_out, sub = hp["ev_in"], hp["ev_done"], hp["ev_out"], hp["sub"]
        xe, ye = x_host.element_size(), y_host.element_size()
        lib = L.lib()
        s_in.wait_stream(cur)
        sub.reset_counters()
        if sub._state is not None:
            sub._state.zero_()
        for c, (f0, f1) in enumerate(slices):
            b = c & 1
            n0, n = f0 * self.hop, (f1 - f0) * self.hop
            xs = xbuf[b].view(-1)[:S * M * n].view(S, M, n)          # a dense [S, M, n] view of the staging buffer
            ys = ybuf[b].view(-1)[:S * n].view(S, n)
            if c >= 2:
                s_in.wait_event(ev_done[b])              # kernels of slice c-2 finished reading xbuf[b]
            L.check(lib.ds_memcpy2d_async(xs.data_ptr(), n * xe, x_host.data_ptr() + n0 * xe, N * xe, n * xe, S * M, 0,
                                          s_in.cuda_stream), "ds_memcpy2d_async")
            ev_in[b].record(s_in)
            cur.wait_event(ev_in[b])
            if c >= 2:
                cur.wait_event(ev_out[b])                # D2H of slice c-2 finished reading ybuf[b]
            sub.process_device(xs, out=ys)               # state carries over from slice to slice
            ev_done[b].record(cur)
            s_out.wait_event(ev_done[b])
            L.check(lib.ds_memcpy2d_async(y_host.data_ptr() + n0 * ye, N * ye, ys.data_ptr(), n * ye, n * ye, S, 1,
                                          s_out.cuda_stream), "ds_memcpy2d_async")
            ev_out[b].record(s_out)
        s_out.synchronize()
        cur.synchronize()
        return y_host

    def reset_counters(self):
        self.frm_cnt, self.ell = 0, 1


class MaskBeamformer(object):
    """Whole-utterance mask-based MVDR / GEV beamformer, batched over streams -- the compositions of
    example/mvdr.ipynb cell 6 ("mvdr") and cell 8 ("gev"), for which the reference has no function:

        D = Transform(n_fft, hop, channel=M).stft(x)
        p[:, n] = estimator.estimation(D[:, n, :])                     (cell 4; here McSppBase, any M <= 8,
                                                                        or a mask handed in by the caller)
        Phi_xx = sum_n p y y^H ; Phi_vv = sum_n (1 - p) y y^H          (cell 6)
        mvdr: w = compute_mvdr_weight(steering(Phi_xx), inv(Phi_vv))   (cell 6)
        gev:  w = blind_analytic_normalization(phase_correction(get_gev_vector(Phi_xx, Phi_vv)), Phi_vv)   (cell 8)
        Y = einsum('inj,ij->in', D, w.conj()) ; y = Transform(channel=1).istft(Y)

    Every stage is a CUDA kernel (transform.cu, mcspp.cu, eig.cu); ``w`` [S, K, M] and ``p`` [S, T, K] of the
    last call stay on the device as attributes.  Each call is one utterance (no state is carried over)."""

    def __init__(self, n_mics, n_fft=512, hop=256, method="mvdr", fft_precision="fp32", ban_eps=0.0):
        if method not in ("mvdr", "gev"):
            raise ValueError("method must be 'mvdr' or 'gev'")
        self.M, self.n_fft, self.hop, self.method = int(n_mics), int(n_fft), int(hop), method
        self.K = self.n_fft // 2 + 1
        self.fft_fp64 = fft_precision == "fp64"
        self.ban_eps = float(ban_eps)
        self.window = _sqrt_hann(self.n_fft)
        self.W0 = float(np.sum(self.window ** 2))
        self.w = None
        self.p = None
        self.Phi_xx = None
        self.Phi_vv = None

    def weights_device(self, Pxx, Pvv):
        """Phi_xx, Phi_vv [S, K, M, M] complex128 CUDA -> w [S, K, M] complex128 CUDA."""
        from .beamformer import beamformer as B
        t = L.require_cuda()
        S, K, M, _ = Pxx.shape
        if self.method == "mvdr":
            a = B.steering(Pxx)
            w = t.empty((S, K, M), dtype=t.complex128, device="cuda")
            L.check(L.lib().ds_mvdr_from_cov_run(S * K, M, L.ptr(a), L.ptr(Pvv), L.ptr(w), L.stream_ptr()),
                    "ds_mvdr_from_cov_run")
            return w
        w = B.get_gev_vector(Pxx, Pvv)
        w = B.phase_correction(w)
        return B.blind_analytic_normalization(w, Pvv, eps=self.ban_eps)

    def _mask_device(self, X):
        """Posterior speech-presence probability of McSppBase (defaults of mcspp_base.py:29-77) for every frame and bin:
        X [S, T, M, K] complex64 CUDA -> p [S, T, K] float64 CUDA.  Runs the output-only per-bin kernel with its p tap;
        the MVDR output it computes on the side (towards a dummy steering vector) is discarded."""
        t = L.require_cuda()
        S, T, M, K = X.shape
        prm = L.McsppParams()
        L.lib().ds_mcspp_default_params(C.byref(prm), self.n_fft, S, M, T)
        prm.frm_cnt, prm.ell, prm.mcra_L, prm.full_state = 0, 1, 15, 0
        state = t.zeros(L.lib().ds_mcspp_state_bytes(C.byref(prm)), dtype=t.uint8, device="cuda")
        a0 = t.ones((M, K), dtype=t.complex128, device="cuda")
        Y = t.empty((S, T, K), dtype=t.complex64, device="cuda")
        p = t.empty((S, T, K), dtype=t.float64, device="cuda")
        taps = L.McsppTaps(p.data_ptr(), None, None, None, None, None, None, None)
        L.check(L.lib().ds_mcspp_run(C.byref(prm), L.ptr(state), L.ptr(a0), L.ptr(X), 0, L.ptr(Y), 1, C.byref(taps),
                                     L.stream_ptr()), "ds_mcspp_run")
        return p

    def process_device(self, x_dev, p_dev=None):
        """x_dev [S, M, N] float32 CUDA (N a multiple of hop), p_dev [S, T, K] float64 CUDA or None
        -> y [S, N] float32 CUDA."""
        from .beamformer import beamformer as B
        from .transform.transform import stft_device, istft_device
        t = L.require_cuda()
        S, M, N = x_dev.shape
        if M != self.M or N % self.hop != 0 or N < self.hop:
            raise ValueError("expected [S, %d, N] with N a positive multiple of hop=%d" % (self.M, self.hop))
        wdev = L.device_window(self.window, self.n_fft)
        hist = t.zeros((S, M, self.n_fft - self.hop), dtype=t.float32, device="cuda")
        X = stft_device(x_dev.contiguous(), self.n_fft, self.hop, wdev, L.DS_STFT_STREAMING, history=hist,
                        fft_fp64=self.fft_fp64)                                   # [S, T, M, K] complex64
        T = X.shape[1]
        if p_dev is None:
            p_dev = self._mask_device(X)                                          # [S, T, K]
        elif tuple(p_dev.shape) != (S, T, self.K):
            raise ValueError("mask must be [S, T, K] = %s, got %s" % ((S, T, self.K), tuple(p_dev.shape)))
        self.p = p_dev
        self.Phi_xx, self.Phi_vv = B.masked_covariances_device(X, p_dev.contiguous())
        self.w = self.weights_device(self.Phi_xx, self.Phi_vv)
        Y = t.empty((S, T, self.K), dtype=t.complex128, device="cuda")
        L.check(L.lib().ds_apply_stream_weights_run(S, T, M, self.K, L.ptr(X), 0, L.ptr(self.w), L.ptr(Y), L.stream_ptr()),
                "ds_apply_stream_weights_run")
        tail = t.zeros((S, 1, self.n_fft - self.hop), dtype=t.float32, device="cuda")
        y = istft_device(Y[:, :, None, :], self.n_fft, self.hop, wdev, L.DS_STFT_STREAMING, tail=tail,
                         scale=self.hop / self.W0, fft_fp64=self.fft_fp64)        # [S, 1, N]
        return y[:, 0, :]

    def process(self, x, p=None):
        """Reference-shaped call: x [N, M] (or [S, N, M]), optional mask p [K, T] (or [S, K, T]) -> y [N] (or [S, N])."""
        t = L.require_cuda()
        as_torch = isinstance(x, t.Tensor)
        xd = L.to_device(x, t.float32)
        batched = xd.dim() == 3
        if not batched:
            xd = xd[None]
        pd = None
        if p is not None:
            pd = L.to_device(p, t.float64)
            pd = (pd if batched else pd[None]).permute(0, 2, 1).contiguous()
        y = self.process_device(xd.permute(0, 2, 1).contiguous(), pd)
        if not batched:
            y = y[0]
        return y if as_torch else y.double().cpu().numpy()
