"""SRP-PHAT direction-of-arrival map -- drop-in for ``DistantSpeech/doa/srp.py`` (srp :10,
compute_angle_spectrum :17).

``compute_angle_spectrum(x[N, M])`` returns ``(angle_spectrum[360, T], p[K, T])`` like the
reference (azimuth only, elevation 0).  ``compute_grid_spectrum`` (extension) evaluates an
arbitrary (azimuth, elevation) grid through ``MicArray.compute_tau`` -- config 5 of
BASELINE.json uses 360 x 90 directions.  The steered response is a per-bin
[D x M] . [M x T] complex contraction; it runs on the tensor cores (tcgen05, tf32) with the
steering matrix generated on chip, or on the CUDA cores (``engine="simt"``).
"""
import numpy as np

from .. import _lib as L
from ..beamformer.MicArray import MicArray
from ..noise_estimation.mcra import NoiseEstimationMCRA
from ..transform.transform import Transform, stft_device


class srp(object):
    def __init__(self, mic_array: MicArray, engine="tensor") -> None:
        self.mic_array = mic_array
        self.transform = Transform(channel=mic_array.M, n_fft=mic_array.n_fft, hop_length=int(mic_array.n_fft / 2))
        self.spp = NoiseEstimationMCRA(nfft=mic_array.n_fft)
        self.spp.L = 65
        self.engine = engine

    def _spectrum(self, x):
        """x [N, M] -> device spectrum X [T, M, K] complex64 (Transform.stft semantics)."""
        t = L.require_cuda()
        L.ensure_init()
        tf = self.transform
        xd = L.to_device(x, t.float32)
        if xd.dim() != 2 or xd.shape[1] != self.mic_array.M:
            raise ValueError("expected x [samples, %d]" % self.mic_array.M)
        tf._state(1)
        xs = xd.t().contiguous()[None]                                        # [1, M, N]
        X = stft_device(xs, tf.n_fft, tf.hop_length, L.device_window(tf.window, tf.n_fft), L.DS_STFT_STREAMING,
                        history=tf._hist)
        return X[0]

    def _steered_response(self, X, tau, phat):
        """X [T, M, K] c64 CUDA, tau [D, M] seconds -> P [D, T] float32 CUDA."""
        t = L.require_cuda()
        T, M, K = X.shape
        D = tau.shape[0]
        Yhat = t.empty((K, T, M), dtype=t.complex64, device="cuda")
        L.check(L.lib().ds_phat_run(T, M, K, int(bool(phat)), L.ptr(X), L.ptr(Yhat), L.stream_ptr()), "ds_phat_run")
        tau_d = t.as_tensor(np.ascontiguousarray(tau, dtype=np.float32)).to("cuda")
        P = t.empty((D, T), dtype=t.float32, device="cuda")
        use_tc = int(self.engine == "tensor" and M in (4, 8, 16))
        nws = L.lib().ds_srp_workspace_bytes(T, M, K, use_tc)
        ws = t.empty(max(nws, 1), dtype=t.uint8, device="cuda") if nws else None      # torch allocations are 512-byte aligned
        L.check(L.lib().ds_srp_run(D, T, M, K, float(self.mic_array.fs), int(self.mic_array.n_fft), L.ptr(tau_d),
                                   L.ptr(Yhat), L.ptr(ws), L.ptr(P), use_tc, L.stream_ptr()), "ds_srp_run")
        return P

    def _mcra_p(self, X):
        """MCRA speech presence on channel 0 (srp.py:38-40; returned, not used by the map)."""
        t = L.require_cuda()
        pw = L.spectral_power(X[:, 0, :].to(t.complex128), via_abs=True)     # [T, K]; complex input -> np.abs(y) ** 2
        _, p = self.spp.estimation_frames(pw, return_p=True)
        return p                                                              # [T, K]

    def compute_angle_spectrum(self, x, phat=True, resolution=1):
        """x [samples, chs] -> (angle_spectrum [360, n_frame], p [half_bin, n_frame])  (srp.py:17-53)."""
        X = self._spectrum(x)
        p = self._mcra_p(X)
        angles = np.arange(0, 360, resolution)
        tau = np.stack([self.mic_array.compute_tau(np.array([a, 0]) * np.pi / 180)[:, 0] for a in angles])
        P = self._steered_response(X, tau, phat).double().cpu().numpy()
        T = P.shape[1]
        out = np.zeros((360, T))
        for i, a in enumerate(angles):
            out[a: a + resolution, :] = P[i]
        return out, p.cpu().numpy().T

    def compute_grid_spectrum(self, x, az_deg, el_deg, phat=True, as_torch=False):
        """Extension: steered response over the grid az x el (degrees) -> [len(az), len(el), n_frame]."""
        X = self._spectrum(x)
        az = np.asarray(az_deg, dtype=np.float64)
        el = np.asarray(el_deg, dtype=np.float64)
        tau = np.stack([self.mic_array.compute_tau(np.array([a, e]) * np.pi / 180)[:, 0] for a in az for e in el])
        P = self._steered_response(X, tau, phat).reshape(len(az), len(el), -1)
        return P if as_torch else P.double().cpu().numpy()
