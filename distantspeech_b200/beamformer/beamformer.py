"""Beamformer base class and weight helpers -- drop-in for
``DistantSpeech/beamformer/beamformer.py`` (compute_mvdr_weight :133,
compute_pmwf_weight :100, update_psd :158, update_csd :182, class beamformer :218;
steering :10, blind_analytic_normalization :34, phase_correction :64, get_gev_vector :77).

Geometry, steering vectors, the one-off fixed-weight design and the plotting
diagnostics are host-side NumPy precompute;
the per-frame work -- batched MVDR / PMWF weights and the weight apply -- runs
in CUDA kernels (csrc/weights.cu).
"""
import ctypes as C
import warnings

import numpy as np

from .. import _lib as L
from ..transform.transform import Transform
from .MicArray import MicArray, compute_tau
from .gen_noise_msc import gen_noise_msc


def _c128_dev(a):
    t = L.require_cuda()
    return L.to_device(np.asarray(a, dtype=np.complex128) if not isinstance(a, t.Tensor) else a, t.complex128)


def compute_mvdr_weight(steer_vector, Rvv_inv, Gmin=0.0631, beta=1):
    """w = R^-1 a / (a^H R^-1 a), batched over bins (beamformer.py:133-155).
    steer_vector [bins, M], Rvv_inv [bins, M, M] -> [bins, M] complex."""
    t = L.require_cuda()
    as_torch = isinstance(steer_vector, t.Tensor)
    a = _c128_dev(steer_vector)
    R = _c128_dev(Rvv_inv)
    B, M = a.shape[-2], a.shape[-1]
    lead = a.shape[:-1]
    a2 = a.reshape(-1, M)
    R2 = R.reshape(-1, M, M)
    if R2.shape[0] != a2.shape[0]:
        raise ValueError("steer_vector %s and Rvv_inv %s do not match" % (tuple(a.shape), tuple(R.shape)))
    out = t.empty_like(a2)
    L.check(L.lib().ds_mvdr_weight_run(a2.shape[0], M, L.ptr(a2), L.ptr(R2), L.ptr(out), L.stream_ptr()),
            "ds_mvdr_weight_run")
    out = out.reshape(*lead, M)
    return out if as_torch else out.cpu().numpy().squeeze()


def compute_pmwf_weight(xi, Rxx, Rvv_inv, Gmin=0.0631, beta=1):
    """w = (Rvv^-1 Rxx) u_1 / (beta + xi)  (beamformer.py:100-130).
    xi [bins], Rxx / Rvv_inv [bins, M, M] -> [bins, M] complex."""
    t = L.require_cuda()
    as_torch = isinstance(xi, t.Tensor)
    X = _c128_dev(Rxx)
    R = _c128_dev(Rvv_inv)
    xid = L.to_device(xi, t.float64)
    B, M = X.shape[0], X.shape[-1]
    out = t.empty((B, M), dtype=t.complex128, device="cuda")
    L.check(L.lib().ds_pmwf_weight_run(B, M, L.ptr(xid), L.ptr(X), L.ptr(R), float(beta), L.ptr(out), L.stream_ptr()),
            "ds_pmwf_weight_run")
    return out if as_torch else out.cpu().numpy().squeeze()


def apply_weights(W, X):
    """Y[..., k] = sum_m conj(W[k, m]) X[k, (t,) m]  -- einsum('ij,ij->i', W.conj(), X_n)
    (fixedbeamformer.py:163).  X [K, M] or [K, T, M] -> [K] or [K, T]."""
    t = L.require_cuda()
    as_torch = isinstance(X, t.Tensor)
    Wd = _c128_dev(W)
    Xd = _c128_dev(X)
    single = Xd.dim() == 2
    if single:
        Xd = Xd[:, None, :]
    K, T, M = Xd.shape
    Xl = Xd.permute(1, 2, 0).contiguous()[None]            # [1, T, M, K]
    Y = t.empty((1, T, K), dtype=t.complex128, device="cuda")
    L.check(L.lib().ds_apply_weights_run(1, T, M, K, L.ptr(Xl), 1, L.ptr(Wd), L.ptr(Y), L.stream_ptr()),
            "ds_apply_weights_run")
    Y = Y[0].permute(1, 0)
    if single:
        Y = Y[:, 0]
    return Y if as_torch else Y.cpu().numpy()


def update_psd(Z, Pxii, alpha=0.8):
    """Recursive auto-PSD (beamformer.py:158-179); notebook helper, host NumPy."""
    return alpha * Pxii + (1 - alpha) * np.real(Z * Z.conj())


def update_csd(Z, Pxij, alpha=0.8):
    """Recursive cross-PSD over mic pairs (beamformer.py:182-215); notebook helper, host NumPy."""
    t = 0
    M = Z.shape[1]
    for i in range(0, M - 1):
        for j in range(i + 1, M):
            Pxij[:, t] = alpha * Pxij[:, t] + (1 - alpha) * (Z[:, i] * Z[:, j].conj())
            t = t + 1
    return Pxij


# ---------------------------------------------------------------------------
# data-driven steering, GEV weights (example/mvdr.ipynb cells 2-8)
# ---------------------------------------------------------------------------
def steering(XXs):
    """Steering vector (rank 1) = principal eigenvector of the spatial correlation matrix, phase
    referenced to sensor 0 (beamformer.py:10-31).  XXs [bins, M, M] (any leading axes) -> [bins, M].
    Like ``np.linalg.eigh`` only the lower triangle is read."""
    t = L.require_cuda()
    as_torch = isinstance(XXs, t.Tensor)
    X = _c128_dev(XXs)
    M = X.shape[-1]
    if X.dim() < 2 or X.shape[-2] != M:
        raise ValueError("steering: expected [..., M, M], got %s" % (tuple(X.shape),))
    lead = X.shape[:-2]
    X2 = X.reshape(-1, M, M)
    out = t.empty((X2.shape[0], M), dtype=t.complex128, device="cuda")
    L.check(L.lib().ds_steering_run(X2.shape[0], M, L.ptr(X2), L.ptr(out), L.stream_ptr()), "ds_steering_run")
    out = out.reshape(*lead, M)
    return out if as_torch else out.cpu().numpy()


def get_gev_vector(target_psd_matrix, noise_psd_matrix):
    """GEV beamforming vectors (beamformer.py:77-97): last generalised eigenvector of
    (target, noise) per bin, scipy.linalg.eigh normalisation (w^H noise w = 1).  [bins, M, M] x2
    (any leading axes) -> [bins, M].  The phase LAPACK leaves open is fixed here (see
    include/ds_b200.h, ds_gev_run): equal to the reference up to a sign per bin."""
    t = L.require_cuda()
    as_torch = isinstance(target_psd_matrix, t.Tensor)
    A = _c128_dev(target_psd_matrix)
    B = _c128_dev(noise_psd_matrix)
    if A.shape != B.shape or A.dim() < 2 or A.shape[-1] != A.shape[-2]:
        raise ValueError("get_gev_vector: shapes %s / %s" % (tuple(A.shape), tuple(B.shape)))
    M = A.shape[-1]
    lead = A.shape[:-2]
    A2, B2 = A.reshape(-1, M, M), B.reshape(-1, M, M)
    out = t.empty((A2.shape[0], M), dtype=t.complex128, device="cuda")
    L.check(L.lib().ds_gev_run(A2.shape[0], M, L.ptr(A2), L.ptr(B2), L.ptr(out), L.stream_ptr()), "ds_gev_run")
    out = out.reshape(*lead, M)
    return out if as_torch else out.cpu().numpy()


def phase_correction(vector):
    """Rotate each bin's beamforming vector onto the previous (already corrected) one
    (beamformer.py:64-74).  vector [bins, M] (extension: [S, bins, M]); returns a copy."""
    t = L.require_cuda()
    as_torch = isinstance(vector, t.Tensor)
    w = _c128_dev(vector).clone()
    if w.dim() not in (2, 3):
        raise ValueError("phase_correction: expected [bins, M] or [S, bins, M]")
    S = 1 if w.dim() == 2 else w.shape[0]
    F, D = w.shape[-2], w.shape[-1]
    L.check(L.lib().ds_phase_correction_run(S, F, D, L.ptr(w), L.stream_ptr()), "ds_phase_correction_run")
    return w if as_torch else w.cpu().numpy()


def blind_analytic_normalization(vector, noise_psd_matrix, eps=0):
    """BAN post-scaling of a beamforming vector (beamformer.py:34-61):
    vector * |sqrt(v^H N N v)| / (|v^H N v| + eps).  vector [..., M], noise [..., M, M]."""
    t = L.require_cuda()
    as_torch = isinstance(vector, t.Tensor)
    v = _c128_dev(vector)
    N = _c128_dev(noise_psd_matrix)
    M = v.shape[-1]
    if N.shape[-2:] != (M, M) or N.shape[:-2] != v.shape[:-1]:
        raise ValueError("blind_analytic_normalization: shapes %s / %s" % (tuple(v.shape), tuple(N.shape)))
    lead = v.shape[:-1]
    v2, N2 = v.reshape(-1, M), N.reshape(-1, M, M)
    out = t.empty_like(v2)
    L.check(L.lib().ds_ban_run(v2.shape[0], M, L.ptr(v2), L.ptr(N2), float(eps), L.ptr(out), L.stream_ptr()), "ds_ban_run")
    out = out.reshape(*lead, M)
    return out if as_torch else out.cpu().numpy()


def masked_covariances(D, p=None, frames=None, scale=1.0):
    """Mask-weighted spatial covariances of example/mvdr.ipynb cell 6 (and the frame-range averages
    of cell 2): Phi_xx = scale * sum_n p[:, n] y_n y_n^H, Phi_vv = scale * sum_n (1 - p[:, n]) y_n y_n^H.
    D [K, T, M] (or [S, K, T, M]) complex, p [K, T] (or [S, K, T]) or None (then only Phi_xx, weight 1),
    frames = (start, stop) restricts the sum.  Returns (Phi_xx, Phi_vv) [K, M, M] (Phi_vv None without p)."""
    t = L.require_cuda()
    as_torch = isinstance(D, t.Tensor)
    Dd = D.to("cuda") if as_torch else t.as_tensor(np.ascontiguousarray(D)).to("cuda")
    if Dd.dtype not in (t.complex64, t.complex128):
        Dd = Dd.to(t.complex128)
    batched = Dd.dim() == 4
    if not batched:
        Dd = Dd[None]
    Xd = Dd.permute(0, 2, 3, 1).contiguous()                                  # [S, T, M, K]
    pd = None
    if p is not None:
        pd = L.to_device(p, t.float64)
        pd = (pd if batched else pd[None]).permute(0, 2, 1).contiguous()      # [S, T, K]
    out = masked_covariances_device(Xd, pd, frames=frames, scale=scale)
    res = [None if o is None else (o if batched else o[0]) for o in out]
    return tuple(res) if as_torch else tuple(None if o is None else o.cpu().numpy() for o in res)


def masked_covariances_device(Xd, pd=None, frames=None, scale=1.0):
    """Device-layout variant: Xd [S, T, M, K] complex64/complex128 CUDA, pd [S, T, K] float64 CUDA or None
    -> (Phi_xx, Phi_vv) [S, K, M, M] complex128 CUDA."""
    t = L.require_cuda()
    S, T, M, K = Xd.shape
    t0, t1 = (0, T) if frames is None else (int(frames[0]), int(frames[1]))
    Pxx = t.zeros((S, K, M, M), dtype=t.complex128, device="cuda")
    Pvv = t.zeros((S, K, M, M), dtype=t.complex128, device="cuda") if pd is not None else None
    L.check(L.lib().ds_masked_cov_run(S, T, M, K, t0, t1, L.ptr(Xd), int(Xd.dtype == t.complex128), L.ptr(pd),
                                      float(scale), L.ptr(Pxx), L.ptr(Pvv), L.stream_ptr()), "ds_masked_cov_run")
    return Pxx, Pvv


class beamformer(object):
    """beamformer base class (beamformer.py:218-534).

    ``c``, ``r`` and ``fs`` always come from the MicArray (quirk 7); the keyword
    arguments of the same name that the subclasses pass are accepted and ignored.
    """

    def __init__(self, mic: MicArray, frame_len=256, hop=None, nfft=None, c=None, fs=None, r=None):
        self.MicArray = mic
        self.M = mic.M
        self.frameLen = frame_len
        self.hop = int(frame_len // 2) if hop is None else int(hop)
        self.overlap = frame_len - self.hop
        self.nfft = int(frame_len) if nfft is None else int(nfft)
        self.c = self.MicArray.c
        self.r = self.MicArray.r
        self.fs = self.MicArray.fs
        self.half_bin = round(self.nfft / 2 + 1)
        self.freq_bin = np.linspace(0, self.half_bin - 1, self.half_bin)
        self.omega = 2 * np.pi * self.freq_bin * self.fs / self.nfft
        eye = np.eye(self.M, dtype=complex)[:, :, None]
        self.Ryy = np.repeat(eye, self.half_bin, axis=2)
        self.Rss = np.repeat(eye, self.half_bin, axis=2)
        self.Rnn = np.repeat(eye, self.half_bin, axis=2)
        self.W = np.zeros((self.half_bin, self.M), dtype=complex)
        self.Fvv = gen_noise_msc(mic=self.MicArray, nfft=self.nfft)
        self.transformer = Transform(n_fft=self.nfft, hop_length=self.hop, channel=self.M)
        self.transform = Transform(n_fft=self.nfft, hop_length=self.hop, channel=self.M)

    def compute_steering_vector_from_doa(self, look_angle=(0, 0)):
        """a0[k, m] = exp(-j w_k tau_m), look_angle in degrees (beamformer.py:267-289).
        Like the reference only the first ``MicArray.half_bin`` bins are filled (:286)."""
        mic_array = self.MicArray
        look_angle_rad = np.array(look_angle) / 180 * np.pi
        tau0 = compute_tau(mic_array, look_angle_rad)
        a0 = np.zeros((self.half_bin, self.M), dtype=complex)
        kk = min(mic_array.half_bin, self.half_bin)
        a0[:kk, :] = np.exp(-1j * self.omega[:kk, None] * tau0[:, 0][None, :])
        return a0

    def get_covariance(self):
        pass

    def get_covariance_yy(self, z, alpha=0.92):
        """Reference helper (beamformer.py:294-304): note it adds the scalar inner
        product z^H z to every element, exactly like the reference."""
        for k in range(self.half_bin):
            self.Ryy[:, :, k] = alpha * self.Ryy[:, :, k] + (1 - alpha) * (z[k, :].conj().T @ z[k, :])
        return self.Ryy

    def getweights(self, a, weightType="DS", Rvv=None, Rvv_inv=None, Ryy=None, Diagonal=1e-3):
        """Single-bin weights (beamformer.py:306-336); a [M, 1].  Per-bin scalar helper
        kept on the host: the per-frame pipelines compute these inside their kernels."""
        a = np.asarray(a)
        if Rvv is None:
            warnings.warn("Rvv not provided,using eye(M,M)\n")
            Rvv = np.eye(self.M)
        if Rvv_inv is None:
            Fvv_k_inv = np.linalg.inv(Rvv + Diagonal * np.eye(self.M))
        else:
            Fvv_k_inv = Rvv_inv
        if weightType == "src":
            weights = a
            weights[1:] = 0
        elif weightType == "DS":
            weights = a / self.M
        elif weightType == "MVDR":
            weights = Fvv_k_inv @ a / (a.conj().T @ Fvv_k_inv @ a)
        elif weightType == "TFGSC":
            u = np.zeros((self.M, 1))
            u[0] = 1
            temp = Fvv_k_inv @ Ryy
            weights = (temp - np.eye(self.M)) @ u / (np.trace(temp) - self.M)
        else:
            raise ValueError("Unknown beamformer weights: %s" % weightType)
        return weights

    def compute_weights(self, look_angle=[90, 0], weightType="DS", diag_value=1e-3):
        """DS: a0 / M; SD: mvdr(a0, inv(Gamma + diag I))  (beamformer.py:338-373).  One-off
        design per look angle on the host."""
        a0 = self.compute_steering_vector_from_doa(look_angle=look_angle)
        if weightType == 'DS':
            W = a0 / self.M
        elif weightType == 'SD':
            Rinv = np.linalg.inv(self.Fvv + np.eye(self.M) * diag_value)
            num = Rinv @ a0[..., None]
            W = (num / (a0[:, None, :].conj() @ num)).squeeze()
        else:
            raise UnboundLocalError("weightType must be 'DS' or 'SD'")   # reference leaves W unbound
        return W

    def process_freframe(self, X_n):
        """The base-class version is broken in the reference (quirk 6: beamformer.py:390-392
        raises ValueError); subclasses override it."""
        raise ValueError("beamformer.process_freframe is not usable in the reference either "
                         "(beamformer.py:390 unpacks np.zeros(X_n.shape)); use a subclass")

    def process(self, x):
        assert x.shape[1] >= 2
        D = self.transform.stft(x)
        half_bin, frameNum, channel = D.shape
        Yf = np.zeros((half_bin, frameNum, 1), dtype=complex)
        for n in range(frameNum):
            Yf[:, n, 0] = self.process_freframe(D[:, n, :])
        output = self.transform.istft(Yf)
        assert output.shape[0] == x.shape[0]
        return output.squeeze()

    # ---- diagnostics (notebook plotting only, host NumPy) ----------------
    def compute_array_gain(self, weights, steer_vector, Rvv, return_db=False):
        """|w^H a|^2 / |w^H Rvv w| (beamformer.py:435-464).  Like the reference the result is **[bins, bins]**, not the
        documented [bins]: the numerator is lifted to [K, 1] and the quadratic form stays [K, 1, 1] (:456-458), so
        NumPy broadcasts to G[i, j] = |num_j|^2 / |den_i| -- the per-bin array gain is the diagonal."""
        num = np.einsum('ij, ij->i', weights.conj(), steer_vector)
        den = weights[:, np.newaxis, :].conj() @ Rvv @ weights[..., None]          # [K, 1, 1]
        G = np.abs(num[..., None]) ** 2 / np.abs(den)                               # [K, K, 1]
        if return_db:
            G = 10 * np.log10(G + 1e-6)
        return G.squeeze()

    def compute_wng_di(self, weights=None, look_angle=[0, 0], return_db=True):
        if weights is None:
            weights = self.compute_weights(look_angle=look_angle)
        steer_vector = self.compute_steering_vector_from_doa(look_angle=look_angle)
        di = self.compute_array_gain(weights, steer_vector, self.Fvv)
        eye = np.broadcast_to(np.eye(self.M), (self.half_bin, self.M, self.M))
        wng = self.compute_array_gain(weights, steer_vector, eye)
        if return_db:
            wng = 10 * np.log10(wng + 1e-6)
            di = 10 * np.log10(di + 1e-6)
        return wng, di

    def compute_beampattern(self, mic_array: MicArray, weights=None, look_angle=np.array([0, 0]) / 180 * np.pi):
        """20 log10 |sum_m conj(W[m,k]) exp(-j w_k tau_m(az))| for az = 0..359 -> [360, half_bin]."""
        if weights is None:
            tau0 = compute_tau(mic_array, np.array(look_angle) / 180 * np.pi)
            H = np.exp(-1j * self.omega[None, :mic_array.half_bin] * tau0) / self.M
        else:
            H = weights
        beamout = np.zeros([360, mic_array.half_bin])
        for az in range(360):
            tau = compute_tau(mic_array, np.array([az, 0]) * np.pi / 180)
            a = np.exp(-1j * self.omega[None, :mic_array.half_bin] * tau)        # [M, K]
            beamout[az] = np.abs(np.sum(H.conj() * a, axis=0))
        return 20 * np.log10(beamout + 1e-12)
