"""Online MVDR beamformer with an MCRA-VAD-gated noise covariance -- drop-in for
``DistantSpeech/beamformer/adaptivebeamformer.py`` (adaptivebeamfomer :10, process :44),
the ``example/run_MVDRbeamformer.py`` path.

``process(x[M, N], angle_rad, method=2)`` runs three kernels: Transform.stft,
ds_amvdr_run (the frame x bin loops of :69-120) and Transform.istft.  A leading
stream axis ``x[S, M, N]`` batches independent streams (``data`` is ``[S, N]``).
"""
import ctypes as C

import numpy as np

from .. import _lib as L
from ..transform.transform import stft_device, istft_device
from .beamformer import beamformer
from .MicArray import MicArray


class _Lazy(object):
    """A host array that is fetched from the device the first time it is read (a device -> host copy into pageable memory
    costs more than the whole kernel chain of a batched call, and most callers never look at these attributes)."""

    def __init__(self, fetch):
        self._fetch, self._value = fetch, None

    def get(self):
        if self._fetch is not None:
            self._value, self._fetch = self._fetch(), None
        return self._value


def _lazy_attr(name):
    def get(self):
        v = self.__dict__.get(name)
        return v.get() if isinstance(v, _Lazy) else v

    def set_(self, value):
        self.__dict__[name] = value
    return property(get, set_)


class _McraView(object):
    """The ``self.mcra`` attribute of the reference: counters on the host, p on the device (read back on first access)."""
    p = _lazy_attr("_p")

    def __init__(self):
        self.L = 15
        self.alpha_d, self.alpha_s, self.delta_s, self.alpha_p = 0.95, 0.8, 5, 0.2
        self.p_max, self.p_min = 0.999, 1e-3
        self.ell, self.frm_cnt = 1, 0
        self.p = None


class adaptivebeamfomer(beamformer):
    H = _lazy_attr("_H")          # weights of the last frame, [M, K] (or [S, M, K]); read back from the device on first access

    def __init__(self, mic: MicArray, frameLen=256, hop=None, nfft=None, c=343, r=0.032, fs=16000):
        beamformer.__init__(self, mic=mic, frame_len=frameLen, hop=hop, nfft=nfft, c=c, fs=fs)
        self.M = mic.M
        self.gamma = mic.gamma
        self.H = np.ones([self.M, self.half_bin], dtype=complex) / self.M
        self.angle = np.array([0, 0]) / 180 * np.pi
        self.method = 'MVDR'
        self.frameCount = 0
        self.calc = 0
        self.estPos = None                      # None -> MCRA-based VAD (:30)
        self.AlgorithmList = ['src', 'DS', 'MVDR', 'TFGSC']
        self.AlgorithmIndex = 0
        self.mcra = _McraView()
        self.update_noise_psd_flag = 0
        self._state = None
        self._S = None
        self._hist = None
        self._tail = None

    # ---- device state ----------------------------------------------------------------
    def _params(self, S, T, method):
        p = L.AmvdrParams()
        L.lib().ds_amvdr_default_params(C.byref(p), self.nfft, S, self.M, T)
        m = self.mcra
        p.frm_cnt, p.ell, p.mcra_L, p.method = int(m.frm_cnt), int(m.ell), int(m.L), int(method)
        p.mcra_alpha_d, p.mcra_alpha_s, p.mcra_delta_s = float(m.alpha_d), float(m.alpha_s), float(m.delta_s)
        p.mcra_alpha_p, p.mcra_p_min, p.mcra_p_max = float(m.alpha_p), float(m.p_min), float(m.p_max)
        return p

    def _ensure(self, S):
        t = L.require_cuda()
        if self._state is None or self._S != S:
            nbytes = L.lib().ds_amvdr_state_bytes(C.byref(self._params(S, 1, 2)))
            self._state = t.zeros(nbytes, dtype=t.uint8, device="cuda")
            ov = max(self.nfft - self.hop, 1)
            self._hist = t.zeros((S, self.M, ov), dtype=t.float32, device="cuda")
            self._tail = t.zeros((S, 1, ov), dtype=t.float32, device="cuda")
            self._S = S
            self.mcra.frm_cnt, self.mcra.ell = 0, 1

    def _field(self, field):
        K, M = self.half_bin, self.M
        if self._state is None:
            return np.zeros((K, M, M), dtype=complex)
        t = L.require_cuda()
        out = t.empty((self._S, K, M, M), dtype=t.complex128, device="cuda")
        L.check(L.lib().ds_amvdr_export(C.byref(self._params(self._S, 1, 2)), L.ptr(self._state), field, L.ptr(out),
                                        L.stream_ptr()), "ds_amvdr_export")
        v = out.cpu().numpy()
        return v[0] if self._S == 1 else v

    Rvv = property(lambda self: self._field(0))
    Rvv_inv = property(lambda self: self._field(1))

    @property
    def Ryy(self):
        return self._field(2)

    @Ryy.setter
    def Ryy(self, value):        # the base class assigns its own (unused) Ryy in __init__
        self._Ryy_base = value

    # ---- processing --------------------------------------------------------------------
    def process(self, x, angle, method=2, retH=False, retWNG=False, retDI=False):
        """x [M, N] (or [S, M, N]); ``angle`` = (azimuth, elevation) in RADIANS (:52).
        Returns {'data', 'WNG', 'DI', 'beampattern'} like the reference (:128)."""
        if retWNG or retDI:
            # the reference calls self.calcWNG / self.calcDI, which do not exist (:113-116)
            raise AttributeError("'adaptivebeamfomer' object has no attribute 'calcWNG'")
        if method not in (0, 1, 2, 3):
            raise IndexError("list index out of range")        # AlgorithmList[method]
        t = L.require_cuda()
        L.ensure_init()
        as_torch = isinstance(x, t.Tensor)
        xd = L.to_device(x, t.float32)
        batched = xd.dim() == 3
        if not batched:
            xd = xd[None]
        S, M, N = xd.shape
        if M != self.M:
            raise ValueError("expected %d channels, got %d" % (self.M, M))
        self._ensure(S)
        angle = np.asarray(angle, dtype=np.float64)
        self.angle = angle
        self.AlgorithmIndex = method
        # circular-array far-field delays, r / gamma / c from the MicArray (:52, quirk 7)
        tao = -1 * self.r * np.cos(angle[1]) * np.cos(angle[0] - self.gamma) / self.c
        a = np.exp(-1j * self.omega[None, :] * tao[:, None])                     # [M, K]
        a_dev = t.as_tensor(np.ascontiguousarray(a)).to("cuda")
        win = L.device_window(self.transformer.window, self.nfft)
        X = stft_device(xd.contiguous(), self.nfft, self.hop, win, L.DS_STFT_STREAMING, history=self._hist)
        T = X.shape[1]
        Y = t.empty((S, T, 1, self.half_bin), dtype=t.complex64, device="cuda")
        Hl = t.empty((S, self.half_bin, M), dtype=t.complex128, device="cuda")
        prm = self._params(S, T, method)
        # no per-frame p tap: the reference only keeps the last frame's mcra.p, which is part of the state blob
        L.check(L.lib().ds_amvdr_run(C.byref(prm), L.ptr(self._state), L.ptr(a_dev), L.ptr(X), L.ptr(Y), L.ptr(Hl),
                                     None, L.stream_ptr()), "ds_amvdr_run")
        f, e = C.c_int32(self.mcra.frm_cnt), C.c_int32(self.mcra.ell)
        L.lib().ds_mcra_advance(int(self.mcra.L), T, C.byref(f), C.byref(e))
        self.mcra.frm_cnt, self.mcra.ell = f.value, e.value
        y = istft_device(Y, self.nfft, self.hop, win, L.DS_STFT_STREAMING, tail=self._tail,
                         scale=self.hop / self.transformer.W0)[:, 0, :]

        def fetch_H(Hl=Hl, S=S):
            Hn = Hl.cpu().numpy()
            return Hn[0].T if S == 1 else Hn.transpose(0, 2, 1)                   # [M, K]

        def fetch_p(st=self._state, S=S, idx=3 * M * M + 3):
            pn = st.view(t.float64).view(S, -1, self.half_bin)[:, idx, :].cpu().numpy()           # state order: Rvv, Rinv, Ryy, mcra[S, Smin, Stmp, p, lambda]
            return pn[0] if S == 1 else pn
        self.H = _Lazy(fetch_H)
        self.mcra.p = _Lazy(fetch_p)
        beampattern = None
        if retH:
            beampattern = self.beampattern(self.omega, self.H if S == 1 else self.H[0])
        if not batched:
            y = y[0]
        data = y if as_torch else y.double().cpu().numpy()
        return {'data': data, 'WNG': None, 'DI': None, 'beampattern': beampattern}

    def beampattern(self, omega, H):
        """10 log10 |sum_m conj(H[m, k]) exp(-j w_k tao_m(az))| over az = 0..359 (beamformer.py:517-534);
        plotting diagnostic, host NumPy."""
        half_bin = H.shape[1]
        beamout = np.zeros([360, half_bin])
        for az in range(360):
            tao = -1 * self.r * np.cos(0) * np.cos(az * np.pi / 180 - self.gamma) / self.c
            a = np.exp(-1j * omega[None, :half_bin] * tao[:, None])
            beamout[az] = np.abs(np.sum(H.conj() * a, axis=0))
        with np.errstate(divide="ignore"):
            return 10 * np.log10(beamout)
