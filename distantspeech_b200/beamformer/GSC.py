"""Frequency-domain GSC -- drop-in for ``DistantSpeech/beamformer/GSC.py`` (GSC :27, process :174-294).

Per bin: fixed beam W = a / (a^H a), Griffiths-Jim blocking matrix, NLMS noise canceller gated by the McMcra
speech-presence probability (mu = 0.01), McMcra's gain as postfilter; STFT, the per-bin kernel (``ds_gsc_run``) and
ISTFT all run on the device.  ``process(x [M, N], angle_rad, method)`` returns ``{'data', 'WNG', 'DI', 'beampattern'}``.

Notes on the reference: ``GSC.process`` ends with ``self.transformer.istft(Y)`` on a 2-D ``[K, T]`` array (:289),
which ``Transform.istft`` reads as one frame x channels and rejects -- the call only works with the array lifted
to ``[K, T, 1]`` (what the oracle harness does); this class implements that intended behaviour.  ``process1`` is the
time-domain variant (alignment, mean beam, pairwise-difference blocking matrix, constrained FDAF canceller,
GSC.py:151-172).  WNG / DI are not built (``calcWNG`` / ``calcDI`` do not exist in the reference either).
Extension: a leading stream axis ``x [S, M, N]``.
"""
import numpy as np

from .. import _lib as L
from ..noise_estimation.mc_mcra import McMcra
from ..transform.transform import Transform, stft_device, istft_device
from .MicArray import MicArray
from .beamformer import beamformer


class GSC(beamformer):
    def __init__(self, mic_array: MicArray, frameLen=256, angle=[197, 0]):
        beamformer.__init__(self, mic_array, frame_len=frameLen)
        self.mic_array = mic_array
        self.angle = np.array(angle) / 180 * np.pi if isinstance(angle, list) else angle
        self.gamma = mic_array.gamma
        self.transformer = Transform(n_fft=self.nfft, hop_length=self.hop, channel=self.M)
        self.AlgorithmList = ['src', 'DS', 'MVDR', 'TFGSC']
        self.AlgorithmIndex = 0
        self.mc_mcra = McMcra(nfft=self.nfft, channels=self.M)
        self.spp = self.mc_mcra
        self.mu = 0.01
        self.W = np.zeros((self.M, self.half_bin), dtype=complex)
        self.BM = np.zeros((self.M, self.M - 1, self.half_bin), dtype=complex)
        self._hist = None
        self._tail = None

    def process1(self, x):
        """Time-domain GSC (GSC.py:151-172): x [samples, chs] (or [S, samples, chs]) -> output [samples]; the input is
        DC-notched in place.  Same chain as TDGSC without the MCRA gate and without the non-causal delay."""
        from .TDGSC import TDGSC
        if getattr(self, "_td", None) is None:
            self._td = TDGSC(self.mic_array, frameLen=self.frameLen, angle=self.angle)
            self._td._gated, self._td._non_causal = False, False
        return self._td.process(x)[0]

    @property
    def G(self):
        """Noise-canceller weights [M-1, K] (GSC.py:71)."""
        spp, M, K = self.spp, self.M, self.half_bin
        if spp._state is None:
            return np.zeros((M - 1, K), dtype=complex)
        off = M * (M + 1)
        blk = spp._blob()[:, off:off + 2 * (M - 1), :].cpu().numpy()
        g = blk[:, :M - 1] + 1j * blk[:, M - 1:]
        return g[0] if spp._S == 1 else g

    def process(self, x, angle, method=2, retH=False, retWNG=False, retDI=False):
        """x [M, N] (or [S, M, N]); ``angle`` = (azimuth, elevation) in RADIANS (:185)."""
        if retWNG or retDI:
            raise AttributeError("'GSC' object has no attribute 'calcWNG'")        # :275-278 call undefined methods
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
        angle = np.asarray(angle, dtype=np.float64)
        self.angle, self.AlgorithmIndex = angle, method
        tao = -1 * self.r * np.cos(angle[1]) * np.cos(angle[0] - self.gamma) / self.c
        a = np.exp(-1j * self.omega[None, :] * tao[:, None])                      # [M, K]
        self.W = a / np.sum(np.conj(a) * a, axis=0, keepdims=True)                # :219
        for i in range(M - 1):                                                    # :220-225
            self.BM[0, i, :] = a[0]
            self.BM[i + 1, i, :] = -a[i + 1]
        ov = self.nfft - self.hop
        if self._hist is None or self._hist.shape[0] != S:
            self._hist = t.zeros((S, M, ov), dtype=t.float32, device="cuda")
            self._tail = t.zeros((S, 1, ov), dtype=t.float32, device="cuda")
            self.spp.reset()
        win = L.device_window(self.transformer.window, self.nfft)
        X = stft_device(xd.contiguous(), self.nfft, self.hop, win, L.DS_STFT_STREAMING, history=self._hist)
        a_dev = t.as_tensor(np.ascontiguousarray(a)).to("cuda")
        out = self.spp._run(X, a_dev=a_dev, want_Y=True, method=method)
        y = istft_device(out["Y"][:, :, None, :], self.nfft, self.hop, win, L.DS_STFT_STREAMING, tail=self._tail,
                         scale=self.hop / self.transformer.W0)[:, 0, :]
        if not batched:
            y = y[0]
        data = y if as_torch else y.double().cpu().numpy()
        return {'data': data, 'WNG': None, 'DI': None, 'beampattern': None}
