"""Zelinski / McCowan coherence postfilter -- drop-in for
``DistantSpeech/postfilter/postfilter.py`` (PostFilter :8, update_CSD_PSD :19, getweights :45).
Channel-major layout ``Z[M, K]`` and the hard-coded 129 bins (nfft = 256) are the reference's."""
import numpy as np

from .. import _lib as L
from ..beamformer.fixedbeamformer import FixedBeamformer


class PostFilter(FixedBeamformer):
    def __init__(self, MicArray, frameLen=256, hop=None, nfft=None, c=343, r=0.032, fs=16000):
        FixedBeamformer.__init__(self, MicArray, frameLen=frameLen, hop=hop, nfft=nfft, c=c, r=r, fs=fs)
        self.M = MicArray.M
        self.half_bin = 129                                   # hard-coded in the reference (:13)
        self.NumSpec = int((self.M * self.M - self.M) / 2)
        self.H = np.ones([1, self.half_bin], dtype=complex)
        self._zstate = None

    def _ensure(self):
        t = L.require_cuda()
        if self._zstate is None:
            n = L.lib().ds_zelinski_state_bytes(1, self.M, self.half_bin)
            self._zstate = t.zeros(n, dtype=t.uint8, device="cuda")
        return t

    def _views(self):
        t = self._ensure()
        st = self._zstate.view(t.float64).view(-1, self.half_bin).cpu().numpy()
        Pxii = st[: self.M].copy()
        Pxij = st[self.M::2] + 1j * st[self.M + 1::2]
        return Pxii, Pxij

    Pxii = property(lambda self: self._views()[0])
    Pxij = property(lambda self: self._views()[1])

    def _run(self, Z, alpha):
        t = self._ensure()
        L.ensure_init()
        Zd = t.as_tensor(np.ascontiguousarray(np.asarray(Z, dtype=np.complex128))).to("cuda")
        M, K = Zd.shape
        assert K == self.half_bin and M == self.M
        Fvv = t.as_tensor(np.ascontiguousarray(self.Fvv[:K], dtype=np.float64)).to("cuda")
        W = t.empty((1, 1, K), dtype=t.float64, device="cuda")
        L.check(L.lib().ds_zelinski_run(1, 1, M, K, float(alpha), 0.7, L.ptr(self._zstate), L.ptr(Zd), L.ptr(Fvv), L.ptr(W),
                                        L.stream_ptr()), "ds_zelinski_run")
        return W[0, 0].cpu().numpy()

    def update_CSD_PSD(self, Z, alpha=0.8):
        """Recursive auto / cross PSD update (:19-43); Z [M, K] complex."""
        self._run(Z, alpha)

    def getweights(self, Z):
        """Postfilter weights W [K] for one frame (:45-84)."""
        return self._run(Z, 0.8)

    def process(self, x, DS, angle, method='DS', retH=False, retWNG=False, retDI=False):
        # the reference body uses undefined names (data_ext, last_output, method): it cannot run
        raise AttributeError("'PostFilter' object has no attribute 'data_ext' (postfilter.py:94 is dead code in the reference)")
