"""TEST INFRASTRUCTURE ONLY -- import harness for the *unmodified* reference.

This file lets the reference package under ``/root/reference`` be imported in
the build container so that (a) the numpy restatement in ``oracle/np_oracle.py``
can be validated against the real thing and (b) golden vectors can be dumped to
``tests/golden`` (see ``tests/golden/make_golden.py``).  It cannot travel to the
GPU box (``/root/reference`` does not exist there); everything that runs on the
GPU box uses the committed fixtures and the numpy oracle instead.

Nothing under ``distantspeech_b200/`` may import this module.

What it does:
  1. stubs third-party modules that are absent here and carry no arithmetic
     (matplotlib, sounddevice, soundfile, pyroomacoustics, pesq, pystoi,
     pyaudio, turtle, tkinter.tix, imp, tqdm is present);
  2. installs a tiny ``librosa`` stand-in that provides only the pure-indexing
     helpers ``transform.py`` needs (pad_center, frame, valid_audio, fix_length,
     MAX_MEM_BLOCK, filters.get_window);
  3. applies three mechanical compatibility patches from the outside:
       (i)   ``np.mat = np.asmatrix``  (NumPy >= 2 dropped ``np.mat``;
             used at beamformer/adaptivebeamformer.py:84);
       (ii)  ``beamformer.__init__`` accepts-and-drops ``c= fs= r=``
             (fixedbeamformer.py:99, adaptivebeamformer.py:13, postfilter.py:11
             pass them; beamformer.py:223 no longer takes them);
       (iii) ``adaptivebeamfomer.transformer.istft`` gets its 2-D ``[K,T]``
             argument lifted to ``[K,T,1]`` (adaptivebeamformer.py:122 vs
             transform.py:463-464).
The reference source files themselves are never modified or copied.
"""
from __future__ import annotations

import os
import sys
import types
from unittest import mock

_COMPILED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def _pick_root():
    """The reference's source tree when it is there (build container), else the byte-compiled copy that
    oracle/build_ref.py wrote into oracle/_ref (what travels to the GPU box: outputs only, no sources)."""
    env = os.environ.get("DS_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/DistantSpeech"):
        return "/root/reference"
    return _COMPILED_ROOT


REFERENCE_ROOT = _pick_root()

_installed = False


def reference_kind() -> str:
    """'source' (the tree under /root/reference), 'compiled' (byte code under oracle/_ref) or 'absent'."""
    base = os.path.join(REFERENCE_ROOT, "DistantSpeech", "transform", "transform")
    return "source" if os.path.exists(base + ".py") else ("compiled" if os.path.exists(base + ".refc") else "absent")


def reference_available() -> bool:
    return reference_kind() != "absent"


class _CompiledFinder(object):
    """Import hook for the byte-compiled reference: ``DistantSpeech.x.y`` -> ``<root>/DistantSpeech/x/y.refc`` (a .pyc
    file under another suffix), packages = directories (the reference has no __init__ files except noise_estimation's)."""

    def __init__(self, root):
        self.root = root

    def find_spec(self, fullname, path=None, target=None):
        import importlib.machinery
        import importlib.util
        if fullname != "DistantSpeech" and not fullname.startswith("DistantSpeech."):
            return None
        rel = os.path.join(self.root, *fullname.split("."))
        if os.path.isdir(rel):
            init = os.path.join(rel, "__init__.refc")
            spec = importlib.machinery.ModuleSpec(fullname, self, origin=init if os.path.exists(init) else None, is_package=True)
            spec.submodule_search_locations = [rel]
            return spec
        if os.path.exists(rel + ".refc"):
            return importlib.machinery.ModuleSpec(fullname, self, origin=rel + ".refc")
        return None

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        import marshal
        origin = module.__spec__.origin
        if origin is None:
            return                                   # plain directory package
        with open(origin, "rb") as fh:
            data = fh.read()
        module.__file__ = origin
        exec(marshal.loads(data[16:]), module.__dict__)      # 16-byte pyc header (PEP 552), then the code object


def _make_librosa_stub():
    import numpy as np
    import scipy.signal

    librosa = types.ModuleType("librosa")
    util = types.ModuleType("librosa.util")
    filters = types.ModuleType("librosa.filters")

    def pad_center(data, size, axis=-1, **kwargs):
        n = data.shape[axis]
        lpad = int((size - n) // 2)
        lengths = [(0, 0)] * data.ndim
        lengths[axis] = (lpad, int(size - n - lpad))
        if lpad < 0:
            raise ValueError("target size smaller than input")
        return np.pad(data, lengths, **kwargs)

    def frame(x, frame_length, hop_length, axis=-1):
        n_frames = 1 + (x.shape[-1] - frame_length) // hop_length
        idx = np.arange(frame_length)[:, None] + hop_length * np.arange(n_frames)[None, :]
        return x[idx]

    def valid_audio(y, mono=True):
        return True

    def fix_length(data, size, axis=-1, **kwargs):
        n = data.shape[axis]
        if n > size:
            sl = [slice(None)] * data.ndim
            sl[axis] = slice(0, size)
            return data[tuple(sl)]
        if n < size:
            lengths = [(0, 0)] * data.ndim
            lengths[axis] = (0, size - n)
            return np.pad(data, lengths, **kwargs)
        return data

    def get_window(window, Nx, fftbins=True):
        return scipy.signal.get_window(window, Nx, fftbins=fftbins)

    util.pad_center = pad_center
    util.frame = frame
    util.valid_audio = valid_audio
    util.fix_length = fix_length
    util.MAX_MEM_BLOCK = 2 ** 18
    filters.get_window = get_window
    librosa.__path__ = []  # behave like a package so ``import librosa.display`` resolves
    librosa.util = util
    librosa.filters = filters
    librosa.load = mock.MagicMock()
    librosa.display = mock.MagicMock(name="librosa.display")
    librosa.power_to_db = mock.MagicMock()
    return librosa, util, filters


def install():
    """Make ``import DistantSpeech...`` work in this container. Idempotent."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/ds_numba_cache")
    sys.dont_write_bytecode = True
    import warnings
    warnings.filterwarnings("ignore", category=SyntaxWarning)      # invalid escapes in the reference's plot labels

    import numpy as np

    if not hasattr(np, "mat"):
        np.mat = np.asmatrix  # patch (i)

    for name in [
        "matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.colors",
        "mpl_toolkits", "mpl_toolkits.mplot3d", "sounddevice", "soundfile",
        "pyroomacoustics", "pesq", "pystoi", "pystoi.stoi", "pyaudio", "turtle",
        "tkinter.tix", "imp", "gpuRIR", "webrtcvad",
    ]:
        if name not in sys.modules:
            sys.modules[name] = mock.MagicMock(name=name)

    if "librosa" not in sys.modules:
        librosa, util, filters = _make_librosa_stub()
        sys.modules["librosa"] = librosa
        sys.modules["librosa.util"] = util
        sys.modules["librosa.filters"] = filters
        sys.modules["librosa.display"] = librosa.display

    if reference_kind() == "source":
        if REFERENCE_ROOT not in sys.path:
            sys.path.insert(0, REFERENCE_ROOT)
    elif not any(isinstance(f, _CompiledFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _CompiledFinder(REFERENCE_ROOT))

    if reference_kind() == "compiled":
        # sourceless modules: numba cannot key an on-disk cache on a source file that is not there
        # (transform.py:224 asks for cache=True) -- same JIT, just not cached across processes
        import numba
        _jit = numba.jit
        if not getattr(numba, "_ds_nocache", False):
            numba.jit = lambda *a, **k: _jit(*a, **dict(k, cache=False))
            numba._ds_nocache = True

    # patch (ii)
    from DistantSpeech.beamformer import beamformer as _bf_mod

    _orig_init = _bf_mod.beamformer.__init__

    def _init(self, mic, frame_len=256, hop=None, nfft=None, c=None, fs=None, r=None):
        _orig_init(self, mic, frame_len=frame_len, hop=hop, nfft=nfft)

    if not getattr(_bf_mod.beamformer, "_ds_patched", False):
        _bf_mod.beamformer.__init__ = _init
        _bf_mod.beamformer._ds_patched = True

    _installed = True


def make_adaptive_mvdr(mic, frameLen, hop, nfft):
    """Construct the reference ``adaptivebeamfomer`` with patch (iii) applied."""
    install()
    from DistantSpeech.beamformer.adaptivebeamformer import adaptivebeamfomer

    obj = adaptivebeamfomer(mic, frameLen=frameLen, hop=hop, nfft=nfft)
    orig = obj.transformer.istft
    obj.transformer.istft = lambda Y: orig(Y[..., None] if Y.ndim == 2 else Y)
    return obj


def make_mcspp(nfft=512, channels=4):
    """Construct the reference ``McSpp``.  For more than 4 channels patch (v) applies: ``McSpp.__init__`` builds its
    prior as ``McCDR(nfft=self.nfft)`` with McCDR's default ``channels=4`` (mcspp.py:54), whose PSD tracker is then
    indexed up to the real channel count and raises IndexError (coherence/BinauralEnhancement.py:21,48-51).  The
    missing argument is supplied from the outside: ``obj.mccdr = McCDR(nfft, channels=channels)``.  Nothing else is
    touched; with 4 channels the object is exactly what the reference builds."""
    install()
    import contextlib
    import io
    from DistantSpeech.noise_estimation.mcspp import McSpp
    from DistantSpeech.noise_estimation.mccdr import McCDR
    with contextlib.redirect_stdout(io.StringIO()):
        obj = McSpp(nfft=nfft, channels=channels)
        if channels != 4:
            obj.mccdr = McCDR(nfft=nfft, channels=channels)
    return obj


def make_wpe(channels=2, filter_len=2, num_bands=512, delay=4, hop_length=None):
    """Make the reference ``Wpe`` (dereverberation/awpe.py:28-193) executable so that the arithmetic of its ``update``
    (:152-187) can be pinned.  As shipped it cannot run; three things are supplied from the outside, the body of
    ``update`` is the reference's own:
      (vi)   ``Subband`` (the Nyquist(M) filter bank, outside the hot path; its designer needs ``np.float_`` and writes
             pickles under /home/wangwei) is replaced, for this class only, by the reference's own streaming ``Transform``;
      (vii)  the method ``check_input_data`` that ``update`` calls (:150) exists nowhere in the reference; the injected one
             does what its siblings' ``update_input_data`` does (SubbandAF.py:53-59): analyse both blocks, one frame each;
      (viii) ``update`` ends in ``return output, self.W`` with ``output`` unassigned (:188-191): ``step`` lets the state
             update finish, swallows that UnboundLocalError and returns the filter state (W, P, var).
    Returns an object whose ``step(x_n[hop, C]) -> (W, P, var)`` runs one reference update."""
    install()
    import numpy as np
    import DistantSpeech.dereverberation.awpe as awpe_mod
    from DistantSpeech.transform.transform import Transform

    class _Bank(Transform):                                   # (vi)
        def __init__(self, n_fft=256, hop_length=128, channel=1):
            Transform.__init__(self, n_fft=n_fft, hop_length=hop_length, channel=channel)

    saved = awpe_mod.Subband
    awpe_mod.Subband = _Bank
    try:
        obj = awpe_mod.Wpe(channels=channels, filter_len=filter_len, num_bands=num_bands, delay=delay, hop_length=hop_length)
    finally:
        awpe_mod.Subband = saved

    def check_input_data(x_delayed, x_n):                     # (vii)
        Xd = obj.transform_x.stft(x_delayed)[:, 0, :]
        Dn = obj.transform_d.stft(x_n)[:, 0, :]
        return Xd, Dn
    obj.check_input_data = check_input_data

    def step(x_n):                                            # (viii)
        try:
            obj.update(np.asarray(x_n, dtype=np.float64))
        except UnboundLocalError:
            pass
        return obj.W, obj.P, obj.var
    obj.step = step
    return obj


def make_gsc(mic, frameLen=256, angle=None):
    """Construct the reference frequency-domain ``GSC`` (beamformer/GSC.py) with the same patch (iii):
    ``GSC.process`` hands its 2-D ``[K, T]`` spectrum to ``Transform.istft`` (:289), which reads 2-D as one
    frame x channels (transform.py:463-464) and asserts."""
    install()
    import contextlib
    import io
    from DistantSpeech.beamformer.GSC import GSC

    with contextlib.redirect_stdout(io.StringIO()):
        obj = GSC(mic, frameLen=frameLen) if angle is None else GSC(mic, frameLen=frameLen, angle=angle)
    orig = obj.transformer.istft
    obj.transformer.istft = lambda Y: orig(Y[..., None] if Y.ndim == 2 else Y)
    return obj


def make_subband_gsc(mic, frameLen=256, angle=None):
    """Construct the reference ``SubbandGSC`` (beamformer/SubbandGSC.py).  Patch (iv): the module cannot be
    imported as it stands -- it does ``from DistantSpeech.beamformer.FDGSC import FDGSC, DelayObj`` (:23) and
    FDGSC.py defines no ``DelayObj`` (SubbandGSC.py defines its own right below, :43) -- so the missing name is
    injected as a placeholder before the import.  Nothing else is touched."""
    install()
    import contextlib
    import io
    import DistantSpeech.beamformer.FDGSC as _fdgsc_mod

    if not hasattr(_fdgsc_mod, "DelayObj"):
        _fdgsc_mod.DelayObj = object
    from DistantSpeech.beamformer.SubbandGSC import SubbandGSC

    with contextlib.redirect_stdout(io.StringIO()):
        return SubbandGSC(mic, frameLen=frameLen) if angle is None else SubbandGSC(mic, frameLen=frameLen, angle=angle)


def reference_chain(x_nm, array_type="circular", r=0.05, M=8, look_angle=(30, 0), n_fft=512, hop=256):
    """The config-4 composition of SURVEY.md 8c run on the REFERENCE's own classes (example/mcsppbase.ipynb cell 3,
    example/mvdr.ipynb cell 4): Transform.stft -> McSppBase.estimation -> compute_mvdr_weight -> compute_omlsa_weight
    -> (w^H y) G -> Transform.istft.  x_nm [N, M] float64 -> y [N].  Used as the timed CPU baseline of bench.py
    (kind "reference") and to pin oracle.mvdr_mcspp_chain."""
    install()
    import contextlib
    import io
    import numpy as np
    from DistantSpeech.transform.transform import Transform
    from DistantSpeech.beamformer.MicArray import MicArray
    from DistantSpeech.beamformer.beamformer import beamformer, compute_mvdr_weight
    from DistantSpeech.noise_estimation.mcspp_base import McSppBase
    with contextlib.redirect_stdout(io.StringIO()):
        mic = MicArray(arrayType=array_type, r=r, M=M, n_fft=n_fft)
    D = Transform(n_fft=n_fft, hop_length=hop, channel=M).stft(x_nm)
    est = McSppBase(nfft=n_fft, channels=M)
    a0 = beamformer(mic, frame_len=n_fft, hop=hop, nfft=n_fft).compute_steering_vector_from_doa(look_angle)
    Y = np.zeros((n_fft // 2 + 1, D.shape[1], 1), dtype=complex)
    for n in range(D.shape[1]):
        yv = D[:, n, :]
        est.estimation(yv)
        w = compute_mvdr_weight(a0, est.Phi_vv_inv)
        est.compute_omlsa_weight(est.xi, est.p)
        Y[:, n, 0] = np.einsum("ij,ij->i", w.conj(), yv) * est.G
    return Transform(n_fft=n_fft, hop_length=hop, channel=1).istft(Y)
