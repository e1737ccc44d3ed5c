"""TEST / BASELINE INFRASTRUCTURE ONLY -- the CPU legs of bench.py.

Per configuration of BASELINE.json: the synthetic input of one stream (SURVEY.md 8d recipe), the reference's CPU
implementation of the path, and a multi-process timing driver (one worker per host core, one stream per worker and
step, OMP/BLAS threads pinned to 1).  Two kinds of runner:

  "reference"  the reference's own classes, imported through oracle/ref_harness.py from /root/reference (build
               container) or from the byte-compiled copy under oracle/_ref (GPU box; oracle/build_ref.py)
  "port"       the NumPy restatement in oracle/np_oracle.py (bit-identical output for configs 1-4, faster than the
               reference because its per-bin Python loops are vectorised) -- used where the reference cannot be
               imported, and for config 5, whose reference loop (360 x T Python iterations, azimuth only) has no
               elevation grid

Nothing under distantspeech_b200/ imports this module.
"""
import os
import time

import numpy as np

FS = 16000

CONFIGS = {
    1: dict(array="linear", r=0.032, M=4, n_fft=512, hop=256, fs=16000, look=(60.0, 0.0), interf=(140.0, 0.0)),
    2: dict(array="circular", r=0.05, M=8, n_fft=512, hop=256, fs=16000, look=(30.0, 0.0), interf=(200.0, 0.0)),
    3: dict(array="linear", r=0.05, M=6, n_fft=256, hop=256, fs=16000, look=(90.0, 0.0), interf=(20.0, 0.0)),
    4: dict(array="circular", r=0.05, M=8, n_fft=512, hop=256, fs=16000, look=(30.0, 0.0), interf=(200.0, 0.0)),
    5: dict(array="circular", r=0.05, M=16, n_fft=1024, hop=512, fs=48000, look=(100.0, 20.0), interf=(300.0, 5.0)),
}


def geometry(config):
    from oracle import np_oracle as O
    c = CONFIGS[config]
    return O.MicGeometry(c["array"], r=c["r"], M=c["M"], n_fft=c["n_fft"], fs=c["fs"])


def make_input(config, stream, seconds):
    """[M, N] float32, N a multiple of the hop (seed 0x5EED + stream index)."""
    from oracle import np_oracle as O
    c = CONFIGS[config]
    n = int(seconds * c["fs"]) // c["hop"] * c["hop"]
    return O.synth_streams(1, geometry(config), n, look_deg=c["look"], interf_deg=c["interf"], first_stream=stream, fs=c["fs"])[0]


def reference_importable():
    """True when the reference's classes really import here (source tree or the byte-compiled copy)."""
    try:
        from oracle import ref_harness as H
        if not H.reference_available():
            return False
        H.install()
        from DistantSpeech.noise_estimation.mcspp_base import McSppBase  # noqa: F401
        from DistantSpeech.transform.transform import Transform  # noqa: F401
        return True
    except Exception:
        return False


def _quiet():
    import contextlib
    import io
    return contextlib.redirect_stdout(io.StringIO())


def run_port(config, x_mn):
    """NumPy restatement; x_mn [M, N] float64 -> output."""
    from oracle import np_oracle as O
    c, geo = CONFIGS[config], geometry(config)
    if config == 1:
        with np.errstate(all="ignore"):
            return O.adaptive_mvdr(x_mn, geo, np.array(c["look"]) / 180 * np.pi, c["n_fft"], c["hop"])
    if config == 2:
        return O.fixed_beamform(x_mn.T, O.fixed_weights(geo, c["n_fft"], c["look"], "SD"), c["n_fft"], c["hop"])
    if config == 3:
        return O.FdgscOracle(geo, 256, np.array(c["look"]) / 180 * np.pi).process(x_mn.T.copy())[0]
    if config == 4:
        return O.mvdr_mcspp_chain(x_mn.T, geo, c["look"], c["n_fft"], c["hop"])
    if config == 5:
        # srp.compute_angle_spectrum's arithmetic (doa/srp.py:45-51) over a 360 x 90 grid of compute_tau([az, el])
        # delays, vectorised over frames (allowed for the timing baseline); bounded to the directions
        # the caller passes through the module-level SRP_DIRECTIONS
        Y = O.Transform(channel=c["M"], n_fft=c["n_fft"], hop_length=c["hop"]).stft(x_mn.T)
        return O.srp_map(Y, geo.omega, SRP_TAU)
    raise ValueError(config)


SRP_TAU = None      # [D, M] delays of the direction subset timed on the CPU (set by srp_prepare)


def srp_prepare(n_dirs, seed=0):
    """Random subset of the 360 x 90 grid for the CPU baseline of config 5 (the full grid takes minutes per frame)."""
    global SRP_TAU
    from oracle import np_oracle as O
    geo = geometry(5)
    rng = np.random.default_rng(seed)
    idx = rng.choice(360 * 90, n_dirs, replace=False)
    SRP_TAU = np.stack([O.method_tau(geo, np.array([i // 90, i % 90]) * np.pi / 180)[:, 0] for i in idx])
    return idx


def run_reference(config, x_mn):
    """The reference's own classes (unmodified; harness patches of SURVEY.md 8c); x_mn [M, N] float64 -> output."""
    from oracle import ref_harness as H
    H.install()
    c = CONFIGS[config]
    from DistantSpeech.beamformer.MicArray import MicArray
    with _quiet():
        mic = MicArray(arrayType=c["array"], r=c["r"], M=c["M"], n_fft=c["n_fft"])
    if config == 1:
        ab = H.make_adaptive_mvdr(mic, c["n_fft"], c["hop"], c["n_fft"])
        with np.errstate(all="ignore"):
            return ab.process(x_mn, np.array(c["look"]) / 180 * np.pi, method=2)["data"]
    if config == 2:
        # FixedBeamformer.process (fixedbeamformer.py:167-207) with SD weights.  The subclass's constructor and its
        # compute_weights hard-code nfft = 256 for the coherence matrix (:107, :140) and raise at 512, so the object is
        # initialised by the base class and gets the base class's weights; the frame loop is the
        # reference's own process_freframe (:147-165) between its own Transform.stft / istft.
        from DistantSpeech.beamformer.beamformer import beamformer
        from DistantSpeech.beamformer.fixedbeamformer import FixedBeamformer
        fb = FixedBeamformer.__new__(FixedBeamformer)
        beamformer.__init__(fb, mic, frame_len=c["n_fft"], hop=c["hop"], nfft=c["n_fft"])
        fb.W = beamformer.compute_weights(fb, c["look"], "SD")
        D = fb.transform.stft(x_mn.T)
        Yf = np.zeros((D.shape[0], D.shape[1], 1), dtype=complex)
        for n in range(D.shape[1]):
            Yf[:, n, 0] = fb.process_freframe(D[:, n, :])
        return fb.transform.istft(Yf)
    if config == 3:
        from DistantSpeech.beamformer.FDGSC import FDGSC
        with _quiet():
            fd = FDGSC(mic, frameLen=256, angle=list(c["look"]))
            return fd.process(x_mn.T.copy(), postfilter=False, dc_notch=True)[0]
    if config == 4:
        return H.reference_chain(x_mn.T, c["array"], c["r"], c["M"], c["look"], c["n_fft"], c["hop"])
    raise ValueError("config %d has no reference runner" % config)


def runner(config, kind):
    return run_reference if kind == "reference" else run_port


def pick_kind(config):
    return "reference" if (config in (1, 2, 3, 4) and reference_importable()) else "port"


# ---------------------------------------------------------------------------------------------------
def _worker(job):
    config, kind, stream, seconds, warmup, steps, srp_dirs = job
    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "NUMBA_NUM_THREADS"):
        os.environ[v] = "1"
    try:                                            # NumPy is already loaded in the forked parent: limit its BLAS pool directly
        import threadpoolctl
        threadpoolctl.threadpool_limits(1)
    except Exception:
        pass
    if config == 5:
        srp_prepare(srp_dirs)
    x = make_input(config, stream, seconds).astype(np.float64)
    run = runner(config, kind)
    c = CONFIGS[config]
    run(config, x[:, : c["hop"] * 8].copy())                                    # import + JIT warm-up
    for _ in range(warmup):
        run(config, x.copy())
    t0 = time.perf_counter()
    for _ in range(steps):
        run(config, x.copy())
    return time.perf_counter() - t0


def time_cpu(config, seconds, steps, warmup, kind=None, cores=None, srp_dirs=64):
    """Times `steps` steps after `warmup` untimed ones; a step = every host core processing one stream of `seconds`
    audio seconds.  Returns dict(value audio-s/s, cores, kind, sample, ms_per_step, steps, warmup)."""
    import multiprocessing as mp
    kind = kind or pick_kind(config)
    if cores is None:
        cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    c = CONFIGS[config]
    n = int(seconds * c["fs"]) // c["hop"] * c["hop"]
    jobs = [(config, kind, i, seconds, warmup, steps, srp_dirs) for i in range(cores)]
    with mp.get_context("fork").Pool(cores) as pool:
        per = pool.map(_worker, jobs)
    wall = max(per)
    audio_per_step = cores * n / c["fs"]
    scale = 1.0
    what = {"reference": "the reference's own classes (unmodified, %s)", "port": "NumPy restatement oracle/np_oracle.py%s"}[kind]
    if kind == "reference":
        from oracle import ref_harness as H
        what = what % ("imported from source" if H.reference_kind() == "source" else "byte-compiled under oracle/_ref")
    else:
        what = what % ""
    if config == 5:
        scale = srp_dirs / float(360 * 90)         # the CPU times a subset of the direction grid; cost is linear in directions
        what += ", %d of 32400 directions timed and scaled linearly" % srp_dirs
    sample = "%d streams x %.2f s per step (one per core), %s" % (cores, n / c["fs"], what)
    return {"value": audio_per_step * steps / wall * scale, "unit": "audio-s/s", "cores": cores, "kind": kind, "sample": sample,
            "ms_per_step": wall / steps * 1e3, "steps": steps, "warmup": warmup, "audio_s_per_step": audio_per_step}


def parity_reference(config, x_mn):
    """Checker output for bench.py's in-run parity: always the NumPy restatement (pinned bit-exactly to the reference
    by tests/test_oracle_golden*.py and tests/test_oracle_vs_reference.py)."""
    return run_port(config, np.asarray(x_mn, dtype=np.float64))
