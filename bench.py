#!/usr/bin/env python
"""bench.py -- headline benchmark of BASELINE.json: audio-seconds per second
(x real time) of the 8-mic MVDR + McSppBase + OMLSA-postfilter chain.

    python bench.py --gpus N --steps K --warmup W          (our CUDA path, config 4 = configs[3] of BASELINE.json)
    python bench.py --impl reference ...                   (the reference's CPU path on the host cores)
    python bench.py --config {1,2,3,5} ...                 (the other BASELINE configurations, same JSON shape)

Workload of the headline (configs[3], the configuration the metric is quoted on): 1024 streams per GPU x 10 s x 8 mics
@ 16 kHz, n_fft 512 / hop 256, synthetic data of the SURVEY.md 8d recipe generated on the device (weak scaling: 8192
streams on 8 GPUs).  A step = one pass of the whole chain over the batch.  Inputs (5.2 GB per GPU) are far larger than
L2, so every step streams from HBM.  Prints ONE JSON line on rank 0; at N = 1 the default run also measures configs
1, 2, 3 and 5 (short runs, `configs` key) so that every BASELINE configuration has a clocked, parity-checked number.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FS = 16000
N_FFT, HOP, M = 512, 256, 8
LOOK, INTERF = (30.0, 0.0), (200.0, 0.0)
ALGO_BYTES_PER_AUDIO_S = M * FS * 4 + FS * 4          # fp32 in (M*fs*4) + fp32 out (fs*4) = 576000
METRIC = "audio-s/s (x realtime) for 8-mic MVDR+postfilter chain"
WORKLOADS = {
    1: "configs[0]: online MVDR (adaptivebeamfomer.process, method 2), 4-mic linear r=0.032 16 kHz, n_fft 512 hop 256",
    2: "configs[1]: fixed superdirective beamformer (FixedBeamformer.process), 8-mic circular r=0.05 16 kHz, n_fft 512 hop 256",
    3: "configs[2]: FDGSC (adaptive blocking matrix + NLMS canceller), 6-mic linear r=0.05 16 kHz, frameLen 256",
    4: "configs[3]: MVDR + McSppBase + OMLSA chain, 8-mic circular r=0.05 16 kHz, n_fft 512 hop 256",
    5: "configs[4]: SRP-PHAT over a 360x90 direction grid, 16-mic circular r=0.05 48 kHz, n_fft 1024 hop 512",
}
METRICS = {1: "audio-s/s (x realtime) for 4-mic online MVDR", 2: "audio-s/s (x realtime) for 8-mic fixed SD beamformer",
           3: "audio-s/s (x realtime) for 6-mic FDGSC", 4: METRIC, 5: "audio-s/s (x realtime) for 16-mic SRP-PHAT 360x90"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=4, choices=[1, 2, 3, 4, 5])
    ap.add_argument("--streams-per-gpu", type=int, default=0, help="0: the configuration's own size")
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--full-state", type=int, default=0)
    ap.add_argument("--fft", default="fp32", choices=["fp32", "fp64"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="headline only: skip the short runs of configs 1, 2, 3, 5")
    ap.add_argument("--cpu-seconds", type=float, default=0.0, help="audio seconds per CPU-baseline stream (0: per config)")
    ap.add_argument("--slice-frames", type=int, default=16, help="frames per time slice of the end-to-end pipeline")
    return ap.parse_args()


# --------------------------------------------------------------------------
# process group / timing helpers
# --------------------------------------------------------------------------
class Env(object):
    def __init__(self):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.torch = None
        self.dist = None

    def init_gpu(self):
        import torch
        import torch.distributed as dist
        assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            # NCCL prints its version banner on STDOUT at communicator creation when NCCL_DEBUG is VERSION or higher;
            # stdout carries exactly one JSON line, so file descriptor 1 points at stderr while NCCL initialises
            sys.stdout.flush()
            saved = os.dup(1)
            os.dup2(2, 1)
            try:
                dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
                dist.barrier()                       # forces the (lazy) communicator creation now
                torch.cuda.synchronize()
            finally:
                sys.stdout.flush()
                os.dup2(saved, 1)
                os.close(saved)
        self.torch, self.dist = torch, dist

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        from distantspeech_b200.sharding import max_over_ranks
        return max_over_ranks(v, device="cuda")


class ClockSampler(object):
    """SM clock and throttle reasons sampled DURING the timed region: an NVML polling thread (2 ms period; the
    timed region of the default run is ~0.4 s, too short for `nvidia-smi -lms`), nvidia-smi as a fallback."""
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"),
               (0x80, "hw_power_brake_slowdown"))

    def __init__(self, gpu_index):
        import threading
        self.sm, self.smax, self.reasons = [], [], set()
        self._stop = threading.Event()
        self._thr = None
        self._how = None
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = gpu_index
        if vis:
            try:
                phys = int(vis.split(",")[gpu_index])
            except Exception:
                phys = gpu_index
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.smax.append(float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)))

            def poll():
                while not self._stop.is_set():
                    try:
                        self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        mask = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                        for bit, name in self.REASONS:
                            if mask & bit:
                                self.reasons.add(name)
                    except Exception:
                        pass
                    time.sleep(0.002)
            self._thr = threading.Thread(target=poll, daemon=True)
            self._thr.start()
            self._how = "nvml"
        except Exception:
            self._how = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self._thr is not None:
            self._stop.set()
            self._thr.join(timeout=2)
        if not self.sm:                       # NVML unavailable: one nvidia-smi reading right after the region
            try:
                q = "clocks.sm,clocks.max.sm"
                r = subprocess.run(["nvidia-smi", "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=10).stdout.splitlines()[0].split(",")
                self.sm, self.smax, self._how = [float(r[0])], [float(r[1])], "nvidia-smi (after the region)"
            except Exception:
                return out
        return {"sm_mhz": statistics.median(self.sm), "sm_max_mhz": max(self.smax), "reasons": sorted(self.reasons),
                "samples": len(self.sm), "source": self._how}


def time_steps(env, step, steps, warmup):
    """W untimed warm-up steps, then exactly K timed steps bracketed by barrier + synchronize, CUDA events on the
    launching stream, max over ranks.  Returns (ms_total, clocks)."""
    t = env.torch
    for _ in range(max(warmup, 3)):
        step()
    env.barrier()
    sampler = ClockSampler(env.local_rank) if env.rank == 0 else None
    e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
    env.barrier()
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    env.barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    return env.max_over_ranks(ms), clocks


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


def measure_fp64_peak(t):
    """DFMA microbenchmark of the library (ds_fp64_peak_run), best of 3, CUDA events -> TFLOP/s (None on failure)."""
    from distantspeech_b200 import _lib as L
    try:
        scratch = t.zeros(8, dtype=t.float64, device="cuda")
        best = None
        for it in range(4):
            e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
            e0.record()
            flops = L.lib().ds_fp64_peak_run(8000, L.ptr(scratch), L.stream_ptr())
            e1.record()
            t.cuda.synchronize()
            if flops <= 0:
                return None
            tf = flops / (e0.elapsed_time(e1) / 1e3) / 1e12
            if it > 0:
                best = tf if best is None else max(best, tf)
        return best
    except Exception:
        return None


# --------------------------------------------------------------------------
# synthetic data on the device (SURVEY 8d recipe)
# --------------------------------------------------------------------------
def synth_device(torch, S, mic, n_samples, seed, look=LOOK, interf=INTERF, fs=FS, chunk=64, dtype=None):
    from distantspeech_b200.beamformer.MicArray import compute_tau
    dev = "cuda"
    Mm = mic.M
    tau_s = torch.as_tensor(compute_tau(mic, np.array(look) / 180 * np.pi)[:, 0], device=dev)
    tau_i = torch.as_tensor(compute_tau(mic, np.array(interf) / 180 * np.pi)[:, 0], device=dev)
    nfft = 1 << int(np.ceil(np.log2(n_samples + 64)))
    f = torch.fft.rfftfreq(nfft, 1.0 / fs, device=dev, dtype=torch.float64)
    t = torch.arange(n_samples, device=dev, dtype=torch.float64) / fs
    env = (torch.sin(2 * np.pi * 0.7 * t) >= 0).to(torch.float32)
    ph_s = torch.exp(-2j * np.pi * f[None, :] * tau_s[:, None]).to(torch.complex64)       # [M, F]
    ph_i = torch.exp(-2j * np.pi * f[None, :] * tau_i[:, None]).to(torch.complex64)
    x = torch.empty((S, Mm, n_samples), dtype=torch.float32, device=dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    for lo in range(0, S, chunk):
        hi = min(S, lo + chunk)
        n = hi - lo
        tgt = torch.randn((n, n_samples), generator=gen, device=dev) * env * 0.3
        itf = torch.randn((n, n_samples), generator=gen, device=dev) * 0.2
        Ft = torch.fft.rfft(tgt, nfft)
        Fi = torch.fft.rfft(itf, nfft)
        for m in range(Mm):
            d = torch.fft.irfft(Ft * ph_s[m] + Fi * ph_i[m], nfft)[:, :n_samples]
            d = d + torch.randn((n, n_samples), generator=gen, device=dev) * 0.05
            x[lo:hi, m, :] = 0.5 * d
    return x


# --------------------------------------------------------------------------
# host placement for the end-to-end path (one rank per GPU: CPUs and pinned memory next to the rank's GPU)
# --------------------------------------------------------------------------
def bind_near_gpu(torch, local_rank):
    """Binds this process's CPU affinity (and, when the kernel allows it, its memory policy) to the NUMA node the
    rank's GPU hangs off, before the pinned staging buffers are allocated (first touch).  Reports what happened."""
    info = {"gpu_numa_node": None, "cpus_visible": None, "cpus_bound": None, "mempolicy": "unchanged"}
    try:
        allowed = sorted(os.sched_getaffinity(0))
        info["cpus_visible"] = len(allowed)
        prop = torch.cuda.get_device_properties(local_rank)
        bus = "%04x:%02x:%02x.0" % (getattr(prop, "pci_domain_id", 0), prop.pci_bus_id, prop.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip())
        info["gpu_numa_node"] = node
        if node < 0:
            return info
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        near = sorted(cpus & set(allowed))
        if near:
            os.sched_setaffinity(0, near)
            info["cpus_bound"] = len(near)
        else:
            info["cpus_bound"] = 0                                      # the container's cpuset has no CPU on that node
        try:                                                            # set_mempolicy(MPOL_PREFERRED, node): x86-64 syscall 238
            libc = ctypes.CDLL(None, use_errno=True)
            mask = ctypes.c_ulong(1 << node)
            rc = libc.syscall(238, 1, ctypes.byref(mask), ctypes.c_ulong(64))
            info["mempolicy"] = "preferred node %d" % node if rc == 0 else "set_mempolicy refused (errno %d)" % ctypes.get_errno()
        except Exception as e:
            info["mempolicy"] = "unavailable (%s)" % type(e).__name__
    except Exception as e:
        info["error"] = "%s: %s" % (type(e).__name__, e)
    return info


def copy_only(env, chain, x_host, y_host, slice_frames, steps):
    """Bare-copy ceiling of process_host: the same pinned buffers, staging buffers, time slices and streams, copies only
    (the strided H2D of every input slice and D2H of every output slice, both directions in flight at once)."""
    t = env.torch
    from distantspeech_b200 import _lib as L
    S, Mm, N = x_host.shape
    slices = chain._slices(N // chain.hop, slice_frames)
    n_max = max(b - a for a, b in slices) * chain.hop
    hp = chain._host_pipeline(S, Mm, n_max, x_host.dtype, y_host.dtype)
    xe, ye = x_host.element_size(), y_host.element_size()
    lib = L.lib()

    def once():
        for c, (f0, f1) in enumerate(slices):
            b = c & 1
            n0, n = f0 * chain.hop, (f1 - f0) * chain.hop
            L.check(lib.ds_memcpy2d_async(hp["xbuf"][b].data_ptr(), n * xe, x_host.data_ptr() + n0 * xe, N * xe, n * xe, S * Mm, 0,
                                          hp["s_in"].cuda_stream))
            L.check(lib.ds_memcpy2d_async(y_host.data_ptr() + n0 * ye, N * ye, hp["ybuf"][b].data_ptr(), n * ye, n * ye, S, 1,
                                          hp["s_out"].cuda_stream))
        hp["s_in"].synchronize()
        hp["s_out"].synchronize()
    once()
    env.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        once()
    env.barrier()
    return env.max_over_ranks((time.perf_counter() - t0) * 1e3)


# --------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation on the host cores
# --------------------------------------------------------------------------
CPU_SECONDS = {1: 1.0, 2: 10.0, 3: 1.0, 4: 2.0, 5: 0.25}      # audio seconds per stream and step: about a second of CPU work


def run_reference(args, env):
    if env.rank != 0:
        return
    from oracle import cpu_baselines as B           # the one other place bench.py may execute oracle/
    cfg = args.config
    secs = args.cpu_seconds or CPU_SECONDS[cfg]
    r = B.time_cpu(cfg, secs, steps=max(args.steps, 1), warmup=max(args.warmup, 0))
    line = {
        "impl": "reference", "metric": METRICS[cfg], "value": r["value"], "unit": "audio-s/s", "n_gpus": args.gpus,
        "steps": r["steps"], "warmup": r["warmup"], "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOADS[cfg],
                   "step": "one bounded sample of the workload: %s; every step and warm-up step really runs" % r["sample"],
                   "audio_s_per_step": r["audio_s_per_step"]},
        "cpu_baseline": {"value": r["value"], "unit": "audio-s/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(cfg, args):
    from oracle import cpu_baselines as B
    secs = args.cpu_seconds or {1: 2.0, 2: 10.0, 3: 2.0, 4: 10.0, 5: 0.25}[cfg]
    r = B.time_cpu(cfg, secs, steps=2 if cfg == 4 else 1, warmup=0)
    return {"value": r["value"], "unit": "audio-s/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}


def _parity_job(job):
    cfg, x, n = job
    from oracle import cpu_baselines as B
    return np.asarray(B.parity_reference(cfg, x[:, :n] if n else x), dtype=np.float64)


def parity_streams(cfg, xs, ys, what):
    """xs: list of [M, N] float32 arrays, ys: list of [N'] outputs; the oracle runs one stream per host core."""
    import multiprocessing as mp
    jobs = [(cfg, np.asarray(x, dtype=np.float64), 0) for x in xs]
    cores = min(len(jobs), len(os.sched_getaffinity(0)))
    with mp.get_context("fork").Pool(cores) as pool:
        refs = pool.map(_parity_job, jobs)
    worst_err, worst_snr = 0.0, 1e9
    for ref, out in zip(refs, ys):
        out = np.asarray(out, dtype=np.float64)[:ref.shape[0]]
        ref = ref[:out.shape[0]]
        worst_err = max(worst_err, float(np.max(np.abs(ref - out))))
        worst_snr = min(worst_snr, float(10 * np.log10(np.sum(ref ** 2) / max(np.sum((ref - out) ** 2), 1e-300))))
    return {"max_abs": worst_err, "snr_db": worst_snr, "streams_checked": len(xs), "seconds": what,
            "tolerance": "max-abs <= 1e-4 and SNR >= 60 dB (BASELINE.json north_star)",
            "ok": bool(worst_err <= 1e-4 and worst_snr >= 60)}


# translation unit of the dominant kernel (mcspp_fast.cu and every header it includes)
DOMINANT_KERNEL_SOURCES = ("mcspp_fast.cu", "chain_step.cuh", "mcspp_args.cuh", "perbin.cuh", "common.cuh")


def kernel_sources_sha():
    """sha1 over the sources of the dominant kernel's translation unit: the ncu-derived figures in profiles/traffic.json
    are only quoted for the kernel build they were captured from."""
    import hashlib
    h = hashlib.sha1()
    d = os.path.join(ROOT, "distantspeech_b200", "csrc")
    for f in DOMINANT_KERNEL_SOURCES:
        h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:12]


# --------------------------------------------------------------------------
# headline: config 4
# --------------------------------------------------------------------------
def run_headline(args, env):
    t, rank, world = env.torch, env.rank, env.world
    from distantspeech_b200 import _lib
    from distantspeech_b200.beamformer.MicArray import MicArray
    from distantspeech_b200.pipelines import MvdrMcsppChain
    from distantspeech_b200.sharding import shard_bounds, gather_validation_streams
    _lib.ensure_init()
    spg = args.streams_per_gpu or 1024
    lo, hi = shard_bounds(spg * world, rank, world)      # weak scaling: fixed streams per GPU
    S = hi - lo
    N = int(args.seconds * FS) // HOP * HOP
    mic = MicArray(arrayType="circular", r=0.05, M=M, n_fft=N_FFT)
    chain = MvdrMcsppChain(mic, look_angle=LOOK, n_fft=N_FFT, hop=HOP, full_state=bool(args.full_state),
                           fft_precision=args.fft)
    x = synth_device(t, S, mic, N, seed=0x5EED + lo)
    y = t.empty((S, N), dtype=t.float32, device="cuda")
    t.cuda.synchronize()

    def step():
        chain.reset_counters()
        chain._state.zero_() if chain._state is not None else None
        chain.process_device(x, out=y)

    ms_max, clocks = time_steps(env, step, args.steps, args.warmup)
    audio_step = world * S * (N / FS)
    value = audio_step * args.steps / (ms_max / 1e3)

    # ---- per-kernel timing for the roofline (CUDA events inside the library, same stream) ----
    phase = np.zeros(3)
    reps = 3
    for _ in range(reps):
        chain.reset_counters()
        chain._state.zero_()
        _, pm = chain.process_device_profiled(x, out=y)
        phase += np.array(pm)
    phase /= reps
    peaks = measured_peaks()
    hbm_peak = peaks["hbm_gbs"] if peaks else 6650.0
    algo_bytes = ALGO_BYTES_PER_AUDIO_S * S * (N / FS)
    dom = int(np.argmax(phase))
    names = ["stft_sq_kernel", "mcspp_fast_kernel" if not args.full_state else "mcspp_kernel", "istft_sq_kernel"]
    traffic, traffic_note = None, "no ncu capture of this build (profiles/traffic.json is keyed by the CUDA sources' sha)"
    flop_bf, pipe_pct = 1928.0, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        flop_bf = float(tj.get("mcspp_fast_kernel", {}).get("fp64_flop_per_bin_frame", flop_bf))
        if tj.get("csrc_sha") == kernel_sources_sha():
            per_stream = tj.get(names[dom], {}).get("dram_bytes_per_stream_10s")
            pipe_pct = tj.get("mcspp_fast_kernel", {}).get("ncu_pipe_fp64_pct")
            if per_stream is not None:
                traffic = per_stream * S * (N / FS) / 10.0      # ncu dram read+write of one launch, scaled to this launch
                traffic_note = "ncu --set full capture of this build (%s), dram__bytes_read.sum + dram__bytes_write.sum" % tj.get("capture", "profiles/")
    except Exception:
        pass
    achieved = algo_bytes / (phase[dom] / 1e3) / 1e9
    # what actually bounds the dominant kernel: the fp64 pipe.  FLOP per (bin, frame) from the executed SASS mix of the
    # frame loop (profiles/traffic.json); the pipe's peak is MEASURED here with the library's DFMA microbenchmark
    bin_frames = S * (N // HOP) * (N_FFT // 2 + 1 - 2)
    fp64_flop = flop_bf * bin_frames
    sm_mhz = (peaks or {}).get("sm_max_mhz", 1965.0)
    fp64_nominal = 148 * 64 * 2 * sm_mhz * 1e6 / 1e12
    fp64_measured = measure_fp64_peak(t)
    fp64_peak = fp64_measured or fp64_nominal
    fp64_ach = fp64_flop / (phase[1] / 1e3) / 1e12
    roofline = {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": traffic, "traffic_note": traffic_note,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                "algorithmic_bytes_per_launch": algo_bytes,
                "kernel_ms": {n: float(v) for n, v in zip(names, phase)},
                "whole_step": {"achieved": algo_bytes / (ms_max / args.steps / 1e3) / 1e9,
                               "frac": algo_bytes / (ms_max / args.steps / 1e3) / 1e9 / hbm_peak},
                "fp64_pipe": {"achieved_tflops": fp64_ach, "peak_tflops": fp64_peak,
                              "peak_source": "measured: ds_fp64_peak_run DFMA microbenchmark, best of 3, CUDA events" if fp64_measured
                              else "nominal 148 SM x 64 FMA/clk x 2 x SM clock (microbenchmark failed)",
                              "nominal_peak_tflops": fp64_nominal, "frac": fp64_ach / fp64_peak,
                              "flop_per_bin_frame": flop_bf, "ncu_pipe_fp64_pct": pipe_pct},
                "note": "the per-bin kernel is bound by the fp64 pipe, not by HBM: the contractual hbm fraction is small by "
                        "construction (SURVEY.md 8d); fp64_pipe is the binding roof -- see DESIGN.md"}

    # ---- parity: 8 streams spread over the whole job, full length, against the oracle ----
    n_sel = max(1, 8 // world)
    sel = [int(round(i)) for i in np.linspace(0, S - 1, n_sel)]
    ys = gather_validation_streams(y[sel].contiguous(), dst=0)     # NCCL gather: validation only, outside the timed region
    xs = gather_validation_streams(x[sel].contiguous(), dst=0)
    parity = None
    if rank == 0:
        xl = [xx.cpu().numpy() for blk in xs for xx in blk]
        yl = [yy.cpu().numpy() for blk in ys for yy in blk]
        parity = parity_streams(4, xl, yl, N / FS)
        parity["streams"] = "local indices %s of every rank's shard" % sel

    # ---- end to end through the public API with HOST buffers (pinned), copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        place = bind_near_gpu(t, env.local_rank)
        cs = args.slice_frames

        def e2e_run(xh, yh):
            chain.process_host(xh, yh, slice_frames=cs)           # warm-up (allocates the staging pipeline once)
            env.barrier()
            e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            for _ in range(args.steps):
                chain.process_host(xh, yh, slice_frames=cs)
            e1.record()
            env.barrier()
            wall = (time.perf_counter() - t0) * 1e3
            return env.max_over_ranks(e0.elapsed_time(e1)), env.max_over_ranks(wall)

        # int16 PCM in and out (the reference's on-disk format: load_audio / save_audio, beamformer/utils.py:182-196)
        x_pcm = t.empty((S, M, N), dtype=t.int16, pin_memory=True)
        x_pcm.copy_((x * 32767.0).round_().clamp_(-32768, 32767))   # x is scratch from here on
        y_pcm = t.empty((S, N), dtype=t.int16, pin_memory=True)
        chain.process_host(x_pcm, y_pcm, slice_frames=cs)         # allocates the staging pipeline copy_only reuses
        ms_c0 = copy_only(env, chain, x_pcm, y_pcm, cs, args.steps)
        ms_p, wall_p = e2e_run(x_pcm, y_pcm)
        ms_c = min(ms_c0, copy_only(env, chain, x_pcm, y_pcm, cs, args.steps))    # a ceiling: the faster of the passes before / after
        h2d, d2h = int(S * M * N * 2), int(S * N * 2)
        v_pcm, v_copy = audio_step * args.steps / (ms_p / 1e3), audio_step * args.steps / (ms_c / 1e3)
        e2e = {"value": v_pcm, "unit": "audio-s/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "api": "MvdrMcsppChain.process_host(int16 PCM x[S,M,N], int16 y[S,N]): pinned host buffers, %d-frame time slices of the "
                      "whole batch (strided DMA), copy/compute overlap on three streams, recursive state carried on the device; "
                      "load_audio / save_audio scalings fused into the kernels" % cs,
               "ms_per_step": ms_p / args.steps, "wall_ms_per_step": wall_p / args.steps,
               "copy_ceiling": {"value": v_copy, "unit": "audio-s/s", "ms_per_step": ms_c / args.steps,
                                "h2d_GBps_per_gpu": h2d / (ms_c / args.steps / 1e3) / 1e9,
                                "what": "same pinned buffers, staging buffers, time slices and streams, cudaMemcpy2DAsync only "
                                        "(H2D and D2H in flight together), wall clock, max over ranks; the faster of one pass before and "
                                        "one after the timed call (the host's copy rate varies by ~10 % from pass to pass)"},
               "frac_of_copy_ceiling": v_pcm / v_copy,
               "limiter": ("host->device copies: the call runs at the pace of the bare copies of the same buffers (kernels hidden); "
                           "at N > 1 the per-GPU copy rate (copy_ceiling.h2d_GBps_per_gpu) drops below one link's ~54 GB/s because all "
                           "ranks share the host's memory and PCIe root") if v_pcm / v_copy > 0.9
               else "see frac_of_copy_ceiling: below the bare-copy ceiling",
               "host_placement": place}
        del x_pcm, y_pcm
        # float32 host buffers beside it (twice the bytes over PCIe)
        x = synth_device(t, S, mic, N, seed=0x5EED + lo)
        x_host = t.empty((S, M, N), dtype=t.float32, pin_memory=True)
        x_host.copy_(x)
        y_host = t.empty((S, N), dtype=t.float32, pin_memory=True)
        del x
        t.cuda.empty_cache()
        chain._hp = None
        ms_f, _ = e2e_run(x_host, y_host)
        e2e["f32"] = {"value": audio_step * args.steps / (ms_f / 1e3), "unit": "audio-s/s",
                      "h2d_bytes_per_step": int(S * M * N * 4), "d2h_bytes_per_step": int(S * N * 4),
                      "api": "MvdrMcsppChain.process_host with float32 host buffers"}
        del x_host, y_host
        chain._hp = None
        t.cuda.empty_cache()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(4, args)

    line = None
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "audio-s/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOADS[4],
                       "streams_per_gpu": S, "seconds_per_stream": N / FS, "streams_total": S * world,
                       "fft": args.fft, "state": "full" if args.full_state else "output-only",
                       "l2": "inputs (%.1f GB per GPU) exceed L2; no flush needed" % (S * M * N * 4 / 1e9),
                       "parallelism": "streams sharded over %d GPU(s), no hot-path collective" % world},
            "clocks": clocks, "e2e": e2e, "gpu_launches": 5 * args.steps,
            "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
        }
    return line


# --------------------------------------------------------------------------
# configs 1, 2, 3, 5 -- same measurement rules, smaller step counts
# --------------------------------------------------------------------------
def run_config(cfg, args, env, steps, warmup, with_cpu=True, with_e2e=True):
    """One line of the same shape as the headline's for configuration cfg of BASELINE.json."""
    t, rank, world = env.torch, env.rank, env.world
    from distantspeech_b200 import _lib as L
    from distantspeech_b200.beamformer.MicArray import MicArray
    from distantspeech_b200.sharding import gather_validation_streams
    from oracle import cpu_baselines as B            # configuration table + checker (parity / cpu_baseline legs only)
    L.ensure_init()
    c = B.CONFIGS[cfg]
    fs, hop, Mm = c["fs"], c["hop"], c["M"]
    N = int(args.seconds * fs) // hop * hop
    mic = MicArray(arrayType=c["array"], r=c["r"], M=Mm, n_fft=c["n_fft"])
    if fs != 16000:                                  # the reference hard-wires 16 kHz (MicArray.py:27)
        mic.fs = fs
        mic.omega = 2 * np.pi * mic.freq_bin * mic.fs / mic.n_fft
    S = args.streams_per_gpu or {1: 2048, 2: 1024, 3: 4096, 5: 1}[cfg]
    peaks = measured_peaks()
    hbm_peak = peaks["hbm_gbs"] if peaks else 6650.0
    x = synth_device(t, S, mic, N, seed=0x5EED + rank * S, look=c["look"], interf=c["interf"], fs=fs)     # [S, M, N]
    extra = {}
    launches = None

    if cfg == 1:
        from distantspeech_b200.beamformer.adaptivebeamformer import adaptivebeamfomer
        ang = np.array(c["look"]) / 180 * np.pi
        ab = adaptivebeamfomer(mic, c["n_fft"], hop, c["n_fft"])
        out = {}

        def step():
            ab._state = None
            out["y"] = ab.process(x, ang, method=2)["data"]
        algo_per_s = Mm * fs * 4 + fs * 4
        launches, api = 6, "adaptivebeamfomer.process(x[S,M,N] CUDA tensor, angle, method=2)['data']"
        get_y = lambda: out["y"]                                                        # noqa: E731
        one = adaptivebeamfomer(mic, c["n_fft"], hop, c["n_fft"])
        x1 = x[0].contiguous()

        def step1():
            one._state = None
            one.process(x1, ang, method=2)
    elif cfg == 2:
        from distantspeech_b200.beamformer.fixedbeamformer import FixedBeamformer
        fb = FixedBeamformer(mic, c["n_fft"], hop, c["n_fft"])
        W = fb.compute_weights(list(c["look"]), "SD")[None]
        Wd = L.to_device(np.asarray(W, dtype=np.complex64), t.complex64)
        p = L.FixedBfParams(c["n_fft"], hop, S, Mm, N, 1, 0, 0, float(hop / fb.transform.W0))
        state = t.zeros(L.lib().ds_fixedbf_state_bytes(ctypes.byref(p)), dtype=t.uint8, device="cuda")
        y2 = t.empty((S, 1, N), dtype=t.float32, device="cuda")
        win = L.device_window(fb.transform.window, c["n_fft"])

        def step():
            state.zero_()
            L.check(L.lib().ds_fixedbf_run(ctypes.byref(p), L.ptr(win), L.ptr(Wd), L.ptr(state), L.ptr(x), L.ptr(y2),
                                           L.stream_ptr()), "ds_fixedbf_run")
        algo_per_s = Mm * fs * 4 + fs * 4
        launches, api = 2, "ds_fixedbf_run (the call FixedBeamformer.process makes) on x[S,M,N] CUDA"
        get_y = lambda: y2[:, 0, :]                                                     # noqa: E731
    elif cfg == 3:
        from distantspeech_b200.beamformer.FDGSC import FDGSC
        fd = FDGSC(mic, frameLen=256, angle=list(c["look"]))
        out = {}
        xw = [t.empty_like(x), t.empty_like(x)]
        y3 = t.empty((S, N), dtype=t.float32, device="cuda")
        side = t.cuda.Stream()
        ev_ready = [t.cuda.Event(), t.cuda.Event()]
        ev_used = [t.cuda.Event(), t.cuda.Event()]
        cnt = {"i": 0}
        xw[0].copy_(x)
        ev_ready[0].record()

        def step():
            # FDGSC.process overwrites its input with the DC-notched signal (FDGSC.py:213), so every step needs a fresh copy
            # of the batch: it is made on a side stream while the previous step computes (two input buffers), the way newly
            # arrived data would be; the copy's HBM traffic still falls inside the timed region
            i = cnt["i"]
            cur, nxt = i & 1, (i + 1) & 1
            main = t.cuda.current_stream()
            with t.cuda.stream(side):
                side.wait_event(ev_used[nxt])                 # the step that last read xw[nxt] is done with it
                xw[nxt].copy_(x, non_blocking=True)
                ev_ready[nxt].record(side)
            main.wait_event(ev_ready[cur])
            fd.reset_state()
            out["y"] = fd.process_device(xw[cur], out=y3)
            ev_used[cur].record(main)
            cnt["i"] = i + 1
        algo_per_s = Mm * fs * 4 + fs * 4
        launches, api = 2, "FDGSC.process_device(x[S,M,N] CUDA tensor) (the kernel calls FDGSC.process makes, output only; a fresh device copy of the batch per step is made on a side stream)"
        get_y = lambda: out["y"]                                                        # noqa: E731
    else:
        from distantspeech_b200.doa.srp import srp
        sp = srp(mic, engine="tensor")
        az, el = np.arange(360), np.arange(90)
        tau = np.stack([mic.compute_tau(np.array([a, e]) * np.pi / 180)[:, 0] for a in az for e in el])
        xin = x[0].t().contiguous()                                                     # [N, M]
        X = sp._spectrum(xin)                                                           # [T, M, K]
        out = {}

        def step():
            out["P"] = sp._steered_response(X, tau, True)
        D, T, K = tau.shape[0], X.shape[0], X.shape[2]
        launches, api = 3, "srp._steered_response (PHAT + re-tiling + tcgen05 contraction) over 32400 directions"

    ms_max, clocks = time_steps(env, step, steps, warmup)
    audio_step = world * S * (N / fs)
    value = audio_step * steps / (ms_max / 1e3)
    ms_step = ms_max / steps

    # ---- roofline ----
    if cfg == 5:
        flops = 8.0 * D * Mm * K * T
        bf16 = (peaks or {}).get("bf16_tflops", 1685.9)
        ach = flops / (ms_step / 1e3) / 1e12
        roofline = {"bound": "tensor", "kernel": "srp_tc_kernel (+ phat_kernel, srp_pack_kernel in the same step)", "achieved": ach,
                    "peak": bf16, "unit": "TFLOP/s", "frac": ach / bf16, "traffic": None,
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst); the kernel runs kind::tf32, whose dense rate is half of bf16",
                    "frac_of_tf32_rate": ach / (bf16 / 2), "algorithmic_flop_per_launch": flops,
                    "note": "8*D*M*K flop per frame (SURVEY.md 8d), D=%d directions, K=%d bins, T=%d frames" % (D, K, T)}
    else:
        algo = algo_per_s * S * (N / fs)
        ach = algo / (ms_step / 1e3) / 1e9
        roofline = {"bound": "hbm", "kernel": {1: "amvdr_kernel (+ stft / istft kernels in the same step)",
                                               2: "fixedbf_seq_kernel", 3: "fdgsc kernels"}[cfg],
                    "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": None,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                    "algorithmic_bytes_per_launch": algo,
                    "note": "whole step against SURVEY.md 8d's bytes (%d B per audio-second per stream)" % algo_per_s}

    # ---- parity against the oracle ----
    parity = None
    if cfg == 5:
        if rank == 0:
            from oracle import np_oracle as O
            geo = B.geometry(5)
            rng = np.random.default_rng(0)
            ti = np.sort(rng.choice(T, min(64, T), replace=False))
            di = np.sort(rng.choice(D, 256, replace=False))
            Y = O.Transform(channel=Mm, n_fft=c["n_fft"], hop_length=hop).stft(xin.cpu().numpy().astype(np.float64))
            tau_o = np.stack([O.method_tau(geo, np.array([i // 90, i % 90]) * np.pi / 180)[:, 0] for i in di])
            Pref = O.srp_map(Y[:, ti, :], geo.omega, tau_o)
            got = out["P"][t.as_tensor(di, device="cuda")][:, t.as_tensor(ti, device="cuda")].double().cpu().numpy()
            rel = float(np.max(np.abs(got - Pref) / np.abs(Pref)))
            Pfull = out["P"].sum(dim=1)
            ia = int(t.argmax(Pfull).item())
            parity = {"max_rel": rel, "frames_checked": int(len(ti)), "directions_checked": int(len(di)), "grid": "full 360 x 90",
                      "tolerance": "map rel-err <= 1e-3 (SURVEY.md 8d)", "argmax_az_el": [ia // 90, ia % 90],
                      "source_az_el": list(c["look"]),
                      "interferer_az_el": list(c["interf"]),
                      # the summed map peaks at the stronger of the two sources (the interferer is on all the time)
                      "ok": bool(rel <= 1e-3 and min(abs(ia // 90 - c["look"][0]), abs(ia // 90 - c["interf"][0])) <= 3)}
            if not parity["ok"]:
                raise RuntimeError("config 5 parity failed: %s" % parity)
    else:
        n_chk = {1: N, 2: N, 3: min(N, 256 * 125)}[cfg]           # FDGSC oracle: 2 s per stream (causal chain)
        n_sel = max(1, {1: 4, 2: 8, 3: 4}[cfg] // world)
        sel = [int(round(i)) for i in np.linspace(0, S - 1, n_sel)]
        yv = get_y()
        ys = gather_validation_streams(yv[sel][:, :n_chk].contiguous(), dst=0)
        xs = gather_validation_streams(x[sel][:, :, :n_chk].contiguous(), dst=0)
        if rank == 0:
            xl = [xx.cpu().numpy() for blk in xs for xx in blk]
            yl = [yy.cpu().numpy() for blk in ys for yy in blk]
            parity = parity_streams(cfg, xl, yl, n_chk / fs)
    if cfg == 1:
        e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
        step1(); step1()
        t.cuda.synchronize()
        e0.record()
        for _ in range(5):
            step1()
        e1.record()
        t.cuda.synchronize()
        extra["single_utterance"] = {"ms": e0.elapsed_time(e1) / 5, "x_realtime": (N / fs) / (e0.elapsed_time(e1) / 5 / 1e3),
                                     "what": "configs[0] as worded: ONE 10 s utterance through adaptivebeamfomer.process (incl. API hand-off)"}

    # ---- end to end: the same call with HOST buffers ----
    e2e = None
    if with_e2e and cfg in (1, 2, 3):
        reps = 2
        Se = min(S, 512)
        xh = t.empty((Se, Mm, N), dtype=t.float32, pin_memory=True)
        xh.copy_(x[:Se])
        yh = t.empty((Se, N), dtype=t.float32, pin_memory=True)

        def host_call():
            xd = xh.to("cuda", non_blocking=True)
            if cfg == 1:
                a1 = adaptivebeamfomer(mic, c["n_fft"], hop, c["n_fft"])
                yd = a1.process(xd, ang, method=2)["data"]
            elif cfg == 2:
                yd = fb._run(xd.permute(0, 2, 1), W)[:, 0, :]
            else:
                fd.reset()
                yd = fd.process_device(xd)
            yh.copy_(yd, non_blocking=True)
            t.cuda.synchronize()
        host_call()
        t0 = time.perf_counter()
        for _ in range(reps):
            host_call()
        dt = (time.perf_counter() - t0) / reps
        e2e = {"value": world * Se * (N / fs) / dt, "unit": "audio-s/s", "h2d_bytes_per_step": int(Se * Mm * N * 4),
               "d2h_bytes_per_step": int(Se * N * 4), "api": "pinned float32 host buffers -> %s -> pinned host buffer, %d streams per call, wall clock" % (api, Se)}
        del xh, yh
    elif with_e2e and cfg == 5:
        xh = xin.cpu().pin_memory()

        def host_call5():
            Ph = sp.compute_grid_spectrum(xh, az, el, as_torch=True)
            am = int(t.argmax(Ph.sum(dim=2)).item())              # the DOA estimate is what comes back to the host
            return am
        host_call5()
        t0 = time.perf_counter()
        for _ in range(2):
            host_call5()
        dt = (time.perf_counter() - t0) / 2
        e2e = {"value": (N / fs) / dt, "unit": "audio-s/s", "h2d_bytes_per_step": int(N * Mm * 4), "d2h_bytes_per_step": 8,
               "api": "srp.compute_grid_spectrum(host x[N,16], 360 az x 90 el) + argmax back to the host (includes the host-side "
                      "compute_tau loop over 32400 directions), wall clock"}

    cpu = None
    if rank == 0 and world == 1 and with_cpu and not args.no_cpu:
        cpu = cpu_baseline(cfg, args)
    line = None
    if rank == 0:
        line = {"metric": METRICS[cfg], "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": steps, "warmup": max(warmup, 3),
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "tf32" if cfg == 5 else "f64", "data": "synthetic",
                "config": {"workload": WORKLOADS[cfg], "streams_per_gpu": S, "seconds_per_stream": N / fs, "api": api,
                           "l2": "inputs exceed L2" if S * Mm * N * 4 > 126e6 else "single utterance: spectrum tiles are L2-resident by design"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": (launches or 0) * steps, "roofline": roofline, "cpu_baseline": cpu,
                "parity": parity}
        line.update(extra)
    del x
    t.cuda.empty_cache()
    return line


def main():
    args = parse()
    env = Env()
    if args.impl == "reference":
        run_reference(args, env)
        return
    env.init_gpu()
    if args.config != 4:
        line = run_config(args.config, args, env, args.steps, args.warmup)
    else:
        line = run_headline(args, env)
        if env.world == 1 and not args.no_configs:
            subs = {}
            for cfg in (1, 2, 3, 5):
                try:
                    sub = run_config(cfg, args, env, steps=5, warmup=3, with_cpu=True, with_e2e=False)
                    subs[str(cfg)] = {k: sub[k] for k in ("metric", "value", "unit", "steps", "warmup", "ms_per_step", "dtype", "config",
                                                          "clocks", "roofline", "cpu_baseline", "parity", "gpu_launches") if k in sub}
                    if "single_utterance" in sub:
                        subs[str(cfg)]["single_utterance"] = sub["single_utterance"]
                except Exception as e:                      # a failing side configuration must not take the headline down
                    subs[str(cfg)] = {"error": "%s: %s" % (type(e).__name__, e)}
                    env.torch.cuda.empty_cache()
            line["configs"] = subs
    if env.rank == 0 and line is not None:
        print(json.dumps(line), flush=True)
    if env.world > 1:
        env.dist.destroy_process_group()


if __name__ == "__main__":
    main()
