#!/usr/bin/env python
"""What the SM clock really does under a sustained FDGSC load: NVML clock / power / throttle reasons every 5 ms and
nvidia-smi every 200 ms while FDGSC.process_device runs back to back for ~4 s; the same for the config-4 chain."""
import os, subprocess, sys, threading, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pynvml
from distantspeech_b200.beamformer.MicArray import MicArray
from distantspeech_b200.beamformer.FDGSC import FDGSC
from distantspeech_b200.pipelines import MvdrMcsppChain
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
def sample(stop, out):
    while not stop.is_set():
        out.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_MEM),
                    pynvml.nvmlDeviceGetPowerUsage(h) / 1e3, pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
        time.sleep(0.005)
def smi(stop, out):
    while not stop.is_set():
        out.append(subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,clocks_throttle_reasons.active", "--format=csv,noheader"],
                                  capture_output=True, text=True).stdout.strip())
        time.sleep(0.2)
def probe(name, fn, secs=4.0):
    fn(); torch.cuda.synchronize()
    stop, a, b = threading.Event(), [], []
    t1, t2 = threading.Thread(target=sample, args=(stop, a)), threading.Thread(target=smi, args=(stop, b))
    t1.start(); t2.start()
    t0 = time.time(); n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < secs:
        fn(); n += 1; torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    stop.set(); t1.join(); t2.join()
    sm = sorted(x[0] for x in a); pw = sorted(x[2] for x in a)
    print(name, "calls", n, "ms/call %.2f" % (e0.elapsed_time(e1) / n), "| nvml sm MHz min/med/max", sm[0], sm[len(sm) // 2], sm[-1],
          "| power W med/max %.0f %.0f" % (pw[len(pw) // 2], pw[-1]), "| reasons", sorted(set(hex(x[3]) for x in a)), flush=True)
    print("   nvidia-smi:", b[:3], "...", b[-2:], flush=True)
S = 4096; N = 256 * 625
mic = MicArray(arrayType="linear", r=0.05, M=6, n_fft=256)
x = torch.randn((S, 6, N), device="cuda") * 0.1
y = torch.empty((S, N), device="cuda")
fd = FDGSC(mic, frameLen=256, angle=[90, 0])
def run_fd():
    fd.reset_state(); fd.process_device(x, out=y)
probe("fdgsc pipeline", run_fd)
del x, fd
torch.cuda.empty_cache()
mic8 = MicArray(arrayType="circular", r=0.05, M=8, n_fft=512)
ch = MvdrMcsppChain(mic8, look_angle=(30, 0))
x8 = torch.randn((1024, 8, N), device="cuda") * 0.1
y8 = torch.empty((1024, N), device="cuda")
def run_ch():
    ch.reset_counters()
    if ch._state is not None: ch._state.zero_()
    ch.process_device(x8, out=y8)
probe("chain (config 4)", run_ch)
