#!/usr/bin/env python
"""Phase-offset experiment on the per-bin chain kernel (build with -DDS_FAST_TRACE): do two warps that share a scheduler keep
their start-up phase offset, or do they fall into lock-step?  Prints, per pair, the offset of frame starts over 64 frames."""
import sys, os, ctypes as C, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from distantspeech_b200 import _lib as L
from distantspeech_b200.beamformer.MicArray import MicArray
from distantspeech_b200.pipelines import MvdrMcsppChain
S = 1024; N = 256 * 625
mic = MicArray(arrayType="circular", r=0.05, M=8, n_fft=512)
ch = MvdrMcsppChain(mic, look_angle=(30, 0))
g = torch.Generator(device="cuda"); g.manual_seed(1)
x = torch.randn((S, 8, N), device="cuda", generator=g) * 0.1
y = torch.empty((S, N), device="cuda")
ch.process_device(x, out=y); torch.cuda.synchronize()
buf = np.zeros(8192 * 66, dtype=np.int64)
rc = L.lib().ds_debug_trace_read(buf.ctypes.data_as(C.c_void_p), C.c_size_t(buf.nbytes)); assert rc == 0
tr = buf.reshape(8192, 66)
nw = 4080 * 2
tr = tr[:nw]
sm, wid = tr[:, 64], tr[:, 65]
ft = np.diff(tr[:, :64], axis=1)
print("frame time cycles: median %.0f  p10 %.0f p90 %.0f" % (np.median(ft), np.percentile(ft, 10), np.percentile(ft, 90)))
# pairs on the same SM and scheduler (warp slot % 4) whose traced windows overlap in time
shown = 0; stats = []
for s_ in np.unique(sm)[:40]:
    idx = np.where(sm == s_)[0]
    for sched in range(4):
        grp = idx[(wid[idx] % 4) == sched]
        for a in range(len(grp)):
            for b in range(a + 1, len(grp)):
                A, B = tr[grp[a], :64], tr[grp[b], :64]
                lo, hi = max(A[0], B[0]), min(A[-1], B[-1])
                if hi - lo < 30 * 7000: continue
                # offset of B's frame starts relative to the latest A start, as a fraction of the frame time
                offs = []
                for tb in B:
                    if tb < A[0] or tb > A[-1]: continue
                    i = np.searchsorted(A, tb, side="right") - 1
                    if i + 1 < len(A): offs.append((tb - A[i]) / (A[i + 1] - A[i]))
                if len(offs) > 20:
                    stats.append((offs[0], offs[-1], np.std(np.unwrap(np.array(offs) * 2 * np.pi)) / (2 * np.pi)))
                    if shown < 12:
                        shown += 1
                        print("sm %d sched %d warps %d,%d: offset first %.2f ... last %.2f  path %s" % (s_, sched, grp[a], grp[b], offs[0], offs[-1], " ".join("%.2f" % o for o in offs[::6])))
st = np.array(stats)
print("pairs", len(st), " mean |drift| over the window (frames): %.3f" % np.mean(np.abs(((st[:, 1] - st[:, 0] + 0.5) % 1) - 0.5)))
print("histogram of offsets at the end of the window:", np.histogram(st[:, 1] % 1, bins=10, range=(0, 1))[0])

# frame time of a warp as a function of the phase offset of the warp that shares its scheduler (only frames during which
# exactly that one partner was resident on the scheduler for the whole frame)
bins = [[] for _ in range(10)]
for s_ in np.unique(sm):
    idx = np.where(sm == s_)[0]
    for sched in range(4):
        grp = idx[(wid[idx] % 4) == sched]
        for a in grp:
            A = tr[a, :64]
            for b in grp:
                if a == b: continue
                B = tr[b, :64]
                for i in range(63):
                    t0, t1 = A[i], A[i + 1]
                    jdx = np.searchsorted(B, t0, side="right") - 1
                    if jdx < 0 or jdx + 2 >= 64: continue
                    off = (t0 - B[jdx]) / (B[jdx + 1] - B[jdx])      # how far the partner is into its frame when A starts one
                    bins[min(9, int(off * 10))].append(t1 - t0)
print("partner's phase at my frame start -> my frame time (cycles): " + "  ".join("%.1f:%.0f(n=%d)" % (i / 10, np.median(b) if b else 0, len(b)) for i, b in enumerate(bins)))
