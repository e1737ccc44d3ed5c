#!/usr/bin/env python
"""Times ds_srp_run (tensor path) for D directions x T frames x 16 mics x 513 bins with CUDA events; CHECK=1 compares
a slice with the CUDA-core path.  D, T via the environment.  Also the ncu target for the SRP kernel."""
import sys, os, ctypes as C, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distantspeech_b200 import _lib as L
L.ensure_init()
D, T, M, K = int(os.environ.get("D", 32400)), int(os.environ.get("T", 937)), 16, 513
g = torch.Generator(device="cuda"); g.manual_seed(3)
Y = torch.randn((K, T, M, 2), device="cuda", generator=g)
Y = torch.view_as_complex(Y / Y.norm(dim=-1, keepdim=True)).contiguous()
tau = (torch.rand((D, M), device="cuda", generator=g) - 0.5) * 4e-4
P = torch.empty((D, T), device="cuda")
ws = torch.empty(max(L.lib().ds_srp_workspace_bytes(T, M, K, 1), 1), dtype=torch.uint8, device="cuda")
def run(tc):
    L.check(L.lib().ds_srp_run(D, T, M, K, 48000.0, 1024, L.ptr(tau), L.ptr(Y), L.ptr(ws), L.ptr(P), tc, L.stream_ptr()), "srp")
run(1); torch.cuda.synchronize()
ts = []
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(1); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
Ptc = P.clone()
chk = ""
if os.environ.get("CHECK"):
    Dc = 256
    Pr = torch.empty((Dc, T), device="cuda")
    L.check(L.lib().ds_srp_run(Dc, T, M, K, 48000.0, 1024, L.ptr(tau), L.ptr(Y), None, L.ptr(Pr), 0, L.stream_ptr()), "srp")
    rel = ((Ptc[:Dc] - Pr).abs().max() / Pr.abs().max()).item()
    chk = " rel-err vs simt %.2e" % rel
ms = min(ts)
print(os.environ.get("DS_B200_LIB", "default"), "ms %.2f  TFLOP/s(alg) %.1f%s" % (ms, 8.0 * D * M * K * T / ms / 1e9, chk))
