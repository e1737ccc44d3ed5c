#!/usr/bin/env python
"""tools/update_traffic.py REPORT.ncu-rep STREAMS [capture-name]: refresh profiles/traffic.json from an `ncu --set full` capture
of the config-4 step (tools/time_chain.py with S=STREAMS): DRAM bytes per stream of the three chain kernels, the fp64 pipe
utilisation of the per-bin kernel, and the sha of the dominant kernel's sources (bench.py quotes the figures only for that build)."""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

rep, S = sys.argv[1], int(sys.argv[2])
name = sys.argv[3] if len(sys.argv) > 3 else rep
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
path = os.path.join(ROOT, "profiles", "traffic.json")
tj = json.load(open(path))
for r in rows[2:]:
    rec = dict(zip(hdr, r))
    kn = rec["Kernel Name"]
    key = next((k for k in ("mcspp_fast_kernel", "stft_sq_kernel", "istft_sq_kernel", "stft_kernel", "istft_seq_kernel") if k in kn), None)
    if key is None:
        continue
    units = dict(zip(hdr, rows[1]))
    def gb(metric):
        v = float(rec[metric]); u = units[metric]
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
    d = tj.setdefault(key, {})
    d["dram_bytes_per_stream_10s"] = (gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")) / S
    d["capture"] = "%s (S=%d)" % (name, S)
    if key == "mcspp_fast_kernel":
        d["ncu_pipe_fp64_pct"] = float(rec["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"])
tj["csrc_sha"] = bench.kernel_sources_sha()
tj["capture"] = name
json.dump(tj, open(path, "w"), indent=1)
print(json.dumps(tj, indent=1))
