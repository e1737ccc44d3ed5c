#!/usr/bin/env python
"""Summarise a .ncu-rep (captured with `ncu --set full --clock-control none`) into the handful of counters
DESIGN.md / bench.py quote:  python tools/ncu_summary.py gpurun_out/x.ncu-rep [kernel-substring] > profiles/x_summary.txt"""
import csv, io, subprocess, sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors.sum",
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        rec = dict(zip(hdr, r))
        name = rec.get("Kernel Name", "")
        if want and want not in name:
            continue
        print("== %s" % name)
        for k in KEEP:
            if k in rec:
                print(k, rec[k], units[hdr.index(k)])
        print("--- stalls per issue")
        for k in sorted(hdr):
            if k.startswith(STALL) and k.endswith("_per_warp_active.pct") is False and k.endswith(".ratio"):
                v = float(rec[k] or 0)
                if v >= 0.1:
                    print("  ", k[len(STALL):].replace("_per_warp_active.ratio", ""), round(v, 3))


if __name__ == "__main__":
    main()
