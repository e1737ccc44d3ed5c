#!/bin/bash
O=gpurun_out/${1:-r2f}
mkdir -p $O
python tools/clock_probe.py > $O/clock_probe.txt 2>&1
cat $O/clock_probe.txt
timeout 600 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max --clock-control none -k regex:"fd_|dcnotch" -c 30 --csv --log-file $O/launches_cfg3.csv python bench.py --config 3 --steps 1 --warmup 1 --no-cpu --no-e2e > $O/cfg3_ncu.log 2>&1; echo "ncu3 rc=$?"
python tools/launch_summary.py $O/launches_cfg3.csv > $O/launches_cfg3_summary.txt 2>&1
cat $O/launches_cfg3_summary.txt
grep "sm__cycles_elapsed" $O/launches_cfg3.csv | head -12 | cut -c1-260
