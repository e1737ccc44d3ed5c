#!/bin/bash
# FDGSC pipeline: tests, config-3 bench, launch list
O=gpurun_out/${1:-r2d}
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q -k "fdgsc or gsc or FDGSC" > $O/pytest_fdgsc.log 2>&1; echo "pytest rc=$?" >> $O/pytest_fdgsc.log
tail -25 $O/pytest_fdgsc.log
timeout 900 python bench.py --config 3 --steps 5 --warmup 3 --no-cpu > $O/bench_cfg3.json 2> $O/bench_cfg3.err; echo "bench rc=$?"
head -c 1500 $O/bench_cfg3.json; tail -3 $O/bench_cfg3.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fd_|fdgsc|dcnotch" -c 60 --csv --log-file $O/launches_cfg3.csv python bench.py --config 3 --steps 1 --warmup 1 --no-cpu --no-e2e > $O/cfg3_ncu.log 2>&1; echo "ncu3 rc=$?"
python tools/launch_summary.py $O/launches_cfg3.csv > $O/launches_cfg3_summary.txt 2>&1
cat $O/launches_cfg3_summary.txt
