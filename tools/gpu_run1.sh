#!/bin/bash
# round-2 GPU run 1: tests, headline bench (+ configs), reference arm, divergence report, launch lists, host topology
mkdir -p gpurun_out/r2a
O=gpurun_out/r2a
nvidia-smi topo -m > $O/topo.txt 2>&1
(lscpu | head -30; numactl -H 2>&1 | head -20; cat /sys/fs/cgroup/cpuset.cpus.effective /sys/fs/cgroup/cpuset.mems.effective 2>&1; nproc) > $O/host.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
timeout 600 python tools/divergence_report.py > $O/divergence_r02.txt 2> $O/divergence.err; echo "div rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_cfg3.csv python bench.py --config 3 --steps 1 --warmup 1 --no-cpu --no-e2e > $O/cfg3_ncu.log 2>&1; echo "ncu3 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"stft|mcspp|pcm" -c 200 --csv --log-file $O/launches_cfg4.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-configs > $O/cfg4_ncu.log 2>&1; echo "ncu4 rc=$?"
python tools/launch_summary.py $O/launches_cfg3.csv > $O/launches_cfg3_summary.txt 2>&1
python tools/launch_summary.py $O/launches_cfg4.csv > $O/launches_cfg4_summary.txt 2>&1
head -c 3000 $O/bench.json; echo; tail -3 $O/bench.err
