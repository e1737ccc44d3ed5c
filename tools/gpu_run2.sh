#!/bin/bash
# round-2 GPU run: tests, headline bench (+ configs), reference arm, FDGSC launch list
O=gpurun_out/${1:-r2b}
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -15 $O/pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fdgsc|dcnotch|fir" -c 40 --csv --log-file $O/launches_cfg3.csv python bench.py --config 3 --steps 1 --warmup 1 --no-cpu --no-e2e > $O/cfg3_ncu.log 2>&1; echo "ncu3 rc=$?"
python tools/launch_summary.py $O/launches_cfg3.csv > $O/launches_cfg3_summary.txt 2>&1
cat $O/launches_cfg3_summary.txt
head -c 6000 $O/bench.json; echo; tail -5 $O/bench.err; head -c 600 $O/bench_ref.json; tail -3 $O/bench_ref.err
