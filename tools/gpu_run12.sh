#!/bin/bash
O=gpurun_out/${1:-r2v}
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q -s -k "srp" > $O/pytest_srp.log 2>&1; grep "rel err\|passed\|failed\|max err" $O/pytest_srp.log | tail
timeout 300 python bench.py --config 5 --steps 10 --warmup 3 --no-cpu --no-e2e > $O/bench_cfg5.json 2> $O/bench_cfg5.err; python -c "
import json; d=json.load(open('$O/bench_cfg5.json')); print('cfg5 ms', d['ms_per_step'], 'TF', d['roofline']['achieved'], 'parity', d['parity'])"
