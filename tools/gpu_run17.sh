#!/bin/bash
O=gpurun_out/${1:-r3b}
mkdir -p $O
run() { timeout 300 python bench.py --config 2 --steps 10 --warmup 3 --no-cpu --no-e2e 2>$O/bench_cfg2_$1.err | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg2 [$1] ms', d['ms_per_step'], 'value', d['value'], 'roofline', d['roofline'].get('frac'), 'parity', d['parity']['snr_db'])"; }
for v in $VARIANTS; do DS_B200_LIB=build/variants/$v.so run $v; done | tee $O/cfg2.txt
