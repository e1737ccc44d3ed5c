#!/bin/bash
O=gpurun_out/${1:-r3u}
mkdir -p $O
run() { timeout 300 python bench.py --config ${CFG:-1} --steps 10 --warmup 3 --no-cpu --no-e2e 2>$O/bench_$1.err | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg${CFG:-1} [$1] ms', d['ms_per_step'], 'value', d['value'], 'parity', d['parity'].get('snr_db'), d['parity']['ok'])"; }
for v in $VARIANTS; do DS_B200_LIB=build/variants/$v.so run $v; done | tee $O/cfg.txt
