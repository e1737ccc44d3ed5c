#!/bin/bash
O=gpurun_out/${1:-r2j}
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q -s -k "mixed_precision or chain_golden" > $O/pytest_mixed.log 2>&1; grep "dB\|passed\|failed" $O/pytest_mixed.log | tail -8
for p in f64 mixed f64 mixed; do PRECISION=$p python tools/time_chain.py 2>&1 | tail -1; done | tee $O/time_chain_mixed.txt
