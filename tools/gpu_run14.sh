#!/bin/bash
O=gpurun_out/${1:-r2y}
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $O/pytest.log
for i in 1 2; do
python tools/time_chain_split.py 2>&1 | tail -1
DS_B200_LIB=build/variants/nosq.so python tools/time_chain_split.py 2>&1 | tail -1
done | tee $O/time_split.txt
