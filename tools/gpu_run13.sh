#!/bin/bash
O=gpurun_out/${1:-r2w}
mkdir -p $O
for i in 1 2; do
python tools/time_chain.py 2>&1 | tail -1
DS_B200_LIB=build/variants/x2.so python tools/time_chain.py 2>&1 | tail -1
done | tee $O/time_chain_x2.txt
IMPLS=pipeline python tools/time_fdgsc.py 2>&1 | tail -1 | tee $O/time_fdgsc_default.txt
IMPLS=pipeline DS_B200_LIB=build/variants/x2.so python tools/time_fdgsc.py 2>&1 | tail -1 | tee $O/time_fdgsc_x2.txt
DS_B200_LIB=build/variants/x2.so timeout 600 python -m pytest tests -m gpu -x -q -k "stft or istft or transform or chain_golden or fdgsc or fixed" 2>&1 | tail -3
