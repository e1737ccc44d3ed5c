#!/bin/bash
O=gpurun_out/${1:-r2z}
mkdir -p $O
for v in "" $VARIANTS; do
  if [ -z "$v" ]; then python tools/time_chain_split.py 2>&1 | tail -1; else DS_B200_LIB=build/variants/$v.so python tools/time_chain_split.py 2>&1 | tail -1; fi
done | tee $O/time_split.txt
