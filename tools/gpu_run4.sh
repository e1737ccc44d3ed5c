#!/bin/bash
O=gpurun_out/${1:-r2e}
mkdir -p $O
python tools/time_fdgsc.py > $O/time_default.txt 2>&1
for v in fd33 fd43 fd22; do DS_B200_LIB=build/variants/$v.so python tools/time_fdgsc.py 2>&1 | grep pipeline > $O/time_$v.txt; done
tail -n 3 $O/time_*.txt
