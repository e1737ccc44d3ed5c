#!/bin/bash
O=gpurun_out/${1:-r3f}
mkdir -p $O
python tools/time_srp.py 2>&1 | tail -1 | tee $O/time_srp.txt
for v in $VARIANTS; do DS_B200_LIB=build/variants/$v.so python tools/time_srp.py 2>&1 | tail -1; done | tee -a $O/time_srp.txt
if [ -n "$NCU" ]; then timeout 600 ncu --set full --clock-control none --import-source on -k regex:srp_tc_kernel -s 1 -c 1 -o $O/srp_tc -f python tools/time_srp.py > $O/ncu_srp.log 2>&1; tail -2 $O/ncu_srp.log; fi
