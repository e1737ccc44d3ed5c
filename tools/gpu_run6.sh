#!/bin/bash
O=gpurun_out/${1:-r2g}
mkdir -p $O
IMPLS=pipeline S=1024 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fd_bm|fd_aic|fd_fir|fd_spec" -c 4 -o $O/fdgsc_pipeline python tools/time_fdgsc.py > $O/ncu_full.log 2>&1; echo "ncu rc=$?"
ls -la $O
