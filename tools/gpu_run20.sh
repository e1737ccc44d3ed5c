#!/bin/bash
O=gpurun_out/${1:-r3r}
mkdir -p $O
python tools/time_multibeam.py 2>&1 | tail -1 | tee $O/time_multibeam.txt
S=16 B=16 ENGINES=tensor,simt python tools/time_multibeam.py 2>&1 | tail -2 | tee -a $O/time_multibeam.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches_multibeam.csv python tools/time_multibeam.py > $O/ncu_mb.log 2>&1
python tools/launch_summary.py $O/launches_multibeam.csv | head -12 | tee $O/launches_multibeam_summary.txt
