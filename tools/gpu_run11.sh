#!/bin/bash
O=gpurun_out/${1:-r2s}
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "fdgsc or FDGSC" > $O/pytest_fdgsc.log 2>&1; tail -3 $O/pytest_fdgsc.log
IMPLS=pipeline python tools/time_fdgsc.py > $O/time_fdgsc.txt 2>&1; cat $O/time_fdgsc.txt
IMPLS=pipeline timeout 600 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max --clock-control none -k regex:"fd_|dcnotch" -c 18 --csv --log-file $O/launches_fdgsc.csv python tools/time_fdgsc.py > $O/ncu.log 2>&1
python tools/launch_summary.py $O/launches_fdgsc.csv | tee $O/launches_fdgsc_summary.txt
