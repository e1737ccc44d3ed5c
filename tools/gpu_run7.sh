#!/bin/bash
O=gpurun_out/${1:-r2h}
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "fdgsc or FDGSC" > $O/pytest_fdgsc.log 2>&1; tail -4 $O/pytest_fdgsc.log
IMPLS=pipeline python tools/time_fdgsc.py > $O/time_fdgsc.txt 2>&1; cat $O/time_fdgsc.txt
