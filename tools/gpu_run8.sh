#!/bin/bash
O=gpurun_out/${1:-r2i}
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "multibeam or fixed" > $O/pytest_mb.log 2>&1; tail -6 $O/pytest_mb.log
timeout 300 python tools/time_multibeam.py > $O/time_multibeam.txt 2>&1; cat $O/time_multibeam.txt
S=16 B=16 ENGINES=tensor,simt timeout 300 python tools/time_multibeam.py >> $O/time_multibeam.txt 2>&1; tail -2 $O/time_multibeam.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"mb_|stft|istft" -c 24 --csv --log-file $O/launches_mb.csv python tools/time_multibeam.py > $O/mb_ncu.log 2>&1
python tools/launch_summary.py $O/launches_mb.csv | tee $O/launches_mb_summary.txt
S=8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mb_tc" -c 1 -o $O/mb_tc python tools/time_multibeam.py > $O/mb_ncu_full.log 2>&1; echo "ncu full rc=$?"
