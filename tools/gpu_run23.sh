#!/bin/bash
O=gpurun_out/${1:-r3w}
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -3 $O/pytest.log
S=512 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mcspp_fast|istft_sq|stft_sq" -s 8 -c 4 -o $O/chain -f python tools/time_chain.py > $O/ncu_chain.log 2>&1; tail -1 $O/ncu_chain.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_cfg4.csv python tools/time_chain.py > $O/ncu_l.log 2>&1
python tools/launch_summary.py $O/launches_cfg4.csv | head -8 | tee $O/launches_cfg4_summary.txt
S=512 IMPLS=pipeline timeout 900 ncu --set full --clock-control none -k regex:"fd_aic|fd_bm|fd_fir" -s 3 -c 3 -o $O/fdgsc -f python tools/time_fdgsc.py > $O/ncu_fd.log 2>&1; tail -1 $O/ncu_fd.log
IMPLS=pipeline timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fd_|dcnotch" -c 27 --csv --log-file $O/launches_fdgsc.csv python tools/time_fdgsc.py > $O/ncu_fdl.log 2>&1
python tools/launch_summary.py $O/launches_fdgsc.csv | head -12 | tee $O/launches_fdgsc_summary.txt
