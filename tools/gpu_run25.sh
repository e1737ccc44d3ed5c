#!/bin/bash
O=gpurun_out/${1:-r4h}
mkdir -p $O
S=512 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mcspp_fast|istft_sq|stft_sq" -s 8 -c 4 -o $O/chain -f python tools/time_chain.py > $O/ncu_chain.log 2>&1; tail -1 $O/ncu_chain.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_cfg4.csv python tools/time_chain.py > $O/ncu_l.log 2>&1
python tools/launch_summary.py $O/launches_cfg4.csv 2>/dev/null | head -6 | tee $O/launches_cfg4_summary.txt
