#!/bin/bash
O=gpurun_out/${1:-r3o}
mkdir -p $O
S=512 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mcspp_fast|istft_seq|stft_sq" -s 8 -c 4 -o $O/chain -f python tools/time_chain.py > $O/ncu_chain.log 2>&1; tail -2 $O/ncu_chain.log
