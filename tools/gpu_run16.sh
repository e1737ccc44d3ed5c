#!/bin/bash
O=gpurun_out/${1:-r3a}
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 --no-configs --no-cpu > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python - $O/bench.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "frac", d["e2e"]["frac_of_copy_ceiling"], d["roofline"]["kernel_ms"], d["parity"])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_sq -s 2 -c 1 -o $O/stft_sq -f python tools/time_chain.py > $O/ncu_stft.log 2>&1; tail -2 $O/ncu_stft.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_cfg4.csv python tools/time_chain.py > $O/ncu_l.log 2>&1
python tools/launch_summary.py $O/launches_cfg4.csv | tee $O/launches_cfg4_summary.txt
