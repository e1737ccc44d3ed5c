#!/bin/bash
O=gpurun_out/${1:-r3x}
mkdir -p $O
timeout 1500 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
timeout 900 python tools/bench_configs.py > $O/bench_configs.jsonl 2> $O/bench_configs.err; echo "configs rc=$?"
python - $O/bench.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "frac", d["e2e"]["frac_of_copy_ceiling"], d["roofline"]["kernel_ms"], d["roofline"]["traffic"], d["parity"]["snr_db"], d["clocks"])
for k,c in d.get("configs", {}).items():
    print(k, c["ms_per_step"], c["value"], c["roofline"].get("frac"), c.get("parity", {}).get("ok"), c.get("cpu_baseline", {}).get("value"))
PY
tail -5 $O/bench_configs.jsonl | cut -c1-300
