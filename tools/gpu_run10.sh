#!/bin/bash
O=gpurun_out/${1:-r2l}
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -5 $O/pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 --no-configs --no-cpu > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python - <<'PY'
import json,sys
d=json.load(open(sys.argv[1] if len(sys.argv)>1 else "gpurun_out/r2l/bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "ceiling", d["e2e"]["copy_ceiling"]["value"], "frac", d["e2e"]["frac_of_copy_ceiling"], "f32", d["e2e"]["f32"]["value"])
PY
for sf in 16 64; do timeout 300 python bench.py --steps 10 --warmup 3 --no-configs --no-cpu --slice-frames $sf 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('slice', $sf, 'e2e', d['e2e']['value'], 'ceiling', d['e2e']['copy_ceiling']['value'])"; done
tail -3 $O/bench.err
