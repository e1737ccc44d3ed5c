#!/bin/bash
O=gpurun_out/${1:-r2x}
mkdir -p $O
nvidia-smi topo -m > $O/topo_n8.txt 2>&1
(nproc; lscpu | grep -i "numa\|socket\|model name"; free -g | head -2) > $O/host_n8.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 > $O/bench_n8.json 2> $O/bench_n8.err; echo "bench n8 rc=$?"
python - <<PY
import json
l=[x for x in open("$O/bench_n8.json").read().splitlines() if x.strip()]
print("stdout lines:", len(l))
d=json.loads(l[-1])
print("value", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "ceiling", d["e2e"]["copy_ceiling"]["value"], "frac", d["e2e"]["frac_of_copy_ceiling"], "f32", d["e2e"]["f32"]["value"], d["e2e"]["host_placement"], d["parity"])
PY
tail -3 $O/bench_n8.err | cut -c1-300
cat $O/host_n8.txt
