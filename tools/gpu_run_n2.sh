#!/bin/bash
O=gpurun_out/${1:-r2k}
mkdir -p $O
nvidia-smi topo -m > $O/topo_n2.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?"
head -c 4000 $O/bench_n2.json; echo; tail -5 $O/bench_n2.err | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err; echo "ref n2 rc=$?"
head -c 600 $O/bench_ref_n2.json; echo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --config 3 --steps 3 --warmup 3 --no-cpu > $O/bench_cfg3_n2.json 2> $O/bench_cfg3_n2.err; echo "cfg3 n2 rc=$?"
head -c 700 $O/bench_cfg3_n2.json; echo
