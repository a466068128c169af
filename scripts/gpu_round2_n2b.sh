#!/bin/bash
# round 2: 2-GPU run of bench config 2 with the staged all-reduce buckets, and 1 GPU on the same box
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2n2b
mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_c2_n2.json 2> $O/bench_c2_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_c2_n2b.json 2> $O/bench_c2_n2b.err
timeout 400 python bench.py --no-cpu-baseline > $O/bench_c2_n1.json 2> $O/bench_c2_n1.err
for f in $O/bench_*.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -2 | tr '\n' ' ') $(grep -o '"value": [0-9.]*' $f | head -1) $(grep -o '"step_ms": {[^}]*}' $f) $(grep -o '"n_gpus": [0-9]*' $f) $(grep -o '"loss": [0-9.]*' $f)"; done
tail -n 3 $O/bench_c2_n2.err
