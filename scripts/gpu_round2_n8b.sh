#!/bin/bash
# round 2, final state: 8- and 4-GPU runs of bench config 2 as the driver launches them
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2n8b
mkdir -p $O
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_c2_n8.json 2> $O/bench_c2_n8.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_c2_n4.json 2> $O/bench_c2_n4.err
for f in $O/bench_*.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -2 | tr '\n' ' ') $(grep -o '"value": [0-9.]*' $f | head -1) $(grep -o '"step_ms": {[^}]*}' $f)"; done
