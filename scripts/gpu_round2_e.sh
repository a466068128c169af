#!/bin/bash
# round 2, state "e": 2-GPU run of the native driver (bucketed all-reduce launched between the halves of backward)
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2g
mkdir -p $O
for drv in native facade; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --driver $drv > $O/bench_n2_$drv.json 2> $O/bench_n2_$drv.err
done
timeout 300 python bench.py --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err
for f in $O/bench_*.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -2 | tr '\n' ' ') $(grep -o '"value": [0-9.]*' $f | head -2 | tr '\n' ' ') $(grep -o '"loss": [0-9.]*' $f)"; done
tail -5 $O/bench_n2_native.err
