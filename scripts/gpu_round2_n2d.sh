#!/bin/bash
# round 2: 2-GPU gradient check of the staged all-reduce with deferred level-0 wgrads, then bench config 2 at N=2 and N=1
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2n2d
mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 scripts/dev_ddp_check.py 2>&1 | grep "rank\|Error\|error" | head; echo "ddp check rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_c2_n2.json 2> $O/bench_c2_n2.err
timeout 400 python bench.py --no-cpu-baseline > $O/bench_c2_n1.json 2> $O/bench_c2_n1.err
timeout 600 python -m pytest tests/test_zz_gpu_step_program.py -q -m gpu --timeout=500 2>&1 | tail -2
for f in $O/bench_*.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -2 | tr '\n' ' ') $(grep -o '"value": [0-9.]*' $f | head -1) $(grep -o '"per_rank": {[^}]*}' $f)"; done
