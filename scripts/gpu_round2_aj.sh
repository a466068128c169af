#!/bin/bash
# round 2, state "aj": per-layer table of config 5 (Res16UNet34D @1cm)
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2aj
mkdir -p $O
LGS_BENCH_LAYERS=1 timeout 600 python bench.py --no-cpu-baseline --config 5 --steps 3 --warmup 3 > $O/bench_c5.json 2> $O/bench_c5.err
grep LAYER $O/bench_c5.err | head -24
