#!/bin/bash
# round 2, state "al": A/B of the fused addend on one box
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2al
mkdir -p $O
timeout 400 python bench.py --no-cpu-baseline > $O/bench_fused.json 2> $O/bench_fused.err
LGS_FUSE_ADDEND=0 timeout 400 python bench.py --no-cpu-baseline > $O/bench_separate.json 2> $O/bench_separate.err
timeout 400 python bench.py --no-cpu-baseline > $O/bench_fused2.json 2> $O/bench_fused2.err
LGS_FUSE_ADDEND=0 timeout 400 python bench.py --no-cpu-baseline > $O/bench_separate2.json 2> $O/bench_separate2.err
for f in $O/bench_*.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -2 | tr '\n' ' ') $(grep -o '"value": [0-9.]*' $f | head -1) $(grep -o '"step_ms": {[^}]*}' $f) $(grep -o '"gpu_launches": [0-9]*' $f)"; done
