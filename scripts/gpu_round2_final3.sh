#!/bin/bash
# round 2: bench with the longer settle phase (3 runs)
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2final3
mkdir -p $O
for i in 1 2 3; do timeout 600 python bench.py --no-cpu-baseline > $O/bench_$i.json 2> $O/err.txt; done
for f in $O/bench_*.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -2 | tr '\n' ' ') $(grep -o '"value": [0-9.]*' $f | head -1) $(grep -o '"step_ms": {[^}]*}' $f) $(grep -o '"warmup_done": [0-9]*' $f)"; done
