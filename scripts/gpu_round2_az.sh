#!/bin/bash
# round 2: where does the occasional long step of the first timed region come from?  6 runs, per-step host issue times
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2az
mkdir -p $O
for i in 1 2 3 4 5 6; do timeout 400 python bench.py --no-cpu-baseline > $O/bench_$i.json 2> $O/err.txt; done
for f in $O/bench_*.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -2 | tr '\n' ' ') $(grep -o '"step_ms": {[^}]*}' $f)"; done
