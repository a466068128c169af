#!/bin/bash
# round 2: pre-reserved spare blocks in the large and small pools of the staging and training streams: 11 runs
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2bd
mkdir -p $O
for i in 1 2 3 4 5 6 7 8 9 10 11; do timeout 400 python bench.py --no-cpu-baseline > $O/bench_$i.json 2> $O/err.txt; done
for f in $O/bench_*.json; do echo "$(basename $f) $(grep -o '"ms_per_step": [0-9.]*' $f | head -2 | tr '\n' ' ') $(grep -o '"max": [0-9.]*, "argmax": [0-9]*' $f) $(grep -o '"cudaMalloc_segments_during_region": [0-9-]*' $f)"; done
