#!/bin/bash
# round 2: stage() stalls vs cudaMalloc: 5 runs without and 7 runs with the pre-reserved staging block
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2bc
mkdir -p $O
for i in 1 2 3 4 5; do LGS_BENCH_NO_PRERESERVE=1 timeout 400 python bench.py --no-cpu-baseline > $O/bench_nores_$i.json 2> $O/err.txt; done
for i in 1 2 3 4 5 6 7; do timeout 400 python bench.py --no-cpu-baseline > $O/bench_res_$i.json 2> $O/err.txt; done
for f in $O/bench_*.json; do echo "$(basename $f) $(grep -o '"ms_per_step": [0-9.]*' $f | head -2 | tr '\n' ' ') $(grep -o '"max": [0-9.]*, "argmax": [0-9]*' $f) $(grep -o '"host_phases_ms_at_max[^]]*]' $f | grep -o '\[.*') $(grep -o '"cudaMalloc_segments_during_region": [0-9-]*' $f)"; done
