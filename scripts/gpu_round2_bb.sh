#!/bin/bash
# round 2: which host phase of a step stalls?  10 runs, host phase times at the longest step
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2bb
mkdir -p $O
for i in 1 2 3 4 5 6 7 8 9 10; do timeout 400 python bench.py --no-cpu-baseline > $O/bench_$i.json 2> $O/err.txt; done
for f in $O/bench_*.json; do echo "$(basename $f) $(grep -o '"ms_per_step": [0-9.]*' $f | head -2 | tr '\n' ' ') $(grep -o '"max": [0-9.]*, "argmax": [0-9]*' $f) $(grep -o '"host_phases_ms_at_max[^]]*]' $f) $(grep -o '"host_phases_ms_median[^]]*]' $f)"; done
