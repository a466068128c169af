#!/bin/bash
# round 2: bench after moving the step events out of the timed region (2 runs), then 8 runs WITHOUT the clock sampler
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2ba
mkdir -p $O
for i in 1 2; do timeout 400 python bench.py --no-cpu-baseline > $O/bench_$i.json 2> $O/err.txt; done
for i in 1 2 3 4 5 6 7 8; do LGS_BENCH_NO_CLOCKS=1 timeout 400 python bench.py --no-cpu-baseline > $O/bench_noclk_$i.json 2> $O/err.txt; done
for f in $O/bench_*.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -2 | tr '\n' ' ') $(grep -o '"max": [0-9.]*, "argmax": [0-9]*' $f) $(grep -o '"clocks": {[^}]*}' $f | head -1)"; done
