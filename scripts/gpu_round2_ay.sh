#!/bin/bash
# round 2: wgrad CTA granularity with the deferred level-0 wgrads in place (one box)
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2ay
mkdir -p $O
for w in 4 6 8 3 4; do
  LGS_WGRAD_WAVES=$w timeout 400 python bench.py --no-cpu-baseline --warmup 3 > $O/bench_w${w}_$RANDOM.json 2> $O/err.txt
done
for f in $O/bench_*.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -2 | tr '\n' ' ')"; done
