#!/bin/bash
# round 2, state "an": wgrad CTA granularity (LGS_WGRAD_WAVES = CTAs' worth of work per SM) on one box
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2an
mkdir -p $O
for w in 2 4 8 3 2; do
  LGS_WGRAD_WAVES=$w timeout 400 python bench.py --no-cpu-baseline > $O/bench_w${w}_$RANDOM.json 2> $O/err.txt
done
for f in $O/bench_*.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -2 | tr '\n' ' ') $(grep -o '"step_ms": {[^}]*}' $f)"; done
