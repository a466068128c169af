#!/bin/bash
# round 2, state "ap": deferral variants (levels 0-1; spread over the encoder stages), A/B on one box
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2ap
mkdir -p $O
timeout 400 python bench.py --no-cpu-baseline > $O/bench_base.json 2> $O/err.txt
LGS_DEFER_WGRAD=1 timeout 400 python bench.py --no-cpu-baseline > $O/bench_defer_l0.json 2> $O/err.txt
LGS_DEFER_WGRAD=1 LGS_DEFER_LEVEL=1 timeout 400 python bench.py --no-cpu-baseline > $O/bench_defer_l01.json 2> $O/err.txt
LGS_DEFER_WGRAD=1 LGS_DEFER_SPREAD=1 timeout 400 python bench.py --no-cpu-baseline > $O/bench_defer_l0_spread.json 2> $O/err.txt
LGS_DEFER_WGRAD=1 LGS_DEFER_LEVEL=1 LGS_DEFER_SPREAD=1 timeout 400 python bench.py --no-cpu-baseline > $O/bench_defer_l01_spread.json 2> $O/err.txt
for f in $O/bench_*.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -2 | tr '\n' ' ') $(grep -o '"loss": [0-9.]*' $f) $(grep -o '"step_ms": {[^}]*}' $f)"; done
