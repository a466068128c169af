#!/bin/bash
# round 2, state "ao": big level-0 decoder wgrads deferred to the coarse encoder phase (1 GPU experiment), A/B on one box
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2ao
mkdir -p $O
timeout 400 python bench.py --no-cpu-baseline > $O/bench_base.json 2> $O/err.txt
LGS_DEFER_WGRAD=1 timeout 400 python bench.py --no-cpu-baseline > $O/bench_defer.json 2> $O/err.txt
LGS_DEFER_WGRAD=1 LGS_WGRAD_WAVES=1 timeout 400 python bench.py --no-cpu-baseline > $O/bench_defer_w1.json 2> $O/err.txt
timeout 400 python bench.py --no-cpu-baseline > $O/bench_base2.json 2> $O/err.txt
LGS_DEFER_WGRAD=1 timeout 400 python bench.py --no-cpu-baseline > $O/bench_defer2.json 2> $O/err.txt
for f in $O/bench_*.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -2 | tr '\n' ' ') $(grep -o '"loss": [0-9.]*' $f) $(grep -o '"step_ms": {[^}]*}' $f)"; done
