#!/bin/bash
# round 2, state "ae": wgrad side stream at high priority vs default
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2ae
mkdir -p $O
timeout 400 python bench.py --no-cpu-baseline > $O/bench_default.json 2> $O/bench_default.err
LGS_SIDE_PRIORITY=1 timeout 400 python bench.py --no-cpu-baseline > $O/bench_sideprio.json 2> $O/bench_sideprio.err
timeout 400 python bench.py --no-cpu-baseline > $O/bench_default2.json 2> $O/bench_default2.err
LGS_SIDE_PRIORITY=1 timeout 400 python bench.py --no-cpu-baseline > $O/bench_sideprio2.json 2> $O/bench_sideprio2.err
for f in $O/bench_*.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -2 | tr '\n' ' ') $(grep -o '"value": [0-9.]*' $f | head -1) $(grep -o '"step_ms": {[^}]*}' $f)"; done
tail -n 3 $O/bench_sideprio.err
