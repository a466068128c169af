#!/bin/bash
# round 2, state "x": bench config 2 with per-layer table, permuted row order, configs 3 / 4(n1) / 5; reference arm
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2x
mkdir -p $O
LGS_BENCH_LAYERS=1 timeout 400 python bench.py --no-cpu-baseline > $O/bench_c2.json 2> $O/bench_c2.err
timeout 400 python bench.py --no-cpu-baseline --permute-rows > $O/bench_c2_permuted.json 2> $O/bench_c2_permuted.err
timeout 300 python bench.py --no-cpu-baseline --config 3 > $O/bench_c3.json 2> $O/bench_c3.err
timeout 300 python bench.py --no-cpu-baseline --config 4 > $O/bench_c4_n1.json 2> $O/bench_c4_n1.err
timeout 400 python bench.py --no-cpu-baseline --config 5 --steps 5 --warmup 3 > $O/bench_c5.json 2> $O/bench_c5.err
for f in $O/bench_*.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -2 | tr '\n' ' ') $(grep -o '"value": [0-9.]*' $f | head -1) $(grep -o '"loss": [0-9.]*' $f) $(grep -o '"warmup_done": [0-9]*' $f) $(grep -o '"frac_of_floor": [0-9.]*' $f)"; done
grep LAYER $O/bench_c2.err | head -30
tail -2 $O/bench_c3.err $O/bench_c5.err $O/bench_c4_n1.err
