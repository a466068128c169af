#!/bin/bash
# round 2, state "z": bench config 2 (x2, per-step times), config 4 n=1, nb tests (async plan status)
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2z
mkdir -p $O
timeout 400 python bench.py --no-cpu-baseline > $O/bench_c2.json 2> $O/bench_c2.err
timeout 400 python bench.py --no-cpu-baseline > $O/bench_c2b.json 2> $O/bench_c2b.err
timeout 400 python bench.py --no-cpu-baseline > $O/bench_c2c.json 2> $O/bench_c2c.err
timeout 400 python bench.py --no-cpu-baseline > $O/bench_c2d.json 2> $O/bench_c2d.err
timeout 300 python bench.py --no-cpu-baseline --config 4 > $O/bench_c4_n1.json 2> $O/bench_c4_n1.err
for f in $O/bench_*.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -2 | tr '\n' ' ') $(grep -o '"value": [0-9.]*' $f | head -1) $(grep -o '"step_ms": {[^}]*}' $f) $(grep -o '"warmup_done": [0-9]*' $f) $(grep -o '"frac_of_floor": [0-9.]*' $f)"; done
timeout 900 python -m pytest tests/test_gpu_conv_nb.py tests/test_zz_gpu_step_program.py -q -m gpu --timeout=600 2>&1 | tail -3
