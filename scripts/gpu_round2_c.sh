#!/bin/bash
# round 2, state "c": native step driver — its GPU tests, then bench native vs facade
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2e
mkdir -p $O
timeout 600 python -m pytest tests/test_zz_gpu_step_program.py -q -m gpu --timeout=300 -x -rf -s 2>&1 | tail -40 > $O/pytest_native.txt
echo "pytest exit ${PIPESTATUS[0]}" >> $O/pytest_native.txt
timeout 300 python bench.py --no-cpu-baseline > $O/bench_native.json 2> $O/bench_native.err
timeout 300 python bench.py --no-cpu-baseline --driver facade > $O/bench_facade.json 2> $O/bench_facade.err
timeout 300 python bench.py --no-cpu-baseline --warmup 25 > $O/bench_native_w25.json 2> $O/bench_native_w25.err
tail -25 $O/pytest_native.txt
tail -3 $O/bench_native.err
for f in $O/bench_*.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -2 | tr '\n' ' ') $(grep -o '"loss": [0-9.]*' $f) $(grep -o '"gpu_launches": [0-9]*' $f)"; done
