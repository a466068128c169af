#!/bin/bash
# round 2, state "b": bf16x3 conv kernel as the default algo — full GPU test-suite, bench, launch list
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2d
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu --timeout=300 -x -rf 2>&1 | tail -30 > $O/pytest.txt
echo "pytest exit ${PIPESTATUS[0]}" >> $O/pytest.txt
timeout 300 python bench.py --no-cpu-baseline > $O/bench_bx3.json 2> $O/bench_bx3.err
timeout 300 python bench.py --no-cpu-baseline --algo tc > $O/bench_tc.json 2> $O/bench_tc.err
tail -8 $O/pytest.txt
for f in $O/bench_*.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -2 | tr '\n' ' ') $(grep -o '"loss": [0-9.]*' $f)"; done
