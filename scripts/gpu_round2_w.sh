#!/bin/bash
# round 2, state "w": programmatic dependent launch on / off (bench config 2), full GPU suite with it on
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2w
mkdir -p $O
timeout 400 python bench.py --no-cpu-baseline > $O/bench_pdl1.json 2> $O/bench_pdl1.err
LGS_TUNE=pdl=0 timeout 400 python bench.py --no-cpu-baseline > $O/bench_pdl0.json 2> $O/bench_pdl0.err
timeout 400 python bench.py --no-cpu-baseline > $O/bench_pdl1b.json 2> $O/bench_pdl1b.err
for f in $O/bench_*.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -2 | tr '\n' ' ') $(grep -o '"value": [0-9.]*' $f | head -1) $(grep -o '"loss": [0-9.]*' $f) $(grep -o '"frac_of_floor": [0-9.]*' $f)"; done
tail -3 $O/bench_pdl1.err
timeout 1500 python -m pytest tests -q -m gpu --timeout=900 -x 2>&1 | tail -8 > $O/pytest_gpu.txt
tail -8 $O/pytest_gpu.txt
