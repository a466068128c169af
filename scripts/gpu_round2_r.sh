#!/bin/bash
# round 2, state "r": neighbourhood-cache convolution on the coarse levels (reduction split over CTAs)
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2r
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_conv_nb.py -q -m gpu -s --timeout=600 2>&1 | tail -60 > $O/pytest_nb.txt
grep -v "^\.\[conv_nb\|^\[conv_nb" $O/pytest_nb.txt | tail -30
timeout 400 python bench.py --no-cpu-baseline > $O/bench_c2.json 2> $O/bench_c2.err
for f in $O/bench_*.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -2 | tr '\n' ' ') $(grep -o '"value": [0-9.]*' $f | head -1) $(grep -o '"loss": [0-9.]*' $f) $(grep -o '"frac_of_floor": [0-9.]*' $f)"; done
tail -3 $O/bench_c2.err
