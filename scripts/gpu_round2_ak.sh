#!/bin/bash
# round 2, state "ak": dgrad + identity-path gradient added in the convolution epilogue (lgs_conv_fwd4): tests + bench
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2ak
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_conv_nb.py tests/test_zz_gpu_step_program.py -q -m gpu --timeout=900 2>&1 | tail -4
timeout 400 python bench.py --no-cpu-baseline > $O/bench_c2.json 2> $O/bench_c2.err
timeout 400 python bench.py --no-cpu-baseline > $O/bench_c2b.json 2> $O/bench_c2b.err
for f in $O/bench_*.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -2 | tr '\n' ' ') $(grep -o '"value": [0-9.]*' $f | head -1) $(grep -o '"loss": [0-9.]*' $f) $(grep -o '"step_ms": {[^}]*}' $f) $(grep -o '"gpu_launches": [0-9]*' $f)"; done
