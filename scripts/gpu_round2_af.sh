#!/bin/bash
# round 2, state "af": BatchNorm statistics + apply in one launch on the coarse levels (grid barrier): tests + bench on / off
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2af
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_nets.py tests/test_zz_gpu_step_program.py tests/test_gpu_conv.py -q -m gpu --timeout=900 -x 2>&1 | tail -4
timeout 400 python bench.py --no-cpu-baseline > $O/bench_small1.json 2> $O/bench_small1.err
LGS_BN_NO_SMALL=1 timeout 400 python bench.py --no-cpu-baseline > $O/bench_small0.json 2> $O/bench_small0.err
timeout 400 python bench.py --no-cpu-baseline > $O/bench_small1b.json 2> $O/bench_small1b.err
for f in $O/bench_*.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -2 | tr '\n' ' ') $(grep -o '"value": [0-9.]*' $f | head -1) $(grep -o '"loss": [0-9.]*' $f) $(grep -o '"step_ms": {[^}]*}' $f)"; done
