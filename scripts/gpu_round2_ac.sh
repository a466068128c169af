#!/bin/bash
# round 2, state "ac": stem wgrad kernel (c_in 4, c_out 32, K 27): parity tests + bench
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2ac
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_nets.py -q -m gpu --timeout=600 -s -k "parity or nets or whole or unet" 2>&1 | grep -i "layer 3->32\|passed\|failed\|error" | head -12
LGS_BENCH_LAYERS=1 timeout 400 python bench.py --no-cpu-baseline > $O/bench_c2.json 2> $O/bench_c2.err
for f in $O/bench_*.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -2 | tr '\n' ' ') $(grep -o '"value": [0-9.]*' $f | head -1) $(grep -o '"step_ms": {[^}]*}' $f) $(grep -o '"frac_of_floor": [0-9.]*' $f)"; done
grep "LAYER wgrad" $O/bench_c2.err | head -9
