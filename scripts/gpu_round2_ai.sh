#!/bin/bash
# round 2, state "ai": conv_nb for up to 1024 output channels (Res16UNet34D block8): tests + bench config 5
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2ai
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_conv_nb.py -q -m gpu -s --timeout=600 -k "vs_table_driven" 2>&1 | grep "512\|passed\|failed"
timeout 600 python bench.py --no-cpu-baseline --config 5 --steps 5 --warmup 3 > $O/bench_c5.json 2> $O/bench_c5.err
for f in $O/bench_*.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -2 | tr '\n' ' ') $(grep -o '"value": [0-9.]*' $f | head -1) $(grep -o '"loss": [0-9.]*' $f) $(grep -o '"step_ms": {[^}]*}' $f) $(grep -o '"frac_of_floor": [0-9.]*' $f)"; done
tail -n 3 $O/bench_c5.err
