#!/bin/bash
# round 2, state "aq": zero-fill kernel (PDL) instead of memset nodes ahead of split-K convolutions: tests + A/B (pdl=0 keeps memset)
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2aq
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_conv_nb.py tests/test_gpu_conv.py -q -m gpu --timeout=600 -k "small or parity or split" 2>&1 | tail -2
timeout 400 python bench.py --no-cpu-baseline > $O/bench_a.json 2> $O/err.txt
timeout 400 python bench.py --no-cpu-baseline > $O/bench_b.json 2> $O/err.txt
for f in $O/bench_*.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -2 | tr '\n' ' ') $(grep -o '"loss": [0-9.]*' $f) $(grep -o '"step_ms": {[^}]*}' $f)"; done
