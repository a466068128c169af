#!/bin/bash
# round 2, state "as": zero-fill kernels also ahead of wgrad / colsum: tests + bench
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2as
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_conv.py tests/test_zz_gpu_step_program.py tests/test_gpu_nets.py -q -m gpu --timeout=600 2>&1 | tail -2
timeout 400 python bench.py --no-cpu-baseline > $O/bench_a.json 2> $O/err.txt
LGS_ZERO_MEMSET=1 timeout 400 python bench.py --no-cpu-baseline > $O/bench_memset.json 2> $O/err.txt
timeout 400 python bench.py --no-cpu-baseline > $O/bench_b.json 2> $O/err.txt
for f in $O/bench_*.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -2 | tr '\n' ' ') $(grep -o '"loss": [0-9.]*' $f) $(grep -o '"gpu_launches": [0-9]*' $f)"; done
