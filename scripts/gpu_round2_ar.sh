#!/bin/bash
# round 2, state "ar": A/B zero-fill kernel vs memset node on one box
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2ar
mkdir -p $O
timeout 400 python bench.py --no-cpu-baseline > $O/bench_kernel.json 2> $O/err.txt
LGS_ZERO_MEMSET=1 timeout 400 python bench.py --no-cpu-baseline > $O/bench_memset.json 2> $O/err.txt
timeout 400 python bench.py --no-cpu-baseline > $O/bench_kernel2.json 2> $O/err.txt
LGS_ZERO_MEMSET=1 timeout 400 python bench.py --no-cpu-baseline > $O/bench_memset2.json 2> $O/err.txt
for f in $O/bench_*.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -2 | tr '\n' ' ') $(grep -o '"step_ms": {[^}]*}' $f)"; done
