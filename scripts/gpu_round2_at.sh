#!/bin/bash
# round 2, state "at": persistent seg_ce kernel, vectorised colsum: tests + bench
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2at
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_loss.py tests/test_zz_gpu_step_program.py tests/test_gpu_nets.py -q -m gpu --timeout=600 2>&1 | tail -2
timeout 400 python bench.py --no-cpu-baseline > $O/bench_a.json 2> $O/err.txt
timeout 400 python bench.py --no-cpu-baseline > $O/bench_b.json 2> $O/err.txt
for f in $O/bench_*.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -2 | tr '\n' ' ') $(grep -o '"loss": [0-9.]*' $f)"; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:seg_ce|colsum" -c 12 --csv python bench.py --profile-run --steps 1 --warmup 3 --no-cpu-baseline 2>/dev/null | grep -i "seg_ce\|colsum" | awk -F'","' '{print $5, $NF}' | tr -d '"' | tail -6
