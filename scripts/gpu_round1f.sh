#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r1m
mkdir -p $O
timeout 300 python scripts/dev_small_layers.py > $O/small_layers.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_nets.py -q -m gpu --timeout=300 -rf 2>&1 | tail -30 > $O/pytest.txt
timeout 300 python bench.py --no-cpu-baseline > $O/bench_a.json 2> $O/bench_a.err
LGS_TC_SPLIT_TARGET=74 timeout 300 python bench.py --no-cpu-baseline > $O/bench_t74.json 2> $O/bench_t74.err
LGS_TC_SPLIT_TARGET=296 timeout 300 python bench.py --no-cpu-baseline > $O/bench_t296.json 2> $O/bench_t296.err
cat $O/small_layers.txt; tail -5 $O/pytest.txt
for f in $O/bench_*.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -2 | tr '\n' ' ') $(grep -o '"loss": [0-9.]*' $f)"; done
