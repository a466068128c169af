#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r1p
mkdir -p $O
for i in 1 2 3 4; do timeout 300 python bench.py --no-cpu-baseline > $O/bench_$i.json 2> $O/bench_$i.err; done
for f in $O/bench_?.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -2 | tr '\n' ' ') $(grep -o '"loss": [0-9.]*' $f) $(grep -o '"samples": [0-9]*' $f | tr '\n' ' ')"; done
