#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r1o
mkdir -p $O
for i in 1 2 3; do timeout 300 python bench.py --no-cpu-baseline > $O/bench_$i.json 2> $O/bench_$i.err; done
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_ref.json 2> $O/bench_ref.err
for f in $O/bench_?.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -2 | tr '\n' ' ') $(grep -o '"loss": [0-9.]*' $f)"; done
cat $O/bench_ref.json | cut -c1-300
