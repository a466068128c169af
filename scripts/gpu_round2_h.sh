#!/bin/bash
# round 2, state "h": bench configs 2 / 3 / 5 on one GPU (native driver), reference arm on the full scene, new tests
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2j
mkdir -p $O
timeout 400 python bench.py > $O/bench_c2.json 2> $O/bench_c2.err
timeout 300 python bench.py --no-cpu-baseline --config 3 > $O/bench_c3.json 2> $O/bench_c3.err
timeout 300 python bench.py --no-cpu-baseline --config 3 --driver facade > $O/bench_c3_facade.json 2> $O/bench_c3_facade.err
timeout 400 python bench.py --no-cpu-baseline --config 5 --steps 5 --warmup 3 > $O/bench_c5.json 2> $O/bench_c5.err
timeout 300 python bench.py --no-cpu-baseline --config 4 > $O/bench_c4_n1.json 2> $O/bench_c4_n1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_nets.py -q -m gpu --timeout=500 -k "bf16 or collate or augmentation" 2>&1 | tail -5
for f in $O/bench_*.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -2 | tr '\n' ' ') $(grep -o '"value": [0-9.]*' $f | head -1) $(grep -o '"loss": [0-9.]*' $f) $(grep -o '"frac": [0-9.]*' $f | head -1) $(grep -o '"frac_of_floor": [0-9.]*' $f)"; done
tail -3 $O/bench_c3.err $O/bench_c5.err
