#!/bin/bash
# round 2, state "o": whole step with the neighbourhood-cache convolution v3 (bench config 2, ncu launch list), then the full GPU suite
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2o
mkdir -p $O
timeout 400 python bench.py --no-cpu-baseline > $O/bench_c2.json 2> $O/bench_c2.err
for f in $O/bench_*.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -2 | tr '\n' ' ') $(grep -o '"value": [0-9.]*' $f | head -1) $(grep -o '"loss": [0-9.]*' $f) $(grep -o '"frac": [0-9.]*' $f | head -1) $(grep -o '"frac_of_floor": [0-9.]*' $f)"; done
tail -3 $O/bench_c2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches.csv python bench.py --profile-run --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
python profiles/summarize_launches.py $O/launches.csv 2>&1 | tail -40
timeout 1500 python -m pytest tests -q -m gpu --timeout=900 -x 2>&1 | tail -15 > $O/pytest_gpu.txt
tail -15 $O/pytest_gpu.txt
