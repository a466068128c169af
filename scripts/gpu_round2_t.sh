#!/bin/bash
# round 2, state "t": nb tests, bench, launch list (split search by waves, multi-block bin scan)
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2t
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_conv_nb.py -q -m gpu -s --timeout=600 2>&1 | tail -60 > $O/pytest_nb.txt
grep -v "^\.\[conv_nb\|^\[conv_nb" $O/pytest_nb.txt | tail -14
timeout 400 python bench.py --no-cpu-baseline > $O/bench_c2.json 2> $O/bench_c2.err
for f in $O/bench_*.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -2 | tr '\n' ' ') $(grep -o '"value": [0-9.]*' $f | head -1) $(grep -o '"loss": [0-9.]*' $f) $(grep -o '"frac_of_floor": [0-9.]*' $f)"; done
tail -3 $O/bench_c2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches.csv python bench.py --profile-run --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
python profiles/summarize_launches.py $O/launches.csv 2>&1 | head -14
