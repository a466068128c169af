#!/bin/bash
# round 2, state "s": ncu launch list of the step with conv_nb on every level
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2s
mkdir -p $O
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches.csv python bench.py --profile-run --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
python profiles/summarize_launches.py $O/launches.csv 2>&1 | head -12
