#!/bin/bash
# round 2, final state: full GPU suite, bench (default invocation), ncu launch list + summary, timeline
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2final
mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu --timeout=900 2>&1 | tail -6 > $O/pytest_gpu.txt; tail -4 $O/pytest_gpu.txt
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err
timeout 600 python bench.py --no-cpu-baseline --config 3 > $O/bench_c3.json 2> $O/bench_c3.err
for f in $O/bench_*.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -2 | tr '\n' ' ') $(grep -o '"value": [0-9.]*' $f | head -1) $(grep -o '"step_ms": {[^}]*}' $f) $(grep -o '"frac_of_floor": [0-9.]*' $f)"; done
timeout 600 python scripts/dev_timeline.py $O/timeline.txt $O/trace.txt > /dev/null 2>&1; head -5 $O/timeline.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches.csv python bench.py --profile-run --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
python profiles/summarize_launches.py $O/launches.csv 2>&1 | head -24
