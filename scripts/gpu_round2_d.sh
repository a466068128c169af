#!/bin/bash
# round 2, state "d": full GPU suite with the native driver + ncu launch list of the native step
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2f
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu --timeout=300 -rf 2>&1 | tail -40 > $O/pytest.txt
echo "pytest exit ${PIPESTATUS[0]}" >> $O/pytest.txt
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1300 --csv --log-file $O/launches.csv python bench.py --profile-run --steps 3 --warmup 6 --no-cpu-baseline > $O/ncu_bench.log 2>&1
tail -12 $O/pytest.txt
python profiles/summarize_launches.py $O/launches.csv 2>/dev/null | head -40
