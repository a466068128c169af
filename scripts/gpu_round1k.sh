#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r1r
mkdir -p $O
timeout 100 python -m pytest tests -q -m gpu --timeout=120 -rf 2>&1 | tail -15 > $O/pytest.txt
echo "pytest exit ${PIPESTATUS[0]}" >> $O/pytest.txt
timeout 60 python bench.py --no-cpu-baseline > $O/bench.json 2> $O/bench.err
tail -4 $O/pytest.txt; wc -l $O/bench.json; echo "$(grep -o '"ms_per_step": [0-9.]*' $O/bench.json | head -2 | tr '\n' ' ')"
