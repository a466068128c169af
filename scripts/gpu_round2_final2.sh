#!/bin/bash
# round 2, final verification: full GPU suite, smoke, default bench twice, reference arm
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2final2
mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu --timeout=900 2>&1 | tail -5 > $O/pytest_gpu.txt; tail -3 $O/pytest_gpu.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err
timeout 900 python bench.py --no-cpu-baseline > $O/bench_default2.json 2> $O/bench_default2.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
for f in $O/bench_*.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -2 | tr '\n' ' ') $(grep -o '"value": [0-9.]*' $f | head -1) $(grep -o '"step_ms": {[^}]*}' $f) $(grep -o '"frac_of_floor": [0-9.]*' $f)"; done
