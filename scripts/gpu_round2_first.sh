#!/bin/bash
# FIRST GPU call of round 2: measure what round 1 could only verify on the CPU (native binding as default, StepProgram,
# host micro-optimisations) next to the last measured state (profiles/r1_j_bench_n1.json: 15.99 ms/step).
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2a
mkdir -p $O
timeout 600 python -m pytest tests -q -m gpu --timeout=300 -rf 2>&1 | tail -30 > $O/pytest.txt
echo "pytest exit ${PIPESTATUS[0]}" >> $O/pytest.txt
timeout 300 python bench.py --no-cpu-baseline > $O/bench_default.json 2> $O/bench_default.err
LGS_FAST_BIND=0 timeout 300 python bench.py --no-cpu-baseline > $O/bench_ctypes.json 2> $O/bench_ctypes.err
timeout 300 python bench.py --no-cpu-baseline --step-program > $O/bench_program.json 2> $O/bench_program.err
timeout 300 python bench.py --no-cpu-baseline > $O/bench_default2.json 2> $O/bench_default2.err
timeout 300 python scripts/dev_hosttime.py > $O/hosttime.txt 2>&1
LGS_FAST_BIND=0 timeout 300 python scripts/dev_hosttime.py > $O/hosttime_ctypes.txt 2>&1
tail -6 $O/pytest.txt; cat $O/hosttime.txt $O/hosttime_ctypes.txt
for f in $O/bench_*.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -2 | tr '\n' ' ') $(grep -o '"loss": [0-9.]*' $f) $(grep -o '"binding": "[a-z]*' $f) $(grep -o '"driver": "[A-Za-z]*' $f)"; done
