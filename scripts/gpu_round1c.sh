#!/bin/bash
# second GPU-box call of the session: full -m gpu suite, determinism probe, host-time breakdown, bench.py twice
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r1j
mkdir -p $O
timeout 1200 python -m pytest tests -q -m gpu --timeout=300 2>&1 | tail -40 > $O/pytest.txt
echo "pytest exit ${PIPESTATUS[0]}" >> $O/pytest.txt
timeout 300 python scripts/dev_determinism.py > $O/determinism.txt 2>&1
timeout 300 python scripts/dev_hosttime.py > $O/hosttime.txt 2>&1
LGS_OVERLAP_WGRAD=0 timeout 300 python scripts/dev_hosttime.py > $O/hosttime_nooverlap.txt 2>&1
timeout 300 python bench.py --no-cpu-baseline > $O/bench_a.json 2> $O/bench_a.err
timeout 300 python bench.py --no-cpu-baseline > $O/bench_b.json 2> $O/bench_b.err
LGS_OVERLAP_WGRAD=0 timeout 300 python bench.py --no-cpu-baseline > $O/bench_nooverlap.json 2> $O/bench_nooverlap.err
tail -6 $O/pytest.txt; cat $O/determinism.txt | head -40; cat $O/hosttime.txt $O/hosttime_nooverlap.txt
for f in $O/bench_*.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -2 | tr '\n' ' ') $(grep -o '"loss": [0-9.]*' $f)"; done
