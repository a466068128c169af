#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r1l
mkdir -p $O
timeout 1200 python -m pytest tests -q -m gpu --timeout=300 -rf 2>&1 | tail -60 > $O/pytest.txt
echo "pytest exit ${PIPESTATUS[0]}" >> $O/pytest.txt
timeout 300 python bench.py --no-cpu-baseline > $O/bench_a.json 2> $O/bench_a.err
LGS_ATEN_CE=1 timeout 300 python bench.py --no-cpu-baseline > $O/bench_atence.json 2> $O/bench_atence.err
LGS_STAGE_PRIORITY=0 timeout 300 python bench.py --no-cpu-baseline > $O/bench_noprio.json 2> $O/bench_noprio.err
timeout 300 python bench.py --no-cpu-baseline > $O/bench_b.json 2> $O/bench_b.err
timeout 300 python scripts/dev_hosttime.py > $O/hosttime.txt 2>&1
LGS_OVERLAP_WGRAD=0 timeout 200 python scripts/dev_profile.py > $O/kernels_nooverlap.txt 2>&1
tail -8 $O/pytest.txt; head -24 $O/kernels_nooverlap.txt; cat $O/hosttime.txt
for f in $O/bench_*.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -2 | tr '\n' ' ') $(grep -o '"loss": [0-9.]*' $f)"; done
