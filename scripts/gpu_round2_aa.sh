#!/bin/bash
# round 2, state "aa": wgrad with table entries two steps ahead (tests + bench), ncu --set full of the remaining kernel families
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2aa
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_conv.py -q -m gpu --timeout=600 -k "wgrad or parity or full" 2>&1 | tail -3
LGS_BENCH_LAYERS=1 timeout 400 python bench.py --no-cpu-baseline > $O/bench_c2.json 2> $O/bench_c2.err
for f in $O/bench_*.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -2 | tr '\n' ' ') $(grep -o '"value": [0-9.]*' $f | head -1) $(grep -o '"step_ms": {[^}]*}' $f) $(grep -o '"frac_of_floor": [0-9.]*' $f)"; done
grep "LAYER wgrad" $O/bench_c2.err | head -8
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bn_ -s 8 -c 4 -o $O/bn_L0_96 python scripts/dev_bn.py > $O/ncu_bn.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:clip_ce_tc_kernel -s 3 -c 1 -o $O/clip_ce_tc python scripts/dev_loss.py > $O/ncu_loss.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:insert_kernel|claim_kernel|kmap_kernel|nb_plan_kernel" -c 4 -o $O/coords python scripts/dev_nb_layer.py 150000 96 96 1 > $O/ncu_coords.log 2>&1
ls -la $O | tail -12
