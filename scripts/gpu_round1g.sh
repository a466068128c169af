#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r1n
mkdir -p $O
timeout 1200 python -m pytest tests -q -m gpu --timeout=300 -rf 2>&1 | tail -30 > $O/pytest.txt
echo "pytest exit ${PIPESTATUS[0]}" >> $O/pytest.txt
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err
LGS_BENCH_LAYERS=1 timeout 300 python bench.py --no-cpu-baseline > $O/bench2.json 2> $O/bench2_layers.txt
timeout 300 python __graft_entry__.py smoke > $O/smoke.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv python bench.py --profile-run --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_tc2 -s 2 -c 1 -o $O/conv_tc2_L1_balanced -f python scripts/dev_one_layer.py 96 96 tc f32 fwd 38500 > $O/ncu_l1.log 2>&1
timeout 600 python scripts/dev_configs.py > $O/configs.txt 2>&1
gzip -f $O/launches.csv
tail -5 $O/pytest.txt; cat $O/bench.json; tail -3 $O/smoke.txt; cat $O/configs.txt; tail -3 $O/ncu_l1.log
