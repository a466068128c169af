#!/bin/bash
# One GPU-box call: parity tests of the conv / net paths, then bench.py with the new host / scheduling paths on, off and
# ablated one at a time.  Everything is bounded by `timeout`; results land in gpurun_out/.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r1i
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_nets.py tests/test_gpu_conv.py -q -m gpu --timeout=300 -x 2>&1 | tail -25 > $O/pytest.txt
echo "pytest exit ${PIPESTATUS[0]}" >> $O/pytest.txt
B="timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5"
LGS_BENCH_LAYERS=1 $B > $O/bench_on.json 2> $O/bench_on.err
ALL_OFF="LGS_FUSE_CONV_BN=0 LGS_OVERLAP_WGRAD=0 LGS_BATCH_PREP=0 LGS_TC_NO_BALANCE=1 LGS_BN_SCALAR=1 LGS_WGRAD_NO_BALANCE=1"
env $ALL_OFF LGS_BENCH_LAYERS=1 $B > $O/bench_off.json 2> $O/bench_off.err
for v in LGS_FUSE_CONV_BN=0 LGS_OVERLAP_WGRAD=0 LGS_BATCH_PREP=0 LGS_TC_NO_BALANCE=1 LGS_BN_SCALAR=1 LGS_WGRAD_NO_BALANCE=1; do
  env $v $B > $O/bench_${v%%=*}.json 2> $O/bench_${v%%=*}.err
done
grep -h -o '"ms_per_step": [0-9.]*' $O/bench_*.json | head -20
for f in $O/bench_*.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -2 | tr '\n' ' ')"; done > $O/summary.txt
cat $O/summary.txt; tail -5 $O/pytest.txt
