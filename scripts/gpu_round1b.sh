#!/bin/bash
# One GPU-box call: parity tests of the conv / net paths, then the in-process ablation of the new host / scheduling paths
# and a kernel-level + host-level profile of the step.  Everything is bounded by `timeout`; results land in gpurun_out/.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r1i
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_nets.py tests/test_gpu_conv.py -q -m gpu --timeout=300 -x 2>&1 | tail -25 > $O/pytest.txt
echo "pytest exit ${PIPESTATUS[0]}" >> $O/pytest.txt
LGS_BENCH_LAYERS=1 timeout 600 python scripts/dev_ablate.py > $O/ablate.txt 2> $O/ablate.err
timeout 200 python scripts/dev_profile.py > $O/kernels.txt 2>&1
timeout 200 python scripts/dev_hostprof.py 2>&1 | head -60 > $O/hostprof.txt
grep ABLATE $O/ablate.txt; tail -4 $O/pytest.txt; head -12 $O/kernels.txt; head -3 $O/hostprof.txt
