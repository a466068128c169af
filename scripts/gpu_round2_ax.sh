#!/bin/bash
# round 2: measured values of the two chaos-sensitive gates, 8 repetitions
cd "$(dirname "$0")/.." || exit 1
for i in 1 2 3 4 5 6 7 8; do
  timeout 300 python -m pytest tests/test_gpu_nets.py tests/test_zz_gpu_step_program.py -q -m gpu --timeout=300 -s -k "fused_node_overlap or native_step_trains" 2>&1 | grep -E "^\.?\[fast|^\.?\[NativeStep trains|passed|failed" | cut -c1-330
done
