#!/bin/bash
# round 2: flakiness hunt — the noise-gated tests 5 times, failures with their assertion lines
cd "$(dirname "$0")/.." || exit 1
for i in 1 2 3 4 5; do
  timeout 600 python -m pytest tests/test_gpu_loss.py tests/test_zz_gpu_step_program.py tests/test_gpu_nets.py -q -m gpu --timeout=600 -x 2>&1 | grep -E "^E  |FAILED|passed|failed" | head -12
done
