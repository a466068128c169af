#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
timeout 1200 python -m pytest tests/test_gpu_loss.py tests/test_zz_gpu_step_program.py tests/test_gpu_nets.py -q -m gpu --timeout=600 2>&1 | grep -B30 "short test summary" | tail -50
timeout 1200 python -m pytest tests/test_gpu_loss.py tests/test_zz_gpu_step_program.py tests/test_gpu_nets.py -q -m gpu --timeout=600 2>&1 | tail -5
