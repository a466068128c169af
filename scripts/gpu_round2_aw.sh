#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
timeout 600 python -m pytest tests/test_gpu_loss.py tests/test_zz_gpu_step_program.py tests/test_gpu_nets.py -q -m gpu --timeout=600 -s 2>&1 | grep "^\[\|^\.\[" | cut -c1-260
