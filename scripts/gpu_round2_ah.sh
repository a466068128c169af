#!/bin/bash
# round 2, state "ah": run-to-run spread of the StepProgram / NativeStep noise gates (3 repetitions)
cd "$(dirname "$0")/.." || exit 1
for i in 1 2 3; do timeout 600 python -m pytest tests/test_zz_gpu_step_program.py -q -m gpu --timeout=500 -s -k "step_program_matches or native_step_matches" 2>&1 | grep "vs exact\|passed\|failed"; done
