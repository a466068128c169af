#!/bin/bash
# round 2, state "ag": weight operands of the late layers on the side stream, label count with 8 loads in flight: tests + bench
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2ag
mkdir -p $O
timeout 1200 python -m pytest tests/test_zz_gpu_step_program.py tests/test_gpu_loss.py tests/test_gpu_nets.py -q -m gpu --timeout=900 2>&1 | tail -4
timeout 400 python bench.py --no-cpu-baseline > $O/bench_split1.json 2> $O/bench_split1.err
LGS_SPLIT_PREP=0 timeout 400 python bench.py --no-cpu-baseline > $O/bench_split0.json 2> $O/bench_split0.err
timeout 400 python bench.py --no-cpu-baseline > $O/bench_split1b.json 2> $O/bench_split1b.err
for f in $O/bench_*.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -2 | tr '\n' ' ') $(grep -o '"value": [0-9.]*' $f | head -1) $(grep -o '"loss": [0-9.]*' $f) $(grep -o '"step_ms": {[^}]*}' $f)"; done
