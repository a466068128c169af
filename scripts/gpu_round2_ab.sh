#!/bin/bash
# round 2, state "ab": GPU timeline of the step (final kernels, PDL on) with a per-kernel trace
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2ab
mkdir -p $O
timeout 600 python scripts/dev_timeline.py $O/timeline.txt $O/trace.txt 2>&1 | tail -3
cat $O/timeline.txt | head -12
