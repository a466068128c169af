#!/bin/bash
# round 2, state "q": GPU timeline of the native step (torch.profiler / CUPTI): busy vs idle, gaps per stream
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2q
mkdir -p $O
timeout 600 python scripts/dev_timeline.py $O/timeline.txt 2>&1 | tail -5
cat $O/timeline.txt
