#!/bin/bash
# round 2: compute-sanitizer memcheck over one native-driver step (round-2 kernels)
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2san
mkdir -p $O
timeout 105 compute-sanitizer --tool memcheck --print-limit 5 python scripts/dev_sanitize_r2.py > $O/memcheck.txt 2>&1
tail -n 6 $O/memcheck.txt
