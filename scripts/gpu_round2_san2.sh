#!/bin/bash
# round 2: compute-sanitizer racecheck (shared-memory hazards) and synccheck over one native-driver step
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2san
mkdir -p $O
timeout 40 compute-sanitizer --tool racecheck --print-limit 5 python scripts/dev_sanitize_r2.py > $O/racecheck.txt 2>&1
tail -n 8 $O/racecheck.txt
timeout 30 compute-sanitizer --tool synccheck --print-limit 5 python scripts/dev_sanitize_r2.py > $O/synccheck.txt 2>&1
tail -n 4 $O/synccheck.txt
