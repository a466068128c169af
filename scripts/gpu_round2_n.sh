#!/bin/bash
# round 2, state "n": neighbourhood-cache convolution v2 (split at fill time, colour-rotated cache rows): tests, then ncu --set full
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2n
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_conv_nb.py -q -m gpu -s --timeout=600 2>&1 | tail -60 > $O/pytest_nb.txt
tail -45 $O/pytest_nb.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_nb_kernel -s 5 -c 1 -o $O/conv_nb_v3_L0_96 \
  python -m pytest tests/test_gpu_conv_nb.py -q -m gpu -s -k full_size > $O/ncu.log 2>&1
tail -3 $O/ncu.log
