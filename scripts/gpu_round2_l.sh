#!/bin/bash
# round 2, state "l": ncu --set full of conv_nb_kernel (96 -> 96, 149 106 rows)
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2l
mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_nb_kernel -s 5 -c 1 -o $O/conv_nb_L0_96 \
  python -m pytest tests/test_gpu_conv_nb.py -q -m gpu -s -k full_size > $O/ncu.log 2>&1
tail -5 $O/ncu.log
ls -la $O
