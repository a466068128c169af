#!/bin/bash
# round 2, state "k": first run of the neighbourhood-cache convolution (nbplan + conv_nb)
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2k
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_conv_nb.py -q -m gpu -s --timeout=600 2>&1 | tail -60 > $O/pytest_nb.txt
cat $O/pytest_nb.txt | tail -45
