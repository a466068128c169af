#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2i
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_nets.py -q -m gpu --timeout=500 -rf -s -k "bf16 or collate or augmentation" 2>&1 | grep -vE "^\[layer|^\.\[layer" | tail -40 > $O/pytest.txt
cat $O/pytest.txt
