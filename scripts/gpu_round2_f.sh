#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2h
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_conv.py -q -m gpu --timeout=500 -rf -s -k "full_size_layer or conv_layer_parity" 2>&1 | grep -E "full size|layer |passed|failed|FAILED|Error" | tail -80 > $O/pytest.txt
cat $O/pytest.txt
