#!/bin/bash
# round 2, state "p": ncu --set full of the small-map bx3 convolution (256 -> 256 at 2.2 K and 500 rows) and of wgrad_tc_kernel (L0 96 -> 96)
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2p
mkdir -p $O
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_bx3_kernel -s 2 -c 1 -o $O/bx3_L3_256 python scripts/dev_one_layer.py 256 256 bx3 f32 fwd 2200 > $O/ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_bx3_kernel -s 2 -c 1 -o $O/bx3_L4_256 python scripts/dev_one_layer.py 256 256 bx3 f32 fwd 500 > $O/ncu2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel -s 2 -c 1 -o $O/wgrad_L0_96 python scripts/dev_one_layer.py 96 96 bx3 f32 bwd > $O/ncu3.log 2>&1
tail -2 $O/ncu1.log $O/ncu2.log $O/ncu3.log
ls -la $O
