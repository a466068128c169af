#!/bin/bash
# round 2, state "u": ncu --set full of conv_nb_kernel on a coarse level (447 rows, 256 -> 256, reduction split over CTAs)
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2u
mkdir -p $O
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_nb_kernel -s 3 -c 1 -o $O/nb_L4_256 python scripts/dev_nb_layer.py 450 256 256 > $O/ncu1.log 2>&1
tail -n 2 $O/ncu1.log
