#!/bin/bash
# round 2, state "v": small-map conv_nb durations (ncu gpu__time_duration) for several split targets
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2v
mkdir -p $O
for cfg in "450 256 256" "2200 256 256" "2200 128 128" "8500 64 64" "8500 128 128"; do
  for tune in "nb_target_ctas=296" "nb_target_ctas=148" "nb_target_ctas=74" "nb_no_split=1"; do
    LGS_TUNE=$tune timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:conv_nb_kernel -s 3 -c 1 --csv python scripts/dev_nb_layer.py $cfg 2>/dev/null | grep conv_nb_kernel | awk -F'","' -v c="$cfg" -v t="$tune" '{print c, "|", t, "| grid", $(NF-6), "|", $NF}' | tr -d '"'
  done
done | tee $O/sweep.txt
