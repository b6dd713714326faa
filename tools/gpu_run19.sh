#!/bin/bash
mkdir -p gpurun_out
i=0
for cfg in "240 427 128 128 1" "480 854 64 64 1" "30 54 512 512 1" "60 107 512 512 1"; do
  i=$((i+1))
  ncu --set full --clock-control none --import-source on -k regex:wgrad_tc -s 1 -c 1 -o gpurun_out/prof_w$i -f python tools/one_wgrad.py $cfg 3 > gpurun_out/ncu_w$i.log 2>&1
  echo "ncu $cfg rc=$?"
done
