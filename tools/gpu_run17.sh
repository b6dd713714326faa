#!/bin/bash
mkdir -p gpurun_out
i=0
for cfg in "480 854 8 64 8" "480 854 64 64 8" "240 427 64 128 8" "240 427 128 16 8"; do
  i=$((i+1))
  ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 1 -c 1 -o gpurun_out/prof_c$i -f python tools/one_conv.py $cfg 3 > gpurun_out/ncu_c$i.log 2>&1
  echo "ncu $cfg rc=$?"
done
