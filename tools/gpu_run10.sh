#!/bin/bash
mkdir -p gpurun_out
python tools/one_conv.py 480 854 64 64 1 3 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 1 -c 1 -o gpurun_out/prof_conv64 -f python tools/one_conv.py 480 854 64 64 1 3 > gpurun_out/ncu_conv64.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_conv64.log
