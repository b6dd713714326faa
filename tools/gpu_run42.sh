#!/bin/bash
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none -k regex:"conv3x3_tc_kernel" -s 17 -c 17 -o gpurun_out/prof_fwd17 -f python tools/profile_step.py 2 16 inf > gpurun_out/ncu_fwd17.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_fwd17.log
