#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"side_heads|side_upsample|bal_loss_fwd" -s 6 -c 3 -o gpurun_out/prof_side -f python tools/side_probe.py 8 > gpurun_out/ncu_side.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_side.log
