#!/bin/bash
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"side_upsample_sep" -s 3 -c 1 -o gpurun_out/prof_side2 -f python tools/side_probe.py 8 > gpurun_out/ncu_side2.log 2>&1
echo "ncu rc=$?"
