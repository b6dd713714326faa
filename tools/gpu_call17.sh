#!/bin/bash
mkdir -p gpurun_out
FOSVOS_TC_STACK_ALL=1 timeout 200 python tools/stack_probe.py 2>&1 | grep dgrad
timeout 400 ncu --set full --clock-control none --import-source on -k regex:stack_tc -s 2 -c 1 -o gpurun_out/prof_stack -f python tools/stack_probe.py > gpurun_out/ncu_stack.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/prof_stack.ncu-rep
