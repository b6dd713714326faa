#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "row_stack or conv3x3_forward or fused_maxpool or pool_only or mask_accumulate" 2>&1 | tail -15
{
timeout 200 python tools/stack_probe.py 2>&1
FOSVOS_TC_NO_STACK=1 timeout 200 python tools/stack_probe.py 2>&1
} > gpurun_out/stack_probe.log 2>&1
cat gpurun_out/stack_probe.log
