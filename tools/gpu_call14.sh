#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "wgrad or fold or conv_step" 2>&1 | tail -8
{
echo "--- default"; timeout 200 python tools/wgrad_probe.py 10 2>&1
echo "--- no stack"; FOSVOS_WG_NO_STACK=1 timeout 200 python tools/wgrad_probe.py 10 2>&1
echo "--- two waves"; FOSVOS_WG_TWO_WAVES=1 timeout 200 python tools/wgrad_probe.py 10 2>&1
} > gpurun_out/wg_stack.log 2>&1
cat gpurun_out/wg_stack.log
