#!/bin/bash
mkdir -p gpurun_out
FOSVOS_WG_SIDE_ONE_PASS=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "wgrad or fold or conv_step" 2>&1 | tail -8
export PROBE_LAYERS=side2,side3,side4,side5
{
echo "--- default"; timeout 200 python tools/wgrad_probe.py 10 2>&1
echo "--- side one pass"; FOSVOS_WG_SIDE_ONE_PASS=1 timeout 200 python tools/wgrad_probe.py 10 2>&1
} > gpurun_out/wg_side.log 2>&1
cat gpurun_out/wg_side.log
