#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "conv_step" 2>&1 | tail -3
timeout 200 python tools/step_probe.py 2>&1 | tail -2
FOSVOS_STEP_NO_VEC=1 timeout 200 python tools/step_probe.py 2>&1 | tail -2
