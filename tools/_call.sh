#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "wgrad" 2>&1 | tail -3
export PROBE_LAYERS=conv1_1
timeout 100 python tools/wgrad_probe.py 10 2>&1 | head -1
FOSVOS_WG_C8_THREE_MMAS=1 timeout 100 python tools/wgrad_probe.py 10 2>&1 | head -1
