#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 600 -p no:cacheprovider -x -k "side" > gpurun_out/t_side.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^E  " gpurun_out/t_side.log | head -20
python tools/side_probe.py 8 2>&1 | tail -8
python tools/side_probe.py 1 2>&1 | tail -8
python tools/side_probe.py 32 2>&1 | tail -8
