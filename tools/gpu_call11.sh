#!/bin/bash
# pair-kernel bring-up: wgrad parity tests, then per-layer timing with and without the CTA-pair kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "wgrad" > gpurun_out/pair_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/pair_tests.log
tail -15 gpurun_out/pair_tests.log
timeout 300 python tools/wgrad_probe.py 20 > gpurun_out/pair_probe_on.log 2>&1; echo "rc=$?" >> gpurun_out/pair_probe_on.log
FOSVOS_WG_NO_CTA_PAIR=1 timeout 300 python tools/wgrad_probe.py 20 > gpurun_out/pair_probe_off.log 2>&1; echo "rc=$?" >> gpurun_out/pair_probe_off.log
echo "--- pair on"; cat gpurun_out/pair_probe_on.log; echo "--- pair off"; cat gpurun_out/pair_probe_off.log
