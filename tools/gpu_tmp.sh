#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_network.py -q -m gpu --timeout 100 -p no:cacheprovider -k "side or forward or infer or predict or mask" 2>&1 | tail -1
echo -n "auto rows: "; timeout 100 python tools/side_sep_probe.py 2>&1 | tail -1
for r in 10 12 14 16 20 24 30; do echo -n "rows=$r: "; FOSVOS_SIDE_SEP2_ROWS=$r timeout 100 python tools/side_sep_probe.py 5 16 2>&1 | tail -1; done
