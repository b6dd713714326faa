#!/bin/bash
for swap in 0 1; do
echo "== FOSVOS_WG_C8_SWAP=$swap"
FOSVOS_WG_C8_SWAP=$swap timeout 200 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 60 -p no:cacheprovider -k "wgrad or fold" 2>&1 | grep -E "passed|failed|^E  +(Assert|assert)|^FAILED" | cut -c1-200 | head -12
done
PROBE_SPLITS=1 timeout 200 python tools/wgrad_probe.py 1 2>&1 | cut -c1-120 | head -3
