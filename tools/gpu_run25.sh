#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 100 -p no:cacheprovider -k wgrad 2>&1 | tail -1
for dbg in 0 2 1; do
echo "== FOSVOS_WG_DEBUG=$dbg"
FOSVOS_WG_DEBUG=$dbg PROBE_SPLITS=1 timeout 300 python tools/wgrad_probe.py 1 2>&1 | cut -c1-120
done
