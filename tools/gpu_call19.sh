#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_network.py tests/test_gpu_extras.py -x -q -m gpu 2>&1 | tail -6
Q="--parity 0 --gpu-reference 0 --config3 0 --config4 0 --fp32-modes 0"
for v in 1 0; do
  FOSVOS_POOL_ARG=$v timeout 600 python bench.py $Q > gpurun_out/bench_arg$v.json 2> gpurun_out/bench_arg$v.err; echo "pool_arg=$v rc=$?"
  python - <<PY
import json
d = json.loads(open('gpurun_out/bench_arg$v.json').read().strip().splitlines()[-1])
print('POOL_ARG=$v', d['value'], d['finetune_s_per_sequence'], d['inference_fps'], d['clocks']['sm_mhz'])
PY
done
