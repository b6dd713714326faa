#!/bin/bash
mkdir -p gpurun_out
Q="--parity 0 --gpu-reference 0 --config3 0 --config4 0 --fp32-modes 0"
for v in epilogue pool; do
  FOSVOS_BWD_FANIN=$v timeout 600 python bench.py $Q > gpurun_out/bench_fanin_$v.json 2> gpurun_out/bench_fanin_$v.err; echo "fanin=$v rc=$?"
  python - <<PY
import json
d = json.loads(open('gpurun_out/bench_fanin_$v.json').read().strip().splitlines()[-1])
print('FANIN=$v', d['value'], d['finetune_s_per_sequence'], d['inference_fps'], d['clocks']['sm_mhz'])
PY
done
FOSVOS_BWD_FANIN=pool timeout 600 python -m pytest tests/test_gpu_network.py -x -q -m gpu 2>&1 | tail -2
