#!/bin/bash
# PDL A/B: quick bench (main job only) with and without programmatic dependent launch on the conv kernels
mkdir -p gpurun_out
Q="--parity 0 --gpu-reference 0 --config3 0 --config4 0 --fp32-modes 0"
for v in 0 1; do
  FOSVOS_PDL=$v timeout 600 python bench.py $Q > gpurun_out/bench_pdl$v.json 2> gpurun_out/bench_pdl$v.err; echo "pdl=$v rc=$?"
  python - <<PY
import json
d = json.loads(open('gpurun_out/bench_pdl$v.json').read().strip().splitlines()[-1])
print('PDL=$v', d['value'], d['config'].get('x'), d['finetune_s_per_sequence'], d['inference_fps'], d['clocks'])
PY
done
FOSVOS_PDL=1 timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -3
