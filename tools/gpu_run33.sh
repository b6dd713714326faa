#!/bin/bash
mkdir -p gpurun_out
for pdl in 0 1; do
FOSVOS_PDL=$pdl timeout 600 python bench.py --iters 200 --steps 1 --warmup 2 > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err; echo "pdl=$pdl rc=$?"; tail -2 gpurun_out/bench_ab.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_ab.json'))
print('   ', d['finetune_s_per_sequence'], d['inference_fps'], d['roofline']['frac'], d['roofline_side_chain']['frac'], d['roofline_loss']['frac'])
PY
done
FOSVOS_PDL=1 timeout 400 python -m pytest tests/test_gpu_network.py -q -m gpu --timeout 120 -p no:cacheprovider 2>&1 | tail -1
