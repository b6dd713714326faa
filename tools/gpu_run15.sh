#!/bin/bash
mkdir -p gpurun_out
for pdl in 0 1; do
FOSVOS_PDL=$pdl timeout 1200 python bench.py --iters 200 --steps 1 --warmup 2 > gpurun_out/bench_pdl$pdl.json 2> gpurun_out/bench_pdl$pdl.err; echo "bench pdl=$pdl rc=$?"; tail -3 gpurun_out/bench_pdl$pdl.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_pdl$pdl.json'))
print({k:d[k] for k in ('value','ms_per_step','inference_fps','finetune_s_per_sequence','finetune_tflops')}, d['roofline']['frac'], d['roofline_side_chain']['frac'])
PY
done
FOSVOS_PDL=1 python -m pytest tests/test_gpu_network.py -q -m gpu --timeout 600 -p no:cacheprovider -x 2>&1 | tail -2
