#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider -x > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^E  " gpurun_out/t_all.log | head -20
timeout 1200 python bench.py --iters 500 --steps 2 --warmup 3 > gpurun_out/bench_tmp.json 2> gpurun_out/bench_tmp.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_tmp.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_tmp.json'))
print({k:d[k] for k in ('value','ms_per_step','inference_fps','finetune_s_per_sequence','finetune_tflops','gpu_launches')}, d['e2e'], d['roofline']['frac'], d['roofline_side_chain']['frac'])
PY
