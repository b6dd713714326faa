#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider -x > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^E  " gpurun_out/t_all.log | head -20
python tools/side_probe.py 8 2>&1 | tail -6
python tools/side_probe.py 1 2>&1 | tail -6 | head -2
timeout 600 python tools/wgrad_probe.py 1 > gpurun_out/wgrad_probe.log 2>&1; echo "wgrad rc=$?"; cat gpurun_out/wgrad_probe.log
timeout 1200 python bench.py --iters 200 --steps 1 --warmup 3 > gpurun_out/bench_tmp.json 2> gpurun_out/bench_tmp.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_tmp.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_tmp.json'))
print({k:d[k] for k in ('value','ms_per_step','inference_fps','finetune_s_per_sequence','finetune_tflops','gpu_launches')}, d['e2e'], d['roofline']['frac'], d['roofline_side_chain']['frac'])
PY
