#!/bin/bash
mkdir -p gpurun_out
timeout 420 python -m pytest tests -q -m gpu --timeout 120 -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^E  +(Assert|assert|Runtime|Type|Attr|Value|Key|Index|Name)|^FAILED" gpurun_out/t_all.log | cut -c1-300 | head -30
timeout 1200 python bench.py --iters 500 --steps 2 --warmup 3 > gpurun_out/bench_tmp.json 2> gpurun_out/bench_tmp.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_tmp.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_tmp.json'))
print({k:d[k] for k in ('value','ms_per_step','inference_fps','finetune_s_per_sequence','finetune_tflops','gpu_launches')}, d['e2e'], d['roofline']['frac'], d['roofline_side_chain']['frac'])
PY
timeout 120 python tools/profile_step.py 3 8 ft > gpurun_out/plain.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_ft.csv python tools/profile_step.py 3 8 ft > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches ft rc=$?"
