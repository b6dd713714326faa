#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python bench.py --iters 500 --steps 2 --warmup 3 > gpurun_out/bench_r01d.json 2> gpurun_out/bench_r01d.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r01d.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r01d.json'))
print({k:d[k] for k in ('value','ms_per_step','inference_fps','finetune_s_per_sequence','finetune_tflops','gpu_launches')}, d['e2e'], d['roofline']['frac'], d['roofline_side_chain']['frac'], d['clocks'])
for l in d['roofline']['per_layer']: print(l)
PY
python tools/profile_step.py 3 8 ft > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_ft.csv python tools/profile_step.py 3 8 ft > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches ft rc=$?"
