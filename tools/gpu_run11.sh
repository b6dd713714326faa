#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider -x > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^E  " gpurun_out/t_all.log | head -20
timeout 300 python __graft_entry__.py smoke > gpurun_out/t_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/t_smoke.log
timeout 1200 python bench.py --iters 500 --steps 2 --warmup 3 > gpurun_out/bench_r01c.json 2> gpurun_out/bench_r01c.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r01c.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r01c.json'))
print({k:d[k] for k in ('value','ms_per_step','inference_fps','finetune_s_per_sequence','finetune_tflops','gpu_launches')}, d['e2e'], d['roofline']['frac'], d['roofline_side_chain'], d['clocks'], d['cpu_baseline']['value'])
PY
python tools/profile_step.py 3 8 ft > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_ft.csv python tools/profile_step.py 3 8 ft > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches ft rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_inf.csv python tools/profile_step.py 3 8 inf > gpurun_out/ncu_launches2.log 2>&1
echo "ncu launches inf rc=$?"
