#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 100 -p no:cacheprovider -k wgrad 2>&1 | tail -1
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_r01e.json 2> gpurun_out/bench_r01e.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r01e.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r01e.json'))
print({k:d[k] for k in ('value','ms_per_step','inference_fps','finetune_s_per_sequence','finetune_tflops','gpu_launches')}, d['e2e'], d['roofline']['frac'], d['roofline_side_chain']['frac'], d['roofline_loss']['frac'], d['cpu_baseline']['value'], d['clocks'])
PY
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref.json
timeout 120 python tools/profile_step.py 3 8 ft > gpurun_out/plain.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_ft_r01e.csv python tools/profile_step.py 3 8 ft > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches ft rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_inf_r01e.csv python tools/profile_step.py 3 8 inf > gpurun_out/ncu_launches2.log 2>&1
echo "ncu launches inf rc=$?"
PROBE_SPLITS=1 timeout 300 python tools/wgrad_probe.py 1 2>&1 | cut -c1-120 | head -5
