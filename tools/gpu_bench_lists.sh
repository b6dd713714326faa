#!/bin/bash
# bench + ncu launch lists only (small outputs); the --set full captures are taken by gpu_validate.sh / separately
mkdir -p gpurun_out
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_r01i.json 2> gpurun_out/bench_r01i.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r01i.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r01i.json'))
print({k:d[k] for k in ('value','ms_per_step','inference_fps','finetune_s_per_sequence','finetune_tflops','gpu_launches')}, d['e2e'], d['roofline']['frac'], d['roofline_side_chain']['frac'], d['roofline_loss']['frac'], d['cpu_baseline']['value'], d['clocks'])
PY
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref_r01i.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 120 python tools/profile_step.py 3 8 ft > gpurun_out/plain.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_ft_r01i.csv python tools/profile_step.py 3 8 ft > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches ft rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_inf_r01i.csv python tools/profile_step.py 3 16 inf > gpurun_out/ncu_launches2.log 2>&1
echo "ncu launches inf rc=$?"
timeout 500 ncu --set full --clock-control none -k regex:"conv3x3_tc_kernel" -s 34 -c 17 -o gpurun_out/prof_conv_r01i -f python tools/profile_step.py 3 16 inf > gpurun_out/ncu_full1.log 2>&1
echo "ncu full conv rc=$?"; ls -la gpurun_out/*.ncu-rep
