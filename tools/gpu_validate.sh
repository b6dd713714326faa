#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests -q -m gpu --timeout 150 -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^E  +(Assert|assert|Runtime)|^FAILED" gpurun_out/t_all.log | cut -c1-300 | head
timeout 200 python __graft_entry__.py smoke > gpurun_out/t_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/t_smoke.log
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_r01i.json 2> gpurun_out/bench_r01i.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r01i.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r01i.json'))
print({k:d[k] for k in ('value','ms_per_step','inference_fps','finetune_s_per_sequence','finetune_tflops','gpu_launches')}, d['e2e'], d['roofline']['frac'], d['roofline_side_chain']['frac'], d['roofline_loss']['frac'], d['cpu_baseline']['value'], d['clocks'])
PY
timeout 120 python tools/profile_step.py 3 8 ft > gpurun_out/plain.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_ft_r01i.csv python tools/profile_step.py 3 8 ft > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches ft rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_inf_r01i.csv python tools/profile_step.py 3 16 inf > gpurun_out/ncu_launches2.log 2>&1
echo "ncu launches inf rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"conv3x3_tc_kernel" -s 34 -c 17 -o gpurun_out/prof_conv_r01i -f python tools/profile_step.py 3 16 inf > gpurun_out/ncu_full1.log 2>&1
echo "ncu full conv rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"wgrad_tc_kernel|side_upsample|side_heads|bal_loss_fused|maxpool_bwd|side_bwd" -s 17 -c 24 -o gpurun_out/prof_misc_r01i -f python tools/profile_step.py 2 8 both > gpurun_out/ncu_full2.log 2>&1
echo "ncu full misc rc=$?"
