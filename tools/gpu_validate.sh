#!/bin/bash
# Round validation on one B200, in parts (one gpurun call each: the merge-back limit of gpurun_out/ is 64 MiB and three
# --set full reports do not fit together).  usage: gpu_validate.sh TAG PART   (PART = a | b | c)
#   a: parity tests, smoke, bench (both arms), ncu launch lists of the fine-tune window and of inference
#   b: ncu --set full of the 17 forward conv launches (batch 16) and of the side-chain up-sampling kernel
#   c: ncu --set full of the backward kernels (weight gradients incl. the CTA-pair kernel, optimizer step, pool / side backward)
# Summarise the outputs into profiles/ with tools/ncu_summary.py / tools/conv_traffic.py.
cd "$(dirname "$0")/.."
TAG=${1:-r02m}
PART=${2:-a}
mkdir -p gpurun_out
if [ "$PART" = a ]; then
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/t_all_$TAG.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED" gpurun_out/t_all_$TAG.log | cut -c1-300 | head
timeout 200 python __graft_entry__.py smoke > gpurun_out/t_smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/t_smoke_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_$TAG.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"
timeout 120 python tools/profile_step.py 3 8 ft > gpurun_out/plain.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_ft_$TAG.csv python tools/profile_step.py 3 8 ft > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches ft rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_inf_$TAG.csv python tools/profile_step.py 3 16 inf > gpurun_out/ncu_launches2.log 2>&1
echo "ncu launches inf rc=$?"
elif [ "$PART" = b ]; then
timeout 120 python tools/profile_step.py 3 16 inf > gpurun_out/plain.log 2>&1 &&
timeout 700 ncu --set full --clock-control none -k regex:"conv3x3_tc_kernel|conv3x3_side_tc_kernel|conv3x3_stack_tc_kernel" -s 34 -c 17 -o gpurun_out/prof_conv_$TAG -f python tools/profile_step.py 3 16 inf > gpurun_out/ncu_full1.log 2>&1
echo "ncu full conv rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"side_upsample_sep2" -s 4 -c 1 -o gpurun_out/prof_side_$TAG -f python tools/side_sep_probe.py 16 > gpurun_out/ncu_side.log 2>&1
echo "ncu full side rc=$?"
timeout 100 python tools/side_sep_probe.py 1 5 16 2>&1 | tail -1
else
timeout 120 python tools/profile_step.py 2 8 ft > gpurun_out/plain.log 2>&1 &&
timeout 500 ncu --set full --clock-control none -k regex:"wgrad_tc|conv_step_kernel|side_bwd|maxpool_bwd" -s 24 -c 24 -o gpurun_out/prof_bwd_$TAG -f python tools/profile_step.py 2 8 ft > gpurun_out/ncu_full2.log 2>&1
echo "ncu full bwd rc=$?"
fi
ls -la gpurun_out | tail -20
