#!/bin/bash
# round-2 call 3: hook test, bench, ncu launch lists + full captures of the forward convs and the side kernel
cd "$(dirname "$0")/.."
TAG=${1:-r02c}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_network.py -x -q -k "hooks_fire or augmentation or step_graph" > gpurun_out/t_hooks.log 2>&1; echo "hook tests rc=$?"; tail -5 gpurun_out/t_hooks.log | cut -c1-300
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_$TAG.err
timeout 120 python tools/profile_step.py 3 8 ft > gpurun_out/plain.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_ft_$TAG.csv python tools/profile_step.py 3 8 ft > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches ft rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_inf_$TAG.csv python tools/profile_step.py 3 16 inf > gpurun_out/ncu_launches2.log 2>&1
echo "ncu launches inf rc=$?"
timeout 700 ncu --set full --clock-control none --import-source on -k regex:"conv3x3_tc_kernel|conv3x3_side_tc_kernel" -s 34 -c 17 -o gpurun_out/prof_conv_$TAG -f python tools/profile_step.py 3 16 inf > gpurun_out/ncu_full1.log 2>&1
echo "ncu full conv rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"side_upsample_sep2" -s 4 -c 1 -o gpurun_out/prof_side_$TAG -f python tools/side_sep_probe.py 16 > gpurun_out/ncu_side.log 2>&1
echo "ncu full side rc=$?"
timeout 100 python tools/side_sep_probe.py 1 5 16 2>&1 | tail -1
