#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/t_all.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/t_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/t_smoke.log
python tools/profile_step.py 2 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01.csv python tools/profile_step.py 2 > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 17 -c 17 -o gpurun_out/prof_conv_tc_r01 python tools/profile_step.py 1 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"side_|bal_loss|sgd_kernel|maxpool|wgrad" -c 24 -o gpurun_out/prof_misc_r01 python tools/profile_step.py 1 > gpurun_out/ncu_full2.log 2>&1
echo "ncu full2 rc=$?"
ls -la gpurun_out
