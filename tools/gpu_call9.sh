#!/bin/bash
cd "$(dirname "$0")/.."
TAG=${1:-r02h}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/t_all.log 2>&1; echo "full gpu suite rc=$?"; tail -5 gpurun_out/t_all.log | cut -c1-300
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_$TAG.err
echo "--- A/B: separate fold + SGD + repack"; FOSVOS_FUSED_STEP=0 timeout 300 python bench.py --steps 2 --warmup 2 --parity 0 --gpu-reference 0 --config3 0 --config4 0 --fp32-modes 0 > gpurun_out/bench_${TAG}_nofusedstep.json 2> gpurun_out/bench_${TAG}_nofusedstep.err; echo "bench nofusedstep rc=$?"
timeout 120 python tools/profile_step.py 3 8 ft > gpurun_out/plain.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_ft_$TAG.csv python tools/profile_step.py 3 8 ft > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches ft rc=$?"
