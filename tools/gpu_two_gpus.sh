#!/bin/bash
# final build on two GPUs: bench under torchrun (value, config4, config5 with the exposed all-reduce), both DP modes checked by the offline tool
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_r02q_n2.json 2> gpurun_out/bench_r02q_n2.err; echo "bench N=2 rc=$?"; tail -3 gpurun_out/bench_r02q_n2.err | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/offline_dp_check.py > gpurun_out/dp_check_r02q.log 2>&1; echo "dp check rc=$?"; tail -6 gpurun_out/dp_check_r02q.log | cut -c1-300
