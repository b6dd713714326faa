#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for i in 1 2 3; do timeout 200 python -m pytest tests/test_gpu_network.py -q -x -s -k "step_graph_follows" 2>&1 | grep -E "eager-vs-graph|passed|failed|Error" | cut -c1-200; done
timeout 200 python -m pytest tests/test_gpu_kernels.py -q -x -k "fused_conv_step" 2>&1 | tail -2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_r02i_n2.json 2> gpurun_out/bench_r02i_n2.err; echo "bench N=2 rc=$?"; tail -5 gpurun_out/bench_r02i_n2.err | cut -c1-400
