#!/bin/bash
# round-2 call 6: split-operand (fp32 on tensor cores) kernels + network tests, then the full suite and bench
cd "$(dirname "$0")/.."
TAG=${1:-r02f}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "split" > gpurun_out/t_split.log 2>&1; echo "split kernel tests rc=$?"; tail -8 gpurun_out/t_split.log | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_network.py -q -x -k "split_operand or fp32_tc" > gpurun_out/t_split_net.log 2>&1; echo "split network tests rc=$?"; tail -8 gpurun_out/t_split_net.log | cut -c1-300
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; echo "full gpu suite rc=$?"; tail -6 gpurun_out/t_all.log | cut -c1-300
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_$TAG.err
