#!/bin/bash
# round-2 validation call: microbenchmark, new-kernel tests, full GPU suite, smoke, bench, yardstick
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 tools/exp/mma_side > gpurun_out/mma_side.log 2>&1; echo "mma_side rc=$?"
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "side_tc or pool_only or fused_heads" > gpurun_out/t_side.log 2>&1; SIDE_RC=$?
echo "side kernel tests rc=$SIDE_RC"; tail -5 gpurun_out/t_side.log
if [ $SIDE_RC -ne 0 ]; then export FOSVOS_SIDE_TC=0; echo "FALLING BACK to FOSVOS_SIDE_TC=0 for the rest of this call"; fi
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; echo "full gpu suite rc=$?"; tail -15 gpurun_out/t_all.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke.log
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r02b.err
timeout 600 python tools/bf16_yardstick.py > gpurun_out/yardstick.log 2>&1; echo "yardstick rc=$?"
cat gpurun_out/mma_side.log
