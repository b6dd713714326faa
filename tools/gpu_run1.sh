#!/bin/bash
# first GPU pass: kernels without tcgen05, then tcgen05, then network tests; each in its own process
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "not tc" --timeout 300 -p no:cacheprovider > gpurun_out/t1_kernels_simt.log 2>&1; echo "simt kernels rc=$?"
tail -5 gpurun_out/t1_kernels_simt.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "tc" --timeout 120 -p no:cacheprovider > gpurun_out/t2_kernels_tc.log 2>&1; echo "tc kernels rc=$?"
tail -15 gpurun_out/t2_kernels_tc.log
timeout 900 python -m pytest tests/test_gpu_network.py -q -m gpu -k "not bf16 or bf16_simt" --timeout 300 -p no:cacheprovider > gpurun_out/t3_net_fp32.log 2>&1; echo "net fp32 rc=$?"
tail -15 gpurun_out/t3_net_fp32.log
timeout 900 python -m pytest tests/test_gpu_network.py -q -m gpu -k "bf16 and not bf16_simt" --timeout 300 -p no:cacheprovider > gpurun_out/t4_net_bf16.log 2>&1; echo "net bf16 rc=$?"
tail -15 gpurun_out/t4_net_bf16.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/t5_smoke.log 2>&1; echo "smoke rc=$?"
tail -5 gpurun_out/t5_smoke.log
