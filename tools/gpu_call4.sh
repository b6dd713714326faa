#!/bin/bash
# round-2 call 4 (2 GPUs): multi-GPU bench legs under torchrun, weights A/B for the power-cap question, upsample rows sweep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_network.py -x -q -k "hooks_fire" > gpurun_out/t_hooks.log 2>&1; echo "hook test rc=$?"; tail -3 gpurun_out/t_hooks.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_r02d_n2.json 2> gpurun_out/bench_r02d_n2.err; echo "bench N=2 rc=$?"; tail -5 gpurun_out/bench_r02d_n2.err | cut -c1-400
for w in structured parent; do
  timeout 300 python bench.py --steps 2 --warmup 2 --parity 0 --gpu-reference 0 --config3 0 --config4 0 --weights $w > gpurun_out/bench_r02d_$w.json 2> gpurun_out/bench_r02d_$w.err; echo "bench $w rc=$?"
done
for r in 9 10 11 12 13 14 15 16 20 24 27; do echo -n "rows=$r: "; FOSVOS_SIDE_SEP2_ROWS=$r timeout 60 python tools/side_sep_probe.py 16 2>&1 | tail -1; done
