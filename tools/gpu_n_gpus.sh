#!/bin/bash
# bench.py under torchrun on N GPUs of one box (value, config4, config5).  usage: gpu_n_gpus.sh N TAG   (run through gpurun --gpus N)
cd "$(dirname "$0")/.."
N=${1:-8}; TAG=${2:-r02r}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err; echo "bench N=$N rc=$?"; tail -3 gpurun_out/bench_${TAG}_n$N.err | cut -c1-300
tail -c 600 gpurun_out/bench_${TAG}_n$N.json
