#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 tools/offline_dp_check.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -5
echo "dp rc=${PIPESTATUS[0]}"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --steps 1 --warmup 3 --iters 200 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench2 rc=$?"; tail -2 gpurun_out/bench_2gpu.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_2gpu.json'):
    if l.startswith('{'):
        d=json.loads(l); print({k:d[k] for k in ('value','n_gpus','ms_per_step','inference_fps','finetune_s_per_sequence')}, d['e2e'])
PY
