#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 4 --steps 1 --warmup 3 > gpurun_out/bench_4gpu.json 2> gpurun_out/bench_4gpu.err; echo "bench4 rc=$?"; tail -2 gpurun_out/bench_4gpu.err | cut -c1-200
python - <<'PY'
import json
for l in open('gpurun_out/bench_4gpu.json'):
    if l.startswith('{'):
        d=json.loads(l); print({k:d[k] for k in ('value','n_gpus','ms_per_step','inference_fps','finetune_s_per_sequence')}, d['e2e']['value'], d['clocks'])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29722 tools/offline_dp_check.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM\|Warning\|return func" | tail -3
