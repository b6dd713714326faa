#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "wgrad_tc" --timeout 120 -p no:cacheprovider > gpurun_out/t_wgrad_tc.log 2>&1; echo "wgrad_tc rc=$?"; grep -E "passed|failed|^E  " gpurun_out/t_wgrad_tc.log | head -30
python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/t_all.log
timeout 900 python bench.py --iters 100 --steps 1 --warmup 1 --graph 0 > gpurun_out/bench_i100.json 2> gpurun_out/bench_i100.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_i100.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_i100.json'))
print({k:d[k] for k in ('value','inference_fps','finetune_s_per_sequence','finetune_tflops','gpu_launches')}, d['e2e'], d['roofline']['frac'])
PY
