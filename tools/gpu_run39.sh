#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests -q -m gpu --timeout 150 -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^E  +(Assert|assert|Runtime|Type|Attr|Value|Key|Index|Name)|^FAILED" gpurun_out/t_all.log | cut -c1-300 | head -20
for fw in 1; do
timeout 600 python bench.py --iters 200 --steps 1 --warmup 2 --fuse-window $fw > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err; echo "fuse_window=$fw rc=$?"; tail -2 gpurun_out/bench_ab.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_ab.json'))
print('   ', d['value'], d['finetune_s_per_sequence'], d['inference_fps'], d['finetune_tflops'])
PY
done
