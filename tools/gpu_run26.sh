#!/bin/bash
mkdir -p gpurun_out
timeout 420 python -m pytest tests -q -m gpu --timeout 120 -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^E  +(Assert|assert|Runtime|Type|Attr|Value|Key|Index|Name)|^FAILED" gpurun_out/t_all.log | cut -c1-300 | head -30
for ov in 1; do
FOSVOS_OVERLAP=$ov timeout 600 python bench.py --iters 200 --steps 1 --warmup 2 > gpurun_out/bench_ov$ov.json 2> gpurun_out/bench_ov$ov.err; echo "bench overlap=$ov rc=$?"; tail -3 gpurun_out/bench_ov$ov.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_ov$ov.json'))
print({k:d[k] for k in ('value','ms_per_step','inference_fps','finetune_s_per_sequence','finetune_tflops')}, d['e2e']['value'], d['roofline']['frac'], d['roofline_side_chain']['frac'], d['roofline_loss']['frac'])
PY
done

