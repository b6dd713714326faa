#!/bin/bash
mkdir -p gpurun_out
timeout 100 python tools/side_probe.py 8 2>&1 | head -3
for b in 8 16; do
timeout 600 python bench.py --iters 100 --steps 1 --warmup 2 --batch $b > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err; echo "batch=$b rc=$?"; tail -2 gpurun_out/bench_ab.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_ab.json'))
print('   ', d['finetune_s_per_sequence'], d['inference_fps'], d['roofline']['frac'], d['roofline_side_chain']['frac'], d['roofline_loss']['frac'])
PY
done
