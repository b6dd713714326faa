#!/bin/bash
for b in 5 16; do PROBE_BN=0 PROBE_MASK=0 timeout 200 python tools/conv_probe.py $b 2>&1 | grep -E "^(60x107 512->512|30x54 512->512|30x54 512->16):" | cut -c1-100; done
timeout 500 python -m pytest tests -q -m gpu --timeout 150 -p no:cacheprovider 2>&1 | grep -E "passed|failed|^E  +(Assert|assert)|^FAILED" | cut -c1-200 | head -12
timeout 600 python bench.py --iters 200 --steps 1 --warmup 2 > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_ab.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_ab.json'))
print('   ', d['value'], d['finetune_s_per_sequence'], d['inference_fps'], d['finetune_tflops'], d['roofline']['frac'], d['roofline_side_chain']['frac'], d['roofline_loss']['frac'])
PY
