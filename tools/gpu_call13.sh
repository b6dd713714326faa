#!/bin/bash
mkdir -p gpurun_out
{
echo "--- default"; timeout 200 python tools/wgrad_probe.py 10 2>&1
echo "--- one wave"; FOSVOS_WG_ONE_WAVE=1 timeout 200 python tools/wgrad_probe.py 10 2>&1
} > gpurun_out/wg_waves.log 2>&1
cat gpurun_out/wg_waves.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/tests_r02l.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/tests_r02l.log
tail -5 gpurun_out/tests_r02l.log
timeout 900 python bench.py > gpurun_out/bench_r02l.json 2> gpurun_out/bench_r02l.err; echo "bench rc=$?"
cat gpurun_out/bench_r02l.json | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'e2e') if k in d})
for k in ('roofline', 'roofline_wgrad', 'roofline_dgrad', 'roofline_step', 'clocks', 'parity'):
    print(k, d.get(k))
print({k: v for k, v in d.get('config', {}).items() if 'finetune' in k or 'inference' in k})
"
