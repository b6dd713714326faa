#!/bin/bash
mkdir -p gpurun_out
for cfg in "pool 1 0" "pool 0 0" "epilogue 1 0" "epilogue 0 0" "epilogue 0 1" "pool 1 1"; do
set -- $cfg
if [ "$3" = "1" ]; then export FOSVOS_TC_MASK_REGS=1; else unset FOSVOS_TC_MASK_REGS; fi
FOSVOS_BWD_FANIN=$1 FOSVOS_HP=$2 timeout 600 python bench.py --iters 200 --steps 1 --warmup 2 > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err; echo "fanin=$1 hp=$2 maskregs=$3 rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_ab.json'))
print('   ', d['finetune_s_per_sequence'], d['inference_fps'])
PY
done
