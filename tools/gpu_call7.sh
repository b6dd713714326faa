#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "split or fold_and_repack" -s > gpurun_out/t_split.log 2>&1; echo "split kernel tests rc=$?"; grep -E "split conv|passed|failed|^E  " gpurun_out/t_split.log | cut -c1-200 | head -40
timeout 600 python -m pytest tests/test_gpu_network.py -q -k "split_operand or fp32_tc" > gpurun_out/t_split_net.log 2>&1; echo "split network tests rc=$?"; grep -E "passed|failed|^E  " gpurun_out/t_split_net.log | cut -c1-250 | head -20
timeout 300 python - <<'PY' 2>&1 | tail -8
import json, sys, torch
sys.path.insert(0, '.')
import bench
class A: pass
dev = torch.device('cuda:0')
sd0 = bench.calibrated_state_dict('parent')
fr, ms = bench.make_sequence_gpu(0, 8)
print(json.dumps(bench.fp32_modes_leg(sd0, fr.to(dev), fr, dev), indent=1))
PY
