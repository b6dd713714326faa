#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider -x > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^E  " gpurun_out/t_all.log | head -20
timeout 600 python tools/conv_probe.py 1 > gpurun_out/probe_b1.log 2>&1; echo "probe1 rc=$?"
timeout 600 python tools/conv_probe.py 8 > gpurun_out/probe_b8.log 2>&1; echo "probe8 rc=$?"
cut -c1-200 gpurun_out/probe_b1.log; cut -c1-200 gpurun_out/probe_b8.log
