#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/conv_probe.py 1 > gpurun_out/probe_b1.log 2>&1; echo "probe1 rc=$?"
timeout 300 python tools/wgrad_probe.py 1 > gpurun_out/wgrad_probe.log 2>&1; echo "wgrad rc=$?"
cat gpurun_out/probe_b1.log | cut -c1-250
tail -30 gpurun_out/wgrad_probe.log
