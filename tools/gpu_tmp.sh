#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests -q -m gpu --timeout 150 -p no:cacheprovider -x > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^E  +(Assert|assert|Runtime)|^FAILED" gpurun_out/t_all.log | cut -c1-300 | head
