#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^E  +(Assert|assert|Runtime|Type|Attr|Value|Key|Index|Name)|^FAILED" gpurun_out/t_all.log | cut -c1-300 | head -40
