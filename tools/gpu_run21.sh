#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do
python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | grep -E "^E   +(Assertion|assert)|passed|failed|^FAILED" | cut -c1-250
done
