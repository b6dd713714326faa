#!/bin/bash
timeout 500 python -m pytest tests/test_gpu_network.py -q -m gpu --timeout 200 -p no:cacheprovider -k "config2 or augmented" 2>&1 | grep -E "passed|failed|^E  +(Assert|assert|Runtime)|^FAILED" | cut -c1-300
