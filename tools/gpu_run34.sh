#!/bin/bash
timeout 300 python -m pytest tests -q -m gpu --timeout 100 -p no:cacheprovider -k "side or loss or golden or full_size" 2>&1 | grep -E "passed|failed|^E  +(Assert|assert)|^FAILED" | cut -c1-200 | head
timeout 100 python tools/side_probe.py 8 2>&1 | tail -9
