#!/bin/bash
timeout 200 python -m pytest tests -q -m gpu --timeout 100 -p no:cacheprovider -k "loss or finetune" 2>&1 | tail -1
timeout 100 python tools/side_probe.py 8 2>&1 | tail -2
timeout 100 python tools/side_probe.py 1 2>&1 | tail -2
python - <<'PY'
import sys, torch
sys.path.insert(0,'.')
from fosvos_b200 import ops
sys.argv=['x','8']
import importlib.util
spec=importlib.util.spec_from_file_location('sp','tools/side_probe.py')
PY
