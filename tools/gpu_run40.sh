#!/bin/bash
PROBE_SPLITS=1 timeout 300 python tools/wgrad_probe.py 5 2>&1 | cut -c1-100
PROBE_BN=1 PROBE_MASK=0 timeout 300 python tools/conv_probe.py 5 2>&1 | cut -c1-170
