#!/bin/bash
mkdir -p gpurun_out
export PROBE_LAYERS=conv3_2 PROBE_NO_BIAS=1
run() { echo "--- $1"; env $1 timeout 120 python tools/wgrad_probe.py 10 2>&1 | grep -v "^sum"; }
{
run "FOSVOS_WG_DEBUG=5"
run "FOSVOS_WG_DEBUG=6"
run "FOSVOS_WG_DEBUG=6 FOSVOS_WG_SPLITS=6"
run "FOSVOS_WG_DEBUG=6 FOSVOS_WG_SPLITS=3"
run "FOSVOS_WG_DEBUG=6 FOSVOS_WG_SPLITS=24"
run "FOSVOS_WG_DEBUG=6 FOSVOS_WG_SPLITS=1000"
run "FOSVOS_WG_DEBUG=1"
run "FOSVOS_WG_DEBUG=0"
run "FOSVOS_WG_NO_CTA_PAIR=1 FOSVOS_WG_DEBUG=6"
run "FOSVOS_WG_NO_CTA_PAIR=1 FOSVOS_WG_DEBUG=6 FOSVOS_WG_SPLITS=12"
run "FOSVOS_WG_NO_CTA_PAIR=1 FOSVOS_WG_DEBUG=6 FOSVOS_WG_SPLITS=1000"
} > gpurun_out/pair_exp2.log 2>&1
cat gpurun_out/pair_exp2.log
