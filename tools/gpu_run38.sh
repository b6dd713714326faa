#!/bin/bash
for b in 1 8; do
for v in "" "FOSVOS_TC_SKIP_B=1" "FOSVOS_TC_SKIP_A=1" "FOSVOS_TC_SKIP_A=1 FOSVOS_TC_SKIP_B=1"; do
echo "== batch $b $v"
env $v PROBE_BN=0 PROBE_MASK=0 timeout 200 python tools/conv_probe.py $b 2>&1 | grep -E "^(480x854 64->64|240x427 64->128|240x427 128->128|120x214 256->256|60x107 512->512|30x54 512->512):" | cut -c1-60
done
done
