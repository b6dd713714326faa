#!/bin/bash
# round-2 call 5: K16 data-gradient mode, multi-buffered first-layer stores, faster up-sampling kernel: tests + probes + bench
cd "$(dirname "$0")/.."
TAG=${1:-r02e}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/t_all.log 2>&1; echo "full gpu suite rc=$?"; tail -6 gpurun_out/t_all.log | cut -c1-300
echo -n "side probe (fast): "; timeout 60 python tools/side_sep_probe.py 1 5 16 2>&1 | tail -1
echo -n "side probe (exact exp): "; FOSVOS_SIDE_SEP2_EXACT_EXP=1 timeout 60 python tools/side_sep_probe.py 16 2>&1 | tail -1
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_$TAG.err
echo "--- A/B: K16 off"; FOSVOS_TC_NO_K16=1 timeout 300 python bench.py --steps 2 --warmup 2 --parity 0 --gpu-reference 0 --config3 0 --config4 0 > gpurun_out/bench_${TAG}_nok16.json 2> gpurun_out/bench_${TAG}_nok16.err; echo "bench nok16 rc=$?"
