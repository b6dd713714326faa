#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider -x > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/t_all.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/t_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/t_smoke.log
timeout 900 python bench.py --iters 20 --steps 1 --warmup 1 --graph 0 > gpurun_out/bench_i20_nograph.json 2> gpurun_out/bench_i20_nograph.err; echo "bench nograph rc=$?"; tail -3 gpurun_out/bench_i20_nograph.err; cat gpurun_out/bench_i20_nograph.json
timeout 900 python bench.py --iters 20 --steps 1 --warmup 1 --graph 1 > gpurun_out/bench_i20_graph.json 2> gpurun_out/bench_i20_graph.err; echo "bench graph rc=$?"; tail -3 gpurun_out/bench_i20_graph.err; cat gpurun_out/bench_i20_graph.json
