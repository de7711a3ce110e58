#!/bin/bash
# round-2 state check: GPU tests, smoke, bench (both arms), per-kernel timings
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_n1.json
timeout 600 python tools/kbench.py --out gpurun_out/kbench.json > gpurun_out/kbench.log 2>&1; echo "kbench rc=$?"
cat gpurun_out/kbench.log | grep -v "^$" | tail -40
