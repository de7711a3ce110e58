#!/bin/bash
# One gpurun call: full GPU tests, smoke(), per-kernel timings, both bench arms.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 400 python tools/kbench.py --out gpurun_out/kbench.json > gpurun_out/kbench.log 2>&1; echo "kbench rc=$?"
tail -60 gpurun_out/kbench.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
cat gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
nproc; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
