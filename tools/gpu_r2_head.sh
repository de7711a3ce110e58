#!/bin/bash
# head-fusion round: full GPU tests, bench with extras, stage sweep, ncu capture of the head kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_n1.json
for t in 5 6 7 8; do timeout 120 python tools/head_check.py --acc --time 1024 --iters 5 --tuning $t --out gpurun_out/head_sweep.jsonl 2>&1 | tail -1 | cut -c1-330; done
timeout 200 python tools/head_check.py --acc --time 4096 --iters 3 --tuning 0 --out gpurun_out/head_b4096.jsonl 2>&1 | tail -1 | cut -c1-700
bash tools/gpu_ncu_head.sh
