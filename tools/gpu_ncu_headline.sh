#!/bin/bash
# r02 evidence for the headline kernel: ncu --set full of the shipped sbp_fused_tma_kernel<GRAD,DECODE>, and the launch list of the bench command
mkdir -p gpurun_out
timeout 120 python tools/profile_fused.py fused > gpurun_out/plain_fused.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sbp_fused -s 2 -c 1 -f -o gpurun_out/ncu_r02_fused python tools/profile_fused.py fused > gpurun_out/ncu_fused.log 2>&1
echo "ncu fused rc=$?"
timeout 300 python bench.py --steps 2 --warmup 3 --no-graph --no-extras --no-aten-baseline > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err; echo "bench short rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches_ncu.csv python bench.py --steps 2 --warmup 3 --no-graph --no-extras --no-aten-baseline > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches rc=$?"; wc -l gpurun_out/r02_bench_launches_ncu.csv
