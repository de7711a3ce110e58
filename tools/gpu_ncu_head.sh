#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/profile_head.py 296 > gpurun_out/plain_head.log 2>&1 && \
timeout 500 ncu --set full --clock-control none --import-source on -k regex:sbp_head_fused -s 1 -c 1 -f -o gpurun_out/ncu_r02_head python tools/profile_head.py 296 > gpurun_out/ncu_head.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_head.log
