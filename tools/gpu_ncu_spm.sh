#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/profile_spm_fused.py 256 > gpurun_out/plain_spm.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:spm_unit -s 4 -c 2 -f -o gpurun_out/ncu_r02_spm_unit python tools/profile_spm_fused.py 256 > gpurun_out/ncu_spm.log 2>&1
echo "ncu rc=$?"
