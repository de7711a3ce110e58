#!/bin/bash
# ncu --set full of the bulk-staged fused kernels: train (grad+decode) and validation loss (read-only)
mkdir -p gpurun_out
for w in fused val_loss; do
  timeout 120 python tools/profile_fused.py $w > gpurun_out/plain_$w.log 2>&1 && \
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:sbp_fused -s 2 -c 1 -f -o gpurun_out/ncu_r02_$w python tools/profile_fused.py $w > gpurun_out/ncu_$w.log 2>&1
  echo "ncu $w rc=$?"
done
