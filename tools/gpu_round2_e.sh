#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 400 python tools/kbench.py --no-spm --out gpurun_out/kbench.json > gpurun_out/kbench.log 2>&1; echo "kbench rc=$?"
cat gpurun_out/kbench.log | grep -v "^$" | tail -25
timeout 600 python tools/tune_fused.py --run > gpurun_out/tune_fused.log 2>&1; echo "tune rc=$?"
cat gpurun_out/tune_fused.log
for w in fused val; do
  timeout 120 python tools/profile_fused.py $w > gpurun_out/plain_$w.log 2>&1 && \
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:sbp_fused -s 2 -c 1 -f -o gpurun_out/ncu_r02_$w python tools/profile_fused.py $w > gpurun_out/ncu_$w.log 2>&1
  echo "ncu $w rc=$?"
done
