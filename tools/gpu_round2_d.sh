#!/bin/bash
# One gpurun call: GPU tests, per-kernel timings, the map-kernel knob sweep.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log
timeout 400 python tools/kbench.py --no-spm --out gpurun_out/kbench.json > gpurun_out/kbench.log 2>&1; echo "kbench rc=$?"
cat gpurun_out/kbench.log | grep -v "^$" | tail -25
echo "== register-staged"
POSE_B200_TMA=0 timeout 400 python tools/kbench.py --no-spm --no-tma --only fused --out gpurun_out/kbench_notma.json 2>&1 | grep -v "^$" | tail -8
timeout 600 python tools/tune_fused.py --run > gpurun_out/tune_fused.log 2>&1; echo "tune rc=$?"
cat gpurun_out/tune_fused.log
