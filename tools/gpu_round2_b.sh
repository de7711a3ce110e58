#!/bin/bash
# One gpurun call: GPU tests, per-kernel timings, the SBP knob sweep.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 400 python tools/kbench.py --out gpurun_out/kbench.json > gpurun_out/kbench.log 2>&1; echo "kbench rc=$?"
grep -E "fused|render|spm_|decode_pred1_randn " gpurun_out/kbench.log
timeout 600 python tools/tune_fused.py --run > gpurun_out/tune_fused.log 2>&1; echo "tune rc=$?"
cat gpurun_out/tune_fused.log
