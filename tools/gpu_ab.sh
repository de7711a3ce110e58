#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sbp_gpu.py tests/test_integration_gpu.py -m gpu -x -q > gpurun_out/pytest_ab.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_ab.log
timeout 240 python tools/tune_fused.py --run --which ab --reps 40 > gpurun_out/tune_ab.log 2>&1; echo "tune rc=$?"
cat gpurun_out/tune_ab.log
timeout 240 python tools/tune_fused.py --run --which ng --reps 40 > gpurun_out/tune_ng.log 2>&1; echo "tune rc=$?"
cat gpurun_out/tune_ng.log
