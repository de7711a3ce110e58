#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_spm_gpu.py tests/test_integration_gpu.py -m gpu -q -x > gpurun_out/pytest_spm.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest_spm.log
timeout 300 python tools/kbench.py --only spm --out gpurun_out/kbench_spm.json 2>&1 | grep spm
timeout 600 python tools/tune_spm.py --run 2>&1 | tee gpurun_out/tune_spm.log
