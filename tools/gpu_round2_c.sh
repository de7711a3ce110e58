#!/bin/bash
# One gpurun call: GPU tests, SBP knob sweep, decode timings for tree vs keep1, ncu --set full of the headline kernel.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sbp_gpu.py tests/test_integration_gpu.py -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python tools/tune_fused.py --run > gpurun_out/tune_fused.log 2>&1; echo "tune rc=$?"
cat gpurun_out/tune_fused.log
echo "== decode, tree"; timeout 200 python tools/kbench.py --only decode_pred --no-spm --out gpurun_out/kb_dec_tree.json 2>&1 | grep decode
echo "== decode, keep1"; POSE_B200_LIB=build/tune/sbp_keep1.so timeout 200 python tools/kbench.py --only decode_pred --no-spm --out gpurun_out/kb_dec_keep1.json 2>&1 | grep decode
timeout 120 python tools/profile_fused.py fused > gpurun_out/plain_fused.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sbp_fused -s 2 -c 1 -f -o gpurun_out/ncu_fused_r02 python tools/profile_fused.py fused > gpurun_out/ncu_fused.log 2>&1
echo "ncu rc=$?"
