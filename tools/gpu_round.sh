#!/bin/bash
# One gpurun call: GPU tests, the pipelined-loop sweep, compute-sanitizer passes, the CPU reference arm.
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python tools/tune_fused.py --run --which pipe --reps 40 > gpurun_out/tune_pipe.log 2>&1; echo "tune rc=$?"
cat gpurun_out/tune_pipe.log
PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout 420 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_integration_gpu.py tests/test_spm_gpu.py tests/test_oks_gpu.py -q -x \
  -k "randomized or nms_rules or crowded or many_persons or score_ties or ground_truth_as" > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"
tail -5 gpurun_out/memcheck.log
timeout 420 $CS --tool racecheck --error-exitcode 9 python -m pytest tests/test_integration_gpu.py tests/test_spm_gpu.py -q -x \
  -k "randomized_shape_sweep_against_oracle and (0 or 1 or 2) or nms_rules or crowded" > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?"
tail -5 gpurun_out/racecheck.log
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
cat gpurun_out/bench_ref.json
