#!/bin/bash
# 2 GPUs: multi-GPU tests, bench at N=2 (includes the exchange parity + stress check), then N=1 bench for reference
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -x > gpurun_out/pytest_multigpu.log 2>&1; echo "pytest multigpu rc=$?"
tail -5 gpurun_out/pytest_multigpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"
cat gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench n1 rc=$?"
cat gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
