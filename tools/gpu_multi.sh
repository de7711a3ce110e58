#!/bin/bash
# N GPUs of one box (default 8): bench with the exchange parity + stress check, both arms
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N rc=$?"
tail -c 2500 gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err
