#!/bin/bash
# One gpurun call: full GPU tests, per-kernel timings, ncu --set full of the read-only fused variants, bench.  (compute-sanitizer is closed on this GPU pool.)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python tools/kbench.py --out gpurun_out/kbench.json > gpurun_out/kbench.log 2>&1; echo "kbench rc=$?"
cat gpurun_out/kbench.log
for w in val val_loss; do
  timeout 120 python tools/profile_fused.py $w > /dev/null 2>&1 && \
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:sbp_fused -s 2 -c 1 -f -o gpurun_out/ncu_$w python tools/profile_fused.py $w > gpurun_out/ncu_$w.log 2>&1
  echo "ncu $w rc=$?"
done
timeout 300 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
cat gpurun_out/bench_n1.json
