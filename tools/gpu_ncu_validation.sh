#!/bin/bash
# ncu --set full of the two-phase validation kernels, then bench.py (both arms)
mkdir -p gpurun_out
for w in val val_loss; do
  timeout 120 python tools/profile_fused.py $w > /dev/null 2>&1 && \
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:sbp_fused -s 2 -c 1 -f -o gpurun_out/ncu_r02_$w python tools/profile_fused.py $w > gpurun_out/ncu_$w.log 2>&1
  echo "ncu $w rc=$?"
done
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err; echo "bench ref rc=$?"; tail -c 600 gpurun_out/bench_ref_n1.json
