#!/bin/bash
# kbench of earlier commits (built in worktrees under build/wt/) on one box: which change cost what
mkdir -p gpurun_out
for s in eef81ca 5a8b3c6 0389015 f38e417; do
  (cd build/wt/$s && timeout 300 python tools/kbench.py --out ../../../gpurun_out/kbench_$s.json > ../../../gpurun_out/kbench_$s.log 2>&1; echo "$s rc=$?")
  echo "=== $s"; grep -E "fused|render|spm_decode|decode_pred1_randn " gpurun_out/kbench_$s.log
done
