#!/usr/bin/env python
"""Times spm_fused (grad / read-only) and spm_render with the usual persons and with none (the bare streaming skeleton)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import pose_b200 as pb
from _inputs import spm_inputs
from kbench import timeit
dev = torch.device("cuda", 0)
for n in (256, 1024):
    c, j, cnt, _t, x = spm_inputs(n, dev)
    zero = torch.zeros_like(cnt)
    for tag, cc in (("persons", cnt), ("empty", zero)):
        for g in (True, False):
            med, best = timeit(lambda: pb.spm_fused(x, c, j, cc, 1, want_grad=g), 30)
            gb = (2 if g else 1) * 35 * 128 * 128 * 4 * n / (med * 1e-3) / 1e9
            print(f"N={n:5d} {tag:8s} grad={int(g)}: {med*1e3:8.1f} us  {gb:7.1f} GB/s", flush=True)
        med, best = timeit(lambda: pb.spm_render_batch(c, j, cc, 128, 1), 30)
        print(f"N={n:5d} {tag:8s} render:  {med*1e3:8.1f} us  {35 * 128 * 128 * 4 * n / (med * 1e-3) / 1e9:7.1f} GB/s", flush=True)
