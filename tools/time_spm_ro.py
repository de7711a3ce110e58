#!/usr/bin/env python
"""Time the read-only (validation) form of pose_spm_fused at config 4 -- CUDA-graph replay, CUDA events."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import pose_b200 as pb  # noqa: E402
from _inputs import spm_inputs  # noqa: E402
from extra_workloads import graph_time  # noqa: E402

dev = torch.device("cuda", 0)
for n in (256, 1024):
    c, j, cnt, _t, x = spm_inputs(n, dev)
    ms = graph_time(lambda: pb.spm_fused(x, c, j, cnt, 1, want_grad=False), 10)
    msg = graph_time(lambda: pb.spm_fused(x, c, j, cnt, 1), 10)
    print(f"N={n}: read-only {ms * 1e3:.1f} us ({n * 35 * 65536 / ms / 1e6 / 6550.7 * 100:.1f} %), loss+grad {msg * 1e3:.1f} us", flush=True)
