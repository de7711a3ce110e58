#!/usr/bin/env python
"""Time pose_spm_decode at config 4 (N images of 35 x 128 x 128, thr 0.5) -- CUDA-graph replay, CUDA events."""
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
for n in (1024, 256):
    c, j, cnt, _t, x = spm_inputs(n, dev)
    ms = graph_time(lambda: pb.spm_decode_batch(x, 512, 1, 0.5, True, 32), 20)
    print(f"N={n}: {ms * 1e3:.2f} us per call", flush=True)
