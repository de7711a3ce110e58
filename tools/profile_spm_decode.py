#!/usr/bin/env python
"""Short driver for ncu: SPM decode at config 4 shapes (N images, default 1024)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pose_b200 as pb  # noqa: E402
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _inputs import spm_inputs  # noqa: E402

dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
c, j, cnt, _t, x = spm_inputs(n, dev)
for _ in range(4):
    pb.spm_decode_batch(x, 512, 1, 0.5, True, 32)
torch.cuda.synchronize()
print("ok")
