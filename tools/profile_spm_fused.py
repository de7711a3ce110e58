#!/usr/bin/env python
"""Short driver for ncu: the fused SPM render+loss+grad kernel at config 4 shapes (N=128 by default)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pose_b200 as pb  # noqa: E402
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _inputs import spm_inputs  # noqa: E402

dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
c, j, cnt, _t, x = spm_inputs(n, dev)
for _ in range(3):
    pb.spm_fused(x, c, j, cnt, 1)
    pb.spm_fused(x, c, j, cnt, 1, want_grad=False)
torch.cuda.synchronize()
print("ok")
