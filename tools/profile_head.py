#!/usr/bin/env python
"""Short driver for ncu: the head-fusion kernel (training form: grad + decode) at B images of 512 x 64 x 48 features."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pose_b200 as pb  # noqa: E402
from tools.head_check import make_inputs  # noqa: E402

b = int(sys.argv[1]) if len(sys.argv) > 1 else 296
tuning = int(sys.argv[2]) if len(sys.argv) > 2 else 0
feats, weight, kp = make_inputs(b, 512, 17, 64, 48)
kpt = torch.from_numpy(kp).cuda()
for _ in range(3):
    pb.sbp_head_fused(feats, weight, kpt, sigma=2, want_grad=True, decode=True, conf_threshold=0.25, coord_scale=4.0, tuning=tuning)
torch.cuda.synchronize()
print("ok")
