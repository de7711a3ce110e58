#!/usr/bin/env python
"""Short driver for ncu: SPM render / loss / decode at config 4 shapes (N=128)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pose_b200 as pb  # noqa: E402
from oracle import cases  # noqa: E402  (input generator only)

dev = torch.device("cuda", 0)
n = 128
people, tgt, lg, meta = cases.spm_case("coco", 16, seed=99)
c, j, cnt = cases.pack_people(people)
rep = n // 16
c = torch.from_numpy(c).repeat(rep, 1, 1).to(dev)
j = torch.from_numpy(j).repeat(rep, 1, 1, 1).to(dev)
cnt = torch.from_numpy(cnt).repeat(rep).to(dev)
x = lg.repeat(rep, 1, 1, 1).to(dev)
for _ in range(3):
    t = pb.spm_render_batch(c, j, cnt, 128, 1)
    pb.spm_loss_fused(x, t)
    pb.spm_decode_batch(x, 512, 1, 0.5, True, 32)
torch.cuda.synchronize()
print("ok")
