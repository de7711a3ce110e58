#!/usr/bin/env python
"""Single-GPU driver for profiling the exchange epilogue: a fake 2-rank descriptor whose 'peers' are plain local buffers."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pose_b200 as pb  # noqa: E402
from pose_b200._cabi import ExchangeDesc, lib  # noqa: E402

dev = torch.device("cuda", 0)
B, K, H, W = 4096, 17, 64, 48
gen = torch.Generator(device=dev).manual_seed(0)
logits = torch.randn(B, K, H, W, device=dev, generator=gen) * 3
kp = torch.rand(B, K, 2, device=dev, generator=gen, dtype=torch.float64) * 48
bbox = torch.rand(B, 4, device=dev, generator=gen, dtype=torch.float64) * 300 + 40
d = ExchangeDesc()
d.world, d.rank, d.batch_local, d.num_keypoints = 2, 0, B, K
nbytes = int(lib().pose_exchange_layout(ctypes.byref(d)))
bufs = [torch.zeros(nbytes, dtype=torch.uint8, device=dev) for _ in range(2)]
for r in range(2):
    d.peer_base[r] = bufs[r].data_ptr()
ids = torch.stack([torch.arange(B, device=dev), torch.ones(B, dtype=torch.int64, device=dev)], 1).contiguous()
d.ids_local = ids.data_ptr()


class Ex:
    desc = d


outs = dict(dlogits=torch.empty_like(logits), joints=torch.empty(B, K, 3, device=dev), loss=torch.empty((), device=dev))
for _ in range(4):
    pb.sbp_fused(logits, keypoints=kp, sigma=2, want_grad=True, decode=True, conf_threshold=0.25, coord_scale=4.0, out=outs,
                 bbox=bbox, input_size=(256, 192), exchange=Ex)
    pb.sbp_fused(logits, keypoints=kp, sigma=2, want_grad=True, decode=True, conf_threshold=0.25, coord_scale=4.0, out=outs,
                 bbox=bbox, input_size=(256, 192))
torch.cuda.synchronize()
print("ok")
