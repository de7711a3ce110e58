#!/usr/bin/env python
"""Short driver for `ncu --set full`: a few launches of the headline fused kernel (and the decode kernel) at B=4096."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pose_b200 as pb  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "fused"
dev = torch.device("cuda", 0)
B, K, H, W = 4096, 17, 64, 48
gen = torch.Generator(device=dev).manual_seed(0)
logits = torch.randn(B, K, H, W, device=dev, generator=gen) * 3
kp = torch.stack([torch.rand(B, K, device=dev, generator=gen, dtype=torch.float64) * W,
                  torch.rand(B, K, device=dev, generator=gen, dtype=torch.float64) * H], -1)
kp[torch.rand(B, K, device=dev, generator=gen) >= 0.85] = -1
for _ in range(4):
    if which == "fused":
        r = pb.sbp_fused(logits, keypoints=kp, sigma=2, want_grad=True, decode=True, conf_threshold=0.25, coord_scale=4.0)
    elif which == "val":          # validation step: loss + decode, no dlogits (read-only variant)
        r = pb.sbp_fused(logits, keypoints=kp, sigma=2, want_grad=False, decode=True, conf_threshold=0.25, coord_scale=4.0)
    elif which == "val_loss":     # validation loss only (read-only variant, no decode)
        r = pb.sbp_fused(logits, keypoints=kp, sigma=2, want_grad=False)
    elif which == "decode":
        r = pb.decode_batch(logits, 0.25, 4.0, True)
    elif which == "dense":
        t = pb.SBPHeatmapGenerator([H, W], K, 2).render_batch(kp)
        r = pb.sbp_fused(logits, target=t)
torch.cuda.synchronize()
print("ok", pb.launch_count())
