"""Synthetic inputs for the measurement tools (config 4 of BASELINE.json: multi-person SPM maps).

Independent of oracle/ (which is test infrastructure): persons are drawn here, the target is rendered by the product's
own kernel and the logits are the inverse activations of (target + noise), computed with torch on the device."""
import numpy as np
import torch


def spm_inputs(n, dev, k=17, res=128, max_people=8, seed=99, noise=0.02):
    import pose_b200 as pb
    rng = np.random.default_rng(seed)
    counts = rng.integers(1, max_people + 1, size=n).astype(np.int32)
    centers = rng.integers(8, res - 8, size=(n, max_people, 2), dtype=np.int64)
    joints = np.clip(centers[:, :, None, :] + rng.integers(-30, 31, size=(n, max_people, k, 2), dtype=np.int64), 1, res - 2)
    joints[rng.uniform(size=(n, max_people, k)) < 0.1] = 0
    c = torch.from_numpy(centers).to(dev)
    j = torch.from_numpy(joints).to(dev)
    cnt = torch.from_numpy(counts).to(dev)
    target = pb.spm_render_batch(c, j, cnt, res, 1)
    gen = torch.Generator(device=dev).manual_seed(seed)
    u = torch.rand(target.shape, device=dev, generator=gen)
    root = (target[:, :1] * (0.90 + 0.09 * u[:, :1]) + noise * u[:, :1].flip(-1)).clamp(1e-4, 1 - 1e-4)
    disp = (target[:, 1:] + noise * (u[:, 1:] - 0.5)).clamp(-0.999, 0.999)
    logits = torch.cat([torch.logit(root), torch.atanh(disp)], dim=1).contiguous()
    return c, j, cnt, target, logits
