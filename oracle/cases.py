"""Seeded input cases shared by oracle/make_golden.py and the tests.

TEST INFRASTRUCTURE ONLY.  Everything here is deterministic; each golden file
stores a sha256 of the inputs it was generated from so a drift in input
regeneration is caught loudly rather than showing up as a parity failure.
"""
import hashlib

import numpy as np
import torch

from . import sbp_oracle as so
from . import spm_oracle as po


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        if isinstance(a, torch.Tensor):
            a = a.detach().cpu().numpy()
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode())
        h.update(str(a.shape).encode())
        h.update(a.tobytes())
    return h.hexdigest()


# --------------------------------------------------------------------------- SBP

SBP_SHAPES = {
    # name: (K, H, W, sigma, input_size[H_in, W_in])
    "coco": (17, 64, 48, 2, (256, 192)),          # configs/sbp_coco.yaml
    "pis": (11, 64, 48, 2, (256, 192)),           # configs/sbp_pis.yaml (num_keypoints 11)
    "hires": (17, 96, 72, -1, (384, 288)),        # sigma<0 -> H/64 = 1.5 (half-to-even patch corners)
    "sigma3": (17, 64, 48, 3, (256, 192)),
    "sigma1": (5, 32, 24, 1, (128, 96)),
}


def sbp_edge_keypoints(k, h, w):
    """[E,K,2] float64 rows that exercise every clipping / truncation rule of the renderer."""
    pts = [
        (0.0, 0.0), (w - 1.0, h - 1.0), (w - 0.1, h - 0.1), (0.0, h - 1.0), (w - 1.0, 0.0),
        (w + 5.0, 10.0), (10.0, h + 9.0), (w * 3.0, h * 3.0),       # beyond the map: clamps to the edge
        (-1.0, -1.0), (-0.5, 10.0), (10.0, -0.001), (-0.0, 5.0),    # negative -> skipped; -0.0 is not < 0
        (10.999999, 20.000001), (10.5, 20.5), (3.2, 3.9), (w / 2.0, h / 2.0),
        (6.0, 6.0), (7.0, 7.0), (8.0, 8.0), (w - 8.0, h - 8.0), (w - 7.0, h - 9.0),
        (0.999999999, 0.999999999), (9.99999999999, 29.99999999999),  # fp64-only truncation cases
        (1e-9, 1e-9), (11.0, 11.0), (12.0, 13.0),
    ]
    e = (len(pts) + k - 1) // k
    arr = np.full((e * k, 2), -1.0)
    arr[:len(pts)] = np.array(pts)
    return arr.reshape(e, k, 2)


def sbp_case(name, batch=6, seed=1234):
    """(kp [B,K,2] f64, logits [B,K,H,W] f32, bbox [B,4] f64, image_id, category_id, meta)."""
    k, h, w, sigma, in_size = SBP_SHAPES[name]
    edge = sbp_edge_keypoints(k, h, w)
    n = max(batch, edge.shape[0] + 2)
    kp, logits, bbox, iid, cid = so.make_config1_inputs(n, k, h, w, seed=seed, torch_seed=seed % 97)
    kp[:edge.shape[0]] = edge    # first rows: the clipping / truncation edge cases
    logits = logits * 3.0        # spread so sigmoid crosses 0.25 / 0.99 thresholds often
    return kp, logits, bbox, iid, cid, dict(k=k, h=h, w=w, sigma=sigma, input_size=in_size)


def sbp_adversarial_maps(k=17, h=64, w=48, seed=5):
    """[A,K,H,W] fp32 logits: exact duplicate maxima, below-threshold maps, saturated plateaus, border peaks."""
    rng = np.random.default_rng(seed)
    base = torch.from_numpy(rng.normal(-4.0, 0.5, size=(6, k, h, w)).astype(np.float32))
    # 0: exact duplicate maxima at several places; the first row-major one must win
    for j in range(k):
        v = np.float32(2.0 + 0.125 * j)
        ys = rng.integers(0, h, 4)
        xs = rng.integers(0, w, 4)
        for y, x in zip(ys, xs):
            base[0, j, y, x] = float(v)
    # 1: everything far below any positive threshold
    base[1] = -12.0
    # 2: saturated plateau: many logits >= 17 with different values -> sigmoid == 1.0 for all of them
    for j in range(k):
        ys = rng.integers(0, h, 6)
        xs = rng.integers(0, w, 6)
        for t, (y, x) in enumerate(zip(ys, xs)):
            base[2, j, y, x] = 17.5 + 3.0 * t
    # 3: single peak on every border / corner
    border = [(0, 0), (0, w - 1), (h - 1, 0), (h - 1, w - 1), (0, w // 2), (h - 1, w // 2), (h // 2, 0), (h // 2, w - 1)]
    for j in range(k):
        y, x = border[j % len(border)]
        base[3, j, y, x] = 3.0
    # 4: the maximum is the very last element; 5: a constant map (all tied -> index 0)
    base[4, :, h - 1, w - 1] = 5.0
    base[5] = 1.25
    return base


NEARTIE_MAGNITUDES = (0.5, 3.0, 8.0, 12.0, 16.5, -2.0, 5.25, 10.0)


def _ulp_step(v, n):
    """v moved by n fp32 ulps (n may be negative)."""
    a = np.float32(v)
    for _ in range(abs(n)):
        a = np.nextafter(a, np.float32(np.inf if n > 0 else -np.inf))
    return a


def sbp_neartie_maps(n=8, k=17, h=64, w=48, seed=11):
    """[n,K,H,W] fp32 logits whose two (or three) largest values are 1-4 ulp apart at the magnitudes where fp32 sigmoid
    starts to merge neighbours: which of them is "the first maximum of sigmoid(x)" depends on the last bit of the sigmoid
    implementation.  Map (i, j): magnitude NEARTIE_MAGNITUDES[(i*K+j) % 8]; the earlier (row-major) pixel holds the
    value that is d = 1 + (i*K+j) % 4 ulp SMALLER on even maps and LARGER on odd maps; every third map has a third
    contender 2d ulp below, placed before both."""
    rng = np.random.default_rng(seed)
    x = rng.normal(-6.0, 0.5, size=(n, k, h, w)).astype(np.float32)
    for i in range(n):
        for j in range(k):
            t = i * k + j
            m = NEARTIE_MAGNITUDES[t % len(NEARTIE_MAGNITUDES)]
            d = 1 + t % 4
            p = np.sort(rng.choice(h * w, size=3, replace=False))
            lo, hi = _ulp_step(m, -d), np.float32(m)
            first, second = (lo, hi) if t % 2 == 0 else (hi, lo)
            x[i, j].flat[p[1]] = first
            x[i, j].flat[p[2]] = second
            if t % 3 == 0:
                x[i, j].flat[p[0]] = _ulp_step(m, -2 * d)
    return torch.from_numpy(x)


# --------------------------------------------------------------------------- SPM

SPM_SHAPES = {
    # name: (K, R, sigma, input_size)
    "coco": (17, 128, 1, 512),            # configs/spm_coco.yaml
    "small": (4, 32, 1, 128),
}


def spm_case(name, n_images=4, seed=4321):
    """people list, target [N,1+2K,R,R] f32 (oracle render), logits (inverse-activated target + noise)."""
    k, res, sigma, in_size = SPM_SHAPES[name]
    people = po.make_config4_people(n_images, k=k, res=res, max_people=8 if res >= 128 else 3, seed=seed)
    target = np.stack([po.spm_render(c, j, res, sigma) for c, j in people])
    logits = po.spm_logits_from_target(target, seed=seed + 1)
    return people, target, logits, dict(k=k, res=res, sigma=sigma, input_size=in_size)


def pack_people(people, pmax=None):
    """Ragged people list -> dense centers [N,Pmax,2] i64, joints [N,Pmax,K,2] i64, counts [N] i32."""
    n = len(people)
    k = people[0][1].shape[1]
    pmax = pmax or max(c.shape[0] for c, _ in people)
    centers = np.zeros((n, pmax, 2), dtype=np.int64)
    joints = np.zeros((n, pmax, k, 2), dtype=np.int64)
    counts = np.zeros((n,), dtype=np.int32)
    for i, (c, j) in enumerate(people):
        p = c.shape[0]
        centers[i, :p] = c[:, 0]
        joints[i, :p] = j
        counts[i] = p
    return centers, joints, counts
