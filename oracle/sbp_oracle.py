"""Simple-Baselines (SBP) oracle: target render, joints-MSE loss, decode, back-projection.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  CPU restatement of

* utils/sbp_utils.py:20-53   (SBPHeatmapGenerator)        -> gauss_template, sbp_render*
* models/loss/sbp_loss.py:20-66 (SBPLoss)                 -> sbp_loss*, sbp_loss_closed_form_f64
* utils/sbp_utils.py:56-118  (nms_sbp, DecodeSBP)         -> sbp_decode*
* utils/sbp_utils.py:131-164 (SBPmAPCOCO.update_state)    -> sbp_backproject, sbp_result_rows
* utils/sbp_pis_utils.py:13-47 (SBPmAPPIS.update_state)   -> sbp_result_rows(pad=18)

Two flavours of each stage: a `*_loop` form that keeps the reference's
per-sample / per-joint control flow (this is what bench.py times as the CPU
baseline, because the loops ARE the reference's cost) and a vectorised form
used by the parity tests on larger inputs.  The two are cross-checked in
tests/test_oracle_selfcheck.py.
"""
import math

import numpy as np
import torch

# --------------------------------------------------------------------------- render


def resolve_sigma(sigma, out_h):
    """sigma < 0 means out_h / 64 (utils/sbp_utils.py:24-25)."""
    return out_h / 64 if sigma < 0 else sigma


def gauss_template(sigma):
    """float64 (6s+3)x(6s+3) Gaussian, centre (3s+1, 3s+1) (utils/sbp_utils.py:27-31)."""
    n = 6 * sigma + 3
    ax = np.arange(0, n, 1, float)
    c = 3 * sigma + 1
    return np.exp(-((ax[None, :] - c) ** 2 + (ax[:, None] - c) ** 2) / (2 * sigma ** 2))


def _patch_bounds(ci, sigma):
    """Half-to-even rounded patch corners for an integer centre (utils/sbp_utils.py:42-43)."""
    lo = int(np.round(ci - 3 * sigma - 1))
    hi = int(np.round(ci + 3 * sigma + 2))
    return lo, hi


def sbp_render_loop(joints, out_h, out_w, sigma=-1):
    """One sample: joints [K,2] (x,y in heat-map pixels; <0 = invisible) -> float32 [K,H,W].

    Same control flow as utils/sbp_utils.py:33-53: skip if x<0 or y<0; truncate
    toward zero then clamp to the map; paste max(old, template slice).
    """
    sigma = resolve_sigma(sigma, out_h)
    g = gauss_template(sigma)
    joints = np.asarray(joints)
    out = np.zeros((joints.shape[0], out_h, out_w), dtype=np.float32)
    for k in range(joints.shape[0]):
        x, y = joints[k]
        if x < 0 or y < 0:
            continue
        cx = min(max(int(x), 0), out_w - 1)
        cy = min(max(int(y), 0), out_h - 1)
        x0, x1 = _patch_bounds(cx, sigma)
        y0, y1 = _patch_bounds(cy, sigma)
        # destination window clipped to the map, and the matching template window
        dx0, dx1 = max(0, x0), min(x1, out_w)
        dy0, dy1 = max(0, y0), min(y1, out_h)
        sx0, sx1 = dx0 - x0, dx1 - x0
        sy0, sy1 = dy0 - y0, dy1 - y0
        out[k, dy0:dy1, dx0:dx1] = np.maximum(out[k, dy0:dy1, dx0:dx1], g[sy0:sy1, sx0:sx1])
    return out


def sbp_render(kp, out_h, out_w, sigma=-1):
    """Batched, vectorised: kp [B,K,2] float64 -> float32 [B,K,H,W].  Same arithmetic as the loop form."""
    sigma = resolve_sigma(sigma, out_h)
    g32 = gauss_template(sigma).astype(np.float32)
    n = g32.shape[0]
    kp = np.asarray(kp, dtype=np.float64)
    x, y = kp[..., 0], kp[..., 1]
    vis = ~((x < 0) | (y < 0))
    # int() truncates toward zero; invisible joints are masked out below so their value is irrelevant
    cx = np.clip(np.trunc(np.where(vis, x, 0.0)).astype(np.int64), 0, out_w - 1)
    cy = np.clip(np.trunc(np.where(vis, y, 0.0)).astype(np.int64), 0, out_h - 1)
    x0 = np.round(cx - 3 * sigma - 1).astype(np.int64)
    y0 = np.round(cy - 3 * sigma - 1).astype(np.int64)
    x1 = np.round(cx + 3 * sigma + 2).astype(np.int64)
    y1 = np.round(cy + 3 * sigma + 2).astype(np.int64)
    cols = np.arange(out_w)[None, None, :]
    rows = np.arange(out_h)[None, None, :]
    gx = cols - x0[..., None]          # [B,K,W] template column of every map column
    gy = rows - y0[..., None]          # [B,K,H]
    okx = (cols >= x0[..., None]) & (cols < x1[..., None]) & (gx < n)
    oky = (rows >= y0[..., None]) & (rows < y1[..., None]) & (gy < n)
    gxc = np.clip(gx, 0, n - 1)
    gyc = np.clip(gy, 0, n - 1)
    val = g32[gyc[..., :, None], gxc[..., None, :]]            # [B,K,H,W]
    m = oky[..., :, None] & okx[..., None, :] & vis[..., None, None]
    return np.where(m, val, np.float32(0)).astype(np.float32)


# --------------------------------------------------------------------------- loss


def sbp_loss(logits, target, lambda_pos=5, lambda_neg=1):
    """Reference op order (models/loss/sbp_loss.py:29-49), CPU fp32.  Returns 0-dim tensor (autograd-capable).

    NHWC permute, sigmoid, per-element mask t>0, two sum-MSE terms weighted 5 / 1,
    each / (2K), total / B.
    """
    b = logits.size(0)
    p = torch.sigmoid(logits.permute(0, 2, 3, 1).contiguous())
    k = p.size(-1)
    t = target.permute(0, 2, 3, 1).contiguous()
    pos = torch.where(t > 0., 1., 0.).type(torch.float32)
    neg = torch.where(t > 0., 0., 1.).type(torch.float32)
    sse = torch.nn.functional.mse_loss
    l_pos = lambda_pos * sse(p * pos, t, reduction='sum') / (k * 2)
    l_neg = lambda_neg * sse(p * neg, t * neg, reduction='sum') / (k * 2)
    return (l_pos + l_neg) / b


def sbp_loss_and_grad(logits, target):
    """(loss float32 scalar tensor, dlogits [B,K,H,W]) via autograd through `sbp_loss`."""
    x = logits.detach().clone().requires_grad_(True)
    loss = sbp_loss(x, target)
    loss.backward()
    return loss.detach(), x.grad.detach()


def sbp_loss_closed_form_f64(logits, target, lambda_pos=5.0, lambda_neg=1.0):
    """float64 closed form of the same loss and gradient.

    t>0:  5 (s-t)^2          grad 5*2 (s-t) s (1-s)
    t<=0: (s-t)^2 + 5 t^2    grad   2 (s-t) s (1-s)        all / (2 K B)
    Used as the high-precision yardstick that both the fp32 oracle and the CUDA path are measured against.
    """
    x = logits.double()
    t = target.double()
    b, k = x.shape[0], x.shape[1]
    s = torch.sigmoid(x)
    posm = t > 0
    e = torch.where(posm, lambda_pos * (s - t) ** 2, lambda_neg * (s - t) ** 2 + lambda_pos * t ** 2)
    norm = 2.0 * k * b
    w = torch.where(posm, torch.full_like(s, lambda_pos), torch.full_like(s, lambda_neg))
    grad = w * 2.0 * (s - t) * s * (1.0 - s) / norm
    return e.sum() / norm, grad


# --------------------------------------------------------------------------- decode


def sbp_decode_loop(x, input_w, conf_threshold, pred=True):
    """One sample, reference control flow: x [1,K,H,W] -> [K,3] (x_px, y_px, conf) float32.

    utils/sbp_utils.py:103-118 + :68-82.  Candidates are pixels with h > thr in row-major
    order; the winner is the first maximum among them; rows with no candidate stay -1 and
    are then scaled too (-> -4 for a 4x ratio); both x and y use input_w / W_out.
    """
    assert x.size(0) == 1
    w_out = x.size(-1)
    h = torch.sigmoid(x) if pred else x
    maps = h[0]
    k = maps.size(0)
    joints = torch.zeros((k, 3)) - 1
    for j in range(k):
        hm = maps[j]
        yy, xx = torch.where(hm > conf_threshold)
        if yy.numel() == 0:
            continue
        vals = hm[yy, xx]
        a = int(np.argmax(vals.numpy()))          # first occurrence of the maximum
        joints[j, 0] = xx[a]
        joints[j, 1] = yy[a]
        joints[j, 2] = vals[a]
    joints[..., :2] *= (input_w / w_out)
    return joints


def sbp_decode_loop_torch(x, input_w, conf_threshold, pred=True):
    """The reference's decode as the ATen OP CHAIN it is (utils/sbp_utils.py:56-82, :103-118), device-agnostic: per joint a
    comparison, torch.where (-> nonzero, host sync), an advanced-index gather, torch.argmax, three element reads (host syncs)
    and a torch.tensor() H2D write.  bench.py times it on CUDA tensors as the `aten_baseline` of decode."""
    assert x.size(0) == 1
    w_out = x.size(-1)
    heatmaps = torch.sigmoid(x) if pred else x
    hm = heatmaps[0]
    k = hm.size(0)
    joints = torch.zeros((k, 3), device=hm.device) - 1
    for idx in torch.arange(k):
        heatmap = hm[idx]
        yy, xx = torch.where(heatmap > conf_threshold)
        if yy.size(0) == 0:
            continue
        conf = heatmap[yy, xx]
        a = torch.argmax(conf)
        joints[idx] = torch.tensor([xx[a], yy[a], conf[a]])
    joints[..., :2] *= (input_w / w_out)
    return joints


def reference_sigmoid(x):
    """torch.sigmoid applied the way DecodeSBP.forward does (utils/sbp_utils.py:104-109): one [1,K,H,W] sample per call.
    Which elements go through ATen's vectorised body (Sleef expf) and which through its scalar tail (glibc expf) depends on
    the size of the tensor handed to torch.sigmoid, so the batched oracle keeps the per-sample calls."""
    if x.device.type != "cpu" or x.size(0) <= 1:
        return torch.sigmoid(x)
    return torch.cat([torch.sigmoid(x[b:b + 1]) for b in range(x.size(0))])


def torch_sigmoid_vector_body(x):
    """torch.sigmoid of a flat fp32 CPU tensor with EVERY element going through ATen's vectorised body (Sleef expf): one
    thread (no chunk boundaries) and a length padded to whole vector pairs.  For pinning the restatements of that body."""
    n = x.numel()
    pad = (-n) % 64
    xp = torch.cat([x.reshape(-1), x.new_zeros(pad)]) if pad else x.reshape(-1)
    threads = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        y = torch.sigmoid(xp)
    finally:
        torch.set_num_threads(threads)
    return y[:n].reshape(x.shape)


def sbp_decode(x, input_w, conf_threshold, pred=True):
    """Batched, vectorised: x [B,K,H,W] fp32 -> [B,K,3] fp32.  Same semantics as the loop form.  Works on CPU tensors (the
    oracle proper: ATen's CPU sigmoid) and on CUDA tensors (ATen's CUDA sigmoid -- what the reference computes when its
    Lightning module runs on a GPU); the result is returned on the CPU."""
    b, k, hh, ww = x.shape
    h = reference_sigmoid(x) if pred else x
    flat = h.reshape(b * k, hh * ww)
    thr = torch.tensor(conf_threshold, dtype=torch.float32, device=x.device)      # torch compares in the tensor dtype
    cand = flat > thr
    masked = torch.where(cand, flat, torch.full_like(flat, -float("inf")))
    best = masked.max(dim=1, keepdim=True).values
    pos = torch.arange(hh * ww, device=x.device).expand_as(flat)
    idx = torch.where(masked == best, pos, torch.full_like(pos, hh * ww)).min(dim=1).values   # first row-major maximum
    any_c = cand.any(dim=1)
    out = torch.full((b * k, 3), -1.0, dtype=torch.float32, device=x.device)
    idx_c = idx.clamp(max=hh * ww - 1)
    out[:, 0] = torch.where(any_c, (idx_c % ww).float(), out[:, 0])
    out[:, 1] = torch.where(any_c, (idx_c // ww).float(), out[:, 1])
    out[:, 2] = torch.where(any_c, flat.gather(1, idx_c[:, None])[:, 0], out[:, 2])
    t = out.reshape(b, k, 3).cpu()
    t[..., :2] *= (input_w / ww)
    return t


def sbp_flip_average(x, x_flipped, flip_pairs, pred=True):
    """PARITY UNPINNED (not in the reference).

    Published Simple-Baselines flip test: the maps of the mirrored image are mirrored back (columns reversed), left /
    right joints swapped, and averaged with the maps of the image: (h + flip_back(h_f)) * 0.5, in fp32.
    x, x_flipped [B,K,H,W] -> post-activation averaged maps [B,K,H,W] (decode them with pred=False).
    """
    h = torch.sigmoid(x) if pred else x
    hf = torch.sigmoid(x_flipped) if pred else x_flipped
    perm = list(range(x.size(1)))
    for a, b in flip_pairs:
        perm[a], perm[b] = b, a
    back = hf[:, perm].flip(-1)
    return (h + back) * 0.5


def sbp_refine_quarter(joints_px, heat):
    """PARITY UNPINNED (not in the reference).

    Published Simple-Baselines rule on heat-map-pixel coordinates: for an interior peak
    (1 < px < W-1 and 1 < py < H-1) shift by 0.25 * sign of the central difference.
    joints_px [B,K,3] in heat-map pixels (conf<0 rows untouched), heat [B,K,H,W] (post-activation).
    """
    out = joints_px.clone()
    b, k, hh, ww = heat.shape
    for i in range(b):
        for j in range(k):
            if out[i, j, 2] < 0:
                continue
            px, py = int(out[i, j, 0]), int(out[i, j, 1])
            if 1 < px < ww - 1 and 1 < py < hh - 1:
                dx = float(heat[i, j, py, px + 1] - heat[i, j, py, px - 1])
                dy = float(heat[i, j, py + 1, px] - heat[i, j, py - 1, px])
                out[i, j, 0] += 0.25 * float(np.sign(dx))
                out[i, j, 1] += 0.25 * float(np.sign(dy))
    return out


# --------------------------------------------------------------------------- back-projection / COCO rows


def sbp_backproject(joints, bbox, input_size):
    """joints [B,K,3] fp32 (input-size scale), bbox [B,4] float64 (x,y,w,h), input_size [H_in,W_in].

    utils/sbp_utils.py:141-146.  The ratio is a float64 0-dim tensor; multiplying an fp32
    tensor by it rounds the ratio to fp32 first, then an fp32 multiply, then an fp32 add
    of fp32(bbox origin) -- two roundings, no FMA.
    """
    out = joints.clone()
    for i in range(out.size(0)):
        j = out[i]
        j[..., :1] *= (bbox[i][2] / input_size[1])
        j[..., 1:2] *= (bbox[i][3] / input_size[0])
        j[..., :1] += bbox[i][0]
        j[..., 1:2] += bbox[i][1]
    return out


def sbp_result_rows(joints_img, image_ids, category_ids, pad=0):
    """COCO-results rows from back-projected joints (utils/sbp_utils.py:148-164).

    conf<0 -> (0,0,0) and contributes 0 to the score; otherwise (x, y, 1) and conf.
    score = left-to-right fp32 sum of confs / K.  `pad` zeros are appended for the
    PIS variant (utils/sbp_pis_utils.py:40).
    """
    rows = []
    for i in range(joints_img.size(0)):
        kps, confs = [], []
        for (x, y, c) in joints_img[i]:
            if c < 0:
                kps.extend([0, 0, 0])
                confs.append(0)
                continue
            kps.extend([float(x), float(y), 1])
            confs.append(c)
        kps.extend([0] * pad)
        rows.append({
            "image_id": int(image_ids[i]),
            "category_id": int(category_ids[i]),
            "keypoints": kps,
            "score": float(sum(confs) / joints_img.size(1)),
        })
    return rows


# --------------------------------------------------------------------------- synthetic inputs (SURVEY.md 8 d)


def make_config1_inputs(batch=32, k=17, h=64, w=48, seed=1234, torch_seed=0):
    """Seeded synthetic inputs of BASELINE.json configs[0]: kp, logits, bbox, ids."""
    rng = np.random.default_rng(seed)
    kp = np.stack([rng.uniform(0, w, (batch, k)), rng.uniform(0, h, (batch, k))], axis=-1)
    vis = rng.uniform(size=(batch, k)) < 0.85
    kp[~vis] = -1.0
    bbox = np.stack([rng.uniform(0, 400, batch), rng.uniform(0, 400, batch),
                     rng.uniform(40, 300, batch), rng.uniform(60, 400, batch)], axis=-1)
    gen = torch.Generator().manual_seed(torch_seed)
    logits = torch.randn(batch, k, h, w, generator=gen)
    image_id = torch.arange(batch, dtype=torch.int64) + 1000
    category_id = torch.ones(batch, dtype=torch.int64)
    return kp, logits, torch.from_numpy(bbox), image_id, category_id


def realistic_logits(kp, h, w, sigma=2, jitter=1.0, noise=0.05, seed=7):
    """Logits whose sigmoid resembles a trained net's output: logit(clip(render(kp+jitter)+noise))."""
    rng = np.random.default_rng(seed)
    kpj = np.where(kp < 0, kp, kp + rng.uniform(-jitter, jitter, kp.shape))
    kpj = np.where((kp >= 0) & (kpj < 0), 0.0, kpj)
    t = sbp_render(kpj, h, w, sigma).astype(np.float64)
    p = np.clip(t + noise * rng.uniform(size=t.shape), 1e-4, 1 - 1e-4)
    return torch.from_numpy(np.log(p / (1 - p)).astype(np.float32))
