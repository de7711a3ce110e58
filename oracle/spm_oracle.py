"""SPM (single-stage multi-person) oracle: target render, loss, root NMS, displacement decode.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  CPU restatement of

* utils/spm_utils.py:16-47   SPMHeatmapGenerator      -> spm_root_map
* utils/spm_utils.py:50-71   SPMMaskGenerator         -> spm_masks
* utils/spm_utils.py:74-95   SPMDisplacementGenerator -> spm_displacements
* dataset/spm_coco_dataset.py:77-86 (concat)          -> spm_render
* models/loss/spm_loss.py:23-105 SPMLoss              -> spm_loss*
* utils/spm_utils.py:98-161  nms_spm                  -> spm_nms
* utils/spm_utils.py:164-200 get_spm_keypoints        -> spm_keypoints
* utils/spm_utils.py:225-250 DecodeSPM.forward        -> spm_decode
* utils/spm_utils.py:293-323 SPMmAPCOCO.update_state  -> spm_result_rows

Tie order of equal root confidences is undefined in the reference (non-stable
argsort, utils/spm_utils.py:115); this oracle and the CUDA path both define it
as row-major (stable) and parity is pinned on distinct-confidence inputs.
"""
import math

import numpy as np
import torch

from .sbp_oracle import gauss_template


def _sigma(sigma, res):
    return res / 64 if sigma < 0 else sigma


def spm_root_map(centers, res, sigma=-1):
    """centers [P,1,2] int64 -> float32 [1,R,R]; per-person max of Gaussian patches.

    utils/spm_utils.py:29-47: skip iff x<=0 and y<=0; no clamp (off-map patches clip away).
    """
    sigma = _sigma(sigma, res)
    g = gauss_template(sigma)
    out = np.zeros((1, res, res), dtype=np.float32)
    for person in centers:
        for (x, y) in person:
            if x <= 0 and y <= 0:
                continue
            x0, y0 = int(np.round(x - 3 * sigma - 1)), int(np.round(y - 3 * sigma - 1))
            x1, y1 = int(np.round(x + 3 * sigma + 2)), int(np.round(y + 3 * sigma + 2))
            dx0, dx1 = max(0, x0), min(x1, res)
            dy0, dy1 = max(0, y0), min(y1, res)
            if dx1 <= dx0 or dy1 <= dy0:
                continue
            out[0, dy0:dy1, dx0:dx1] = np.maximum(out[0, dy0:dy1, dx0:dx1],
                                                  g[dy0 - y0:dy1 - y0, dx0 - x0:dx1 - x0])
    return out


def spm_masks(centers, res, sigma=-1):
    """centers [P,1,2] int64 -> float32 [P,R,R] box masks of half-size int((6s+2)/2) (utils/spm_utils.py:55-69)."""
    sigma = _sigma(sigma, res)
    half = int((6 * sigma + 2) / 2)
    m = np.zeros((len(centers), res, res), dtype=np.float32)
    for i, person in enumerate(centers):
        for (x, y) in person:
            if x <= 0 and y <= 0:
                continue
            # raw python slice semantics, as in the reference (domain: coordinates >= 0)
            xs, xe = max(0, x - half), min(res, x + half + 1)
            ys, ye = max(0, y - half), min(res, y + half + 1)
            m[i, slice(ys, ye), slice(xs, xe)] = 1.
    return m


def spm_displacements(joints, masks, res):
    """joints [P,K,2] int64, masks [P,R,R] -> float32 [2K,R,R].

    utils/spm_utils.py:84-95.  For each person, each joint not (x<=0 and y<=0):
    disp[2j] += mask*(x - col)/z, disp[2j+1] += mask*(y - row)/z, z = sqrt(2 R^2).  The
    right-hand side is float64; the in-place add upcasts the fp32 accumulator, adds in
    fp64 and rounds back to fp32 -- persons accumulate (they do not overwrite).
    """
    k = joints.shape[1]
    col = np.arange(res, dtype=np.int64)[None, :].repeat(res, 0)
    row = col.T
    z = math.sqrt(res ** 2 + res ** 2)
    d = np.zeros((2 * k, res, res), dtype=np.float32)
    for p in range(joints.shape[0]):
        for j in range(k):
            x, y = joints[p, j]
            if x <= 0 and y <= 0:
                continue
            d[2 * j] += masks[p] * (x - col) / z
            d[2 * j + 1] += masks[p] * (y - row) / z
    return d


def spm_render(centers, joints, res, sigma=-1):
    """One image: centers [P,1,2] i64, joints [P,K,2] i64 -> float32 [1+2K,R,R] (dataset/spm_coco_dataset.py:77-86)."""
    root = spm_root_map(centers, res, sigma)
    masks = spm_masks(centers, res, sigma)
    disp = spm_displacements(joints, masks, res)
    return np.concatenate([root, disp], axis=0)


# --------------------------------------------------------------------------- loss


def spm_loss(logits, target, lambda_root=1, lambda_disp=0.1):
    """Reference op order (models/loss/spm_loss.py:32-83), CPU fp32, autograd-capable."""
    b = logits.size(0)
    p = logits.permute(0, 2, 3, 1).contiguous()
    root = torch.sigmoid(p[..., :1])
    disp = torch.tanh(p[..., 1:])
    t = target.permute(0, 2, 3, 1).contiguous()
    t_root, t_disp = t[..., :1], t[..., 1:]
    m = torch.where(t_root > 0., 1., 0.).type(torch.float32)
    l_root = lambda_root * torch.nn.functional.mse_loss(root * m, t_root, reduction='sum')
    l_disp = lambda_disp * torch.nn.functional.smooth_l1_loss(disp * m, t_disp, reduction='sum')
    return (l_root + l_disp) / b


def spm_loss_and_grad(logits, target):
    x = logits.detach().clone().requires_grad_(True)
    loss = spm_loss(x, target)
    loss.backward()
    return loss.detach(), x.grad.detach()


def spm_loss_closed_form_f64(logits, target, lambda_root=1.0, lambda_disp=0.1):
    """float64 closed form: root (s*m - t0)^2, disp 0.1*SmoothL1(tanh(p)*m - t), / B."""
    x, t = logits.double(), target.double()
    b = x.shape[0]
    m = (t[:, :1] > 0).double()
    s = torch.sigmoid(x[:, :1])
    e_root = (s * m - t[:, :1]) ** 2
    g_root = lambda_root * 2 * (s * m - t[:, :1]) * m * s * (1 - s) / b
    th = torch.tanh(x[:, 1:])
    d = th * m - t[:, 1:]
    ad = d.abs()
    e_disp = torch.where(ad < 1, 0.5 * d * d, ad - 0.5)
    g_disp = lambda_disp * d.clamp(-1, 1) * m * (1 - th * th) / b
    loss = (lambda_root * e_root.sum() + lambda_disp * e_disp.sum()) / b
    return loss, torch.cat([g_root, g_disp], dim=1)


# --------------------------------------------------------------------------- decode


def spm_nms(heat, conf_threshold=0.8, dist_threshold=7.):
    """heat [1,R,R] fp32 (post-activation) -> [N,3] (x, y, conf) fp32, or an empty 1-D tensor.

    utils/spm_utils.py:112-161.  Candidates h>thr in row-major order, sorted by descending
    confidence (STABLE here), greedy: take the head, keep only those strictly farther than
    dist_threshold, repeat.
    """
    hm = heat[0].numpy()
    ys, xs = np.nonzero(hm > np.float32(conf_threshold))
    if ys.size == 0:
        return torch.zeros((0,), dtype=torch.float32)
    conf = hm[ys, xs]
    order = np.argsort(-conf, kind='stable')
    ys, xs, conf = ys[order].astype(np.int64), xs[order].astype(np.int64), conf[order]
    roots = []
    while ys.size:
        cy, cx, cc = ys[0], xs[0], conf[0]
        roots.append((float(cx), float(cy), cc))
        d = np.sqrt(((xs - cx) ** 2 + (ys - cy) ** 2).astype(np.float64))
        keep = d > dist_threshold
        keep[0] = False
        ys, xs, conf = ys[keep], xs[keep], conf[keep]
    return torch.tensor(np.array(roots, dtype=np.float32))


def spm_keypoints(root_joints, disp, dist_threshold):
    """root_joints [N,3], disp [2K,R,R] (post-activation) -> [N,K,3].

    utils/spm_utils.py:175-200.  kx = disp[2i][y,x]*z + x in fp32 with two roundings;
    d = float64 sqrt of the fp32 squared distance; d < dist_threshold -> (0,0,0).
    """
    if root_joints.size(0) == 0:
        return root_joints
    k2, res, _ = disp.shape
    z = math.sqrt(res ** 2 + res ** 2)
    out = []
    for r in root_joints:
        x, y, c = r
        xi, yi = int(x), int(y)
        row = []
        for i in range(k2 // 2):
            kx = disp[2 * i][yi, xi] * z + x
            ky = disp[2 * i + 1][yi, xi] * z + y
            d = math.sqrt((x - kx) ** 2 + (y - ky) ** 2)
            if d < dist_threshold:
                row.append(torch.zeros(3))
            else:
                row.append(torch.stack([kx, ky, c]))
        out.append(torch.stack(row))
    return torch.stack(out)


def spm_keypoints_chained(root_joints, disp, parents, dist_threshold, max_depth=16):
    """PARITY UNPINNED (not in the reference: get_spm_keypoints is single hop, utils/spm_utils.py:187-189).

    Hierarchical SPR decode in the reference's arithmetic: joint k is read at the decoded position of parents[k] (-1 = the root),
    truncated to a pixel, and added to that position (fp32 mul by z, fp32 add); the fp64 sqrt of the fp32 squared distance to the
    parent `< dist_threshold`, an off-map parent or a chain deeper than `max_depth` make the joint and its descendants (0,0,0)."""
    if root_joints.size(0) == 0:
        return root_joints
    k2, res, _ = disp.shape
    z = math.sqrt(res ** 2 + res ** 2)
    out = []
    for r in root_joints:
        x0, y0, c = r
        row = []
        for k in range(k2 // 2):
            path, j = [], k
            while j >= 0 and len(path) < max_depth:
                path.append(j)
                j = int(parents[j])
            ok = j < 0
            px, py = x0, y0
            for j in reversed(path):
                if not ok:
                    break
                xi, yi = int(px), int(py)
                if not (0 <= xi < res and 0 <= yi < res):
                    ok = False
                    break
                kx = disp[2 * j][yi, xi] * z + px
                ky = disp[2 * j + 1][yi, xi] * z + py
                if math.sqrt((px - kx) ** 2 + (py - ky) ** 2) < dist_threshold:
                    ok = False
                    break
                px, py = kx, ky
            row.append(torch.stack([px, py, c]) if ok else torch.zeros(3))
        out.append(torch.stack(row))
    return torch.stack(out)


def spm_decode(x, input_size, sigma, conf_threshold, pred=True):
    """x [1,1+2K,R,R] -> (root_joints [N,3], keypoints [N,K,3]) at input-size scale (utils/spm_utils.py:225-250)."""
    assert x.size(0) == 1
    res = x.size(-1)
    dist = (6 * sigma + 2) / 2
    if pred:
        heat = torch.sigmoid(x[0, 0:1])
        disp = torch.tanh(x[0, 1:])
    else:
        heat, disp = x[0, 0:1], x[0, 1:]
    roots = spm_nms(heat, conf_threshold, dist)
    kps = spm_keypoints(roots, disp, dist)
    if roots.size(0) == 0:
        return roots, kps
    roots = roots.clone()
    kps = kps.clone()
    roots[..., :2] = roots[..., :2] * input_size / res
    kps[..., :2] = kps[..., :2] * input_size / res
    return roots, kps


def spm_result_rows(x_batch, image_sizes, image_ids, category_ids, input_size, sigma, conf_threshold, pred=True):
    """COCO-results rows (utils/spm_utils.py:293-323): one row per detected person.

    image_sizes = [widths [B], heights [B]] as collated by the reference DataLoader.
    A joint counts as missing iff x==0 and y==0 (after scaling).
    """
    rows = []
    for i in range(x_batch.size(0)):
        _, kps = spm_decode(x_batch[i:i + 1], input_size, sigma, conf_threshold, pred)
        if kps.dim() < 3:
            continue
        kps = kps.clone()
        kps[..., :1] *= (image_sizes[0][i] / input_size)
        kps[..., 1:2] *= (image_sizes[1][i] / input_size)
        for person in kps:
            flat, confs = [], []
            for x, y, c in person:
                if x == 0. and y == 0.:
                    flat.extend([0, 0, 0])
                    confs.append(0)
                    continue
                flat.extend([float(x), float(y), 1])
                confs.append(c)
            rows.append({
                "image_id": int(image_ids[i]),
                "category_id": int(category_ids[i]),
                "keypoints": flat,
                "score": float(sum(confs) / person.size(0)),
            })
    return rows


# --------------------------------------------------------------------------- synthetic inputs (SURVEY.md 8 d, config 4)


def make_config4_people(n_images, k=17, res=128, max_people=8, seed=4321):
    """Per image: centers [P,1,2] i64 and joints [P,K,2] i64 (some joints dropped to (0,0))."""
    rng = np.random.default_rng(seed)
    people = []
    for _ in range(n_images):
        p = int(rng.integers(1, max_people + 1))
        c = rng.integers(8, res - 8, size=(p, 1, 2), dtype=np.int64)
        j = c + rng.integers(-30, 31, size=(p, k, 2), dtype=np.int64)
        j = np.clip(j, 1, res - 2)
        drop = rng.uniform(size=(p, k)) < 0.1
        j[drop] = 0
        people.append((c, j))
    return people


def spm_logits_from_target(target, seed=11, noise=0.02):
    """Invert the activations of a rendered target (+ small noise) with DISTINCT peak confidences.

    root: logit(clip(t0*(0.9+jitter) + noise)), disp: atanh(clip(t + noise)).
    """
    rng = np.random.default_rng(seed)
    t = np.asarray(target, dtype=np.float64)
    out = np.empty_like(t)
    root = t[:, :1] * (0.90 + 0.09 * rng.uniform(size=t[:, :1].shape)) + noise * rng.uniform(size=t[:, :1].shape)
    root = np.clip(root, 1e-4, 1 - 1e-4)
    out[:, :1] = np.log(root / (1 - root))
    d = np.clip(t[:, 1:] + noise * (rng.uniform(size=t[:, 1:].shape) - 0.5), -0.999, 0.999)
    out[:, 1:] = np.arctanh(d)
    return torch.from_numpy(out.astype(np.float32))
