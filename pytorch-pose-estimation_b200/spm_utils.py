"""SPM utilities with the reference's names and call signatures (utils/spm_utils.py) on sm_100a kernels.

The reference builds one image at a time with three generator objects (root heat map, per-person
masks, displacement maps); here the three keep their per-image NumPy contracts and
`spm_render_batch` renders whole batches [N,1+2K,R,R] on the device in one launch.
"""
import json
import math
import os

import numpy as np
import torch
from torch import nn

from . import _cabi
from ._cabi import check, dense, lib, ptr, stream_ptr
from .sbp_utils import _gauss_template, _ids, _load_coco, _templates


def _device(device=None):
    if device is not None:
        return torch.device(device)
    if not torch.cuda.is_available():
        raise _cabi.PoseB200Error("pose_b200 needs a CUDA device (no CPU path)")
    return torch.device("cuda", torch.cuda.current_device())


def _i64(a, device):
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=torch.int64).contiguous()
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.int64))).to(device)


def _render_args(centers, joints, counts, output_res, sigma, dev):
    """Device-side person arrays + Gaussian template shared by `spm_render_batch` and `spm_loss.spm_fused`."""
    c = _i64(centers, dev)
    j = _i64(joints, dev)
    n, pmax = j.size(0), j.size(1)
    assert tuple(c.shape) == (n, pmax, 2) and j.dim() == 4 and j.size(3) == 2
    cnt = (counts.to(device=dev, dtype=torch.int32) if isinstance(counts, torch.Tensor)
           else torch.from_numpy(np.asarray(counts, dtype=np.int32)).to(dev)).contiguous()
    assert cnt.numel() == n
    sig = float(output_res / 64 if sigma < 0 else sigma)
    g = _gauss_template(sig)
    return c, j, cnt, sig, _templates.get(g, sig, dev), g.shape[0]


def spm_render_batch(centers, joints, counts, output_res, sigma=-1, device=None):
    """centers [N,Pmax,2] i64, joints [N,Pmax,K,2] i64, counts [N] i32 -> CUDA fp32 [N,1+2K,R,R].

    Channel 0 = root heat map (SPMHeatmapGenerator), channels 1.. = displacement maps
    (SPMMaskGenerator + SPMDisplacementGenerator), as concatenated at dataset/spm_coco_dataset.py:86.
    """
    dev = _device(device if device is not None else (centers.device if isinstance(centers, torch.Tensor) and centers.is_cuda else None))
    c, j, cnt, sig, lut, lut_n = _render_args(centers, joints, counts, output_res, sigma, dev)
    n, pmax, k = j.size(0), j.size(1), j.size(2)
    out = torch.empty((n, 1 + 2 * k, output_res, output_res), dtype=torch.float32, device=dev)
    ws = _cabi.workspace(dev, int(lib().pose_spm_fused_workspace_bytes(n, k, output_res)))      # single-pass render needs the geometry records
    with torch.cuda.device(dev):
        check(lib().pose_spm_render(ptr(c), ptr(j), ptr(cnt), ptr(out), n, pmax, k, output_res, sig, ptr(lut), lut_n,
                                    ptr(ws), ws.numel(), stream_ptr(dev)), "pose_spm_render")
    return out


def _render_one(centers, joints, output_res, sigma):
    """One image on the device: centers [P,2], joints [P,K,2] (int64 numpy) -> CUDA fp32 [1+2K,R,R]."""
    p = centers.shape[0]
    k = joints.shape[1]
    c = centers.reshape(1, p, 2) if p else np.zeros((1, 1, 2), dtype=np.int64)
    j = joints.reshape(1, p, k, 2) if p else np.zeros((1, 1, k, 2), dtype=np.int64)
    return spm_render_batch(c, j, np.array([p], dtype=np.int32), output_res, sigma)[0]


################################################################################################################
# Single-Stage Multi-Person Pose Machines Utils
################################################################################################################
class SPMHeatmapGenerator:
    """Drop-in for utils/spm_utils.py:16-47: `__call__(joints [P,J,2])` -> np.float32 [J,R,R], per-joint max over persons."""

    def __init__(self, output_res, num_joints, sigma=-1):
        self.output_res = output_res
        self.num_joints = num_joints
        if sigma < 0:
            sigma = self.output_res / 64
        self.sigma = sigma
        self.g = _gauss_template(sigma)

    def __call__(self, joints):
        jt = np.asarray(joints, dtype=np.int64).reshape(-1, self.num_joints, 2)
        none = np.zeros((jt.shape[0], 1, 2), dtype=np.int64)            # no body joints: only channel 0 is used
        maps = [_render_one(np.ascontiguousarray(jt[:, idx]), none, self.output_res, self.sigma)[0]
                for idx in range(self.num_joints)]
        return torch.stack(maps).cpu().numpy()


class SPMMasks(np.ndarray):
    """np.float32 [P,R,R] box masks that remember the centres / sigma they were made from (needed by
    SPMDisplacementGenerator, whose kernel rebuilds the boxes from the centres instead of reading dense masks)."""

    def __new__(cls, array, centers, sigma):
        obj = np.asarray(array, dtype=np.float32).view(cls)
        obj.centers, obj.sigma = centers, sigma
        return obj

    def __array_finalize__(self, obj):
        if obj is not None:
            self.centers = getattr(obj, "centers", None)
            self.sigma = getattr(obj, "sigma", None)


class SPMMaskGenerator:
    """Drop-in for utils/spm_utils.py:50-71: `__call__(centers [P,J,2])` -> np.float32 [P,R,R] box masks."""

    def __init__(self, output_res, sigma=-1):
        self.output_res = output_res
        if sigma < 0:
            sigma = self.output_res / 64
        self.sigma = sigma
        self.size = int((6 * sigma + 2) / 2)

    def __call__(self, joints):
        c = np.asarray(joints, dtype=np.int64)
        c = c.reshape(c.shape[0], -1, 2)
        p, r = c.shape[0], self.output_res
        if p == 0:
            return SPMMasks(np.zeros((0, r, r), dtype=np.float32), c, self.sigma)
        # one single-person image per person whose only body joint lies right of every column: its x-displacement is
        # strictly positive exactly on the person's box, so mask = (disp_x != 0) comes straight out of the render kernel
        far = np.full((p, 1, 1, 2), r + 10, dtype=np.int64)
        mask = None
        for j in range(c.shape[1]):
            t = spm_render_batch(np.ascontiguousarray(c[:, j]).reshape(p, 1, 2), far, np.ones(p, dtype=np.int32), r, self.sigma)
            m = t[:, 1] != 0
            mask = m if mask is None else (mask | m)
        return SPMMasks(mask.to(torch.float32).cpu().numpy(), c, self.sigma)


class SPMDisplacementGenerator:
    """Drop-in for utils/spm_utils.py:74-95: `__call__(joints [P,K,2], masks)` -> np.float32 [2K,R,R].

    `masks` must come from this package's SPMMaskGenerator (one centre per person, as in
    dataset/spm_coco_dataset.py:80): the kernel rebuilds each box from its centre.
    """

    def __init__(self, output_res, num_joints):
        self.output_res = output_res
        self.num_joints = num_joints
        self.z = math.sqrt(output_res ** 2 + output_res ** 2)

    def __call__(self, joints, masks):
        if not isinstance(masks, SPMMasks) or masks.centers is None:
            raise TypeError("masks must be the SPMMasks returned by pose_b200's SPMMaskGenerator")
        if masks.centers.shape[1] != 1:
            raise ValueError("SPMDisplacementGenerator supports one centre per person")
        j = np.asarray(joints, dtype=np.int64).reshape(-1, self.num_joints, 2)
        return _render_one(np.ascontiguousarray(masks.centers[:, 0]), j, self.output_res, masks.sigma)[1:].cpu().numpy()


def spm_decode_batch(x, input_size, sigma, conf_threshold, pred=True, max_people=64, dist_threshold=None, sigmoid_ref=None):
    """x [N,1+2K,R,R] CUDA -> (roots [N,Pmax,3], kps [N,Pmax,K,3], counts [N], counts_total [N]) on the device.
    Only the first counts[i] rows of image i are defined.  `sigmoid_ref` ("cpu" | "cuda"): with `pred`, the root confidences
    -- hence the threshold test and the greedy order -- are the reference's torch.sigmoid on CPU / CUDA tensors bit for bit."""
    t = dense(x, "x")
    n, c, r, _ = t.shape
    k = (c - 1) // 2
    dev = t.device
    # rows >= counts[i] are never written (and never read by the drop-ins): no memset launches in front of the kernel
    roots = torch.empty((n, max_people, 3), dtype=torch.float32, device=dev)
    kps = torch.empty((n, max_people, k, 3), dtype=torch.float32, device=dev)
    counts = torch.empty((n,), dtype=torch.int32, device=dev)
    total = torch.empty((n,), dtype=torch.int32, device=dev)
    dist = (6 * sigma + 2) / 2 if dist_threshold is None else dist_threshold
    with torch.cuda.device(dev):
        check(lib().pose_spm_decode(ptr(t), ptr(roots), ptr(kps), ptr(counts), ptr(total), n, max_people, k, r,
                                    float(conf_threshold), float(dist), int(bool(pred)), _cabi.sigmoid_ref_code(sigmoid_ref),
                                    float(input_size), stream_ptr(dev)), "pose_spm_decode")
    return roots, kps, counts, total


def nms_spm(heatmaps, conf_threshold=0.8, dist_threshold=7., max_people=256):
    """Drop-in for utils/spm_utils.py:98-161: heatmaps [1,R,R] (activated) -> [N,3] = [x, y, conf] (or empty 1-D)."""
    h = dense(heatmaps, "heatmaps")
    r = h.size(-1)
    # reuse the decode kernel on a 1-joint problem with zero displacements (K=1 -> 3 channels)
    x = torch.zeros((1, 3, r, r), dtype=torch.float32, device=h.device)
    x[0, 0] = h[0]
    roots, _, counts, total = spm_decode_batch(x, r, 0, conf_threshold, pred=False, max_people=max_people,
                                               dist_threshold=dist_threshold)
    n = int(counts[0])
    if n == 0:
        return torch.zeros((0,), dtype=torch.float32, device=h.device)
    if int(total[0]) > max_people:
        return nms_spm(heatmaps, conf_threshold, dist_threshold, max_people=int(total[0]))
    return roots[0, :n].clone()


def get_spm_keypoints(root_joints, displacements, dist_threshold):
    """Drop-in for utils/spm_utils.py:164-200: root_joints [N,3], displacements [2K,R,R] -> [N,K,3] (map pixels)."""
    if root_joints.size(0) == 0:
        return root_joints
    r = dense(root_joints, "root_joints")
    d = dense(displacements, "displacements")
    k, res = d.size(0) // 2, d.size(-1)
    out = torch.empty((r.size(0), k, 3), dtype=torch.float32, device=r.device)
    with torch.cuda.device(r.device):
        check(lib().pose_spm_gather(ptr(r), ptr(d), ptr(out), r.size(0), k, res, float(dist_threshold), stream_ptr(r.device)),
              "pose_spm_gather")
    return out


def get_spm_keypoints_chained(root_joints, displacements, parents, dist_threshold):
    """Hierarchical displacement chaining -- NOT in the reference (single hop, utils/spm_utils.py:187-189); opt-in, parity unpinned.

    `parents[k]` = the joint that joint k's displacement is relative to (-1 = the root joint).  Same arguments and result as
    `get_spm_keypoints` otherwise; with every parent -1 the result is bit-identical to it."""
    if root_joints.size(0) == 0:
        return root_joints
    r = dense(root_joints, "root_joints")
    d = dense(displacements, "displacements")
    k, res = d.size(0) // 2, d.size(-1)
    par = torch.as_tensor(parents, dtype=torch.int32).reshape(-1)
    assert par.numel() == k, "parents must have one entry per joint"
    assert int(par.max()) < k and int(par.min()) >= -1, "parents must be joint indices or -1"
    par = par.to(r.device)
    out = torch.empty((r.size(0), k, 3), dtype=torch.float32, device=r.device)
    with torch.cuda.device(r.device):
        check(lib().pose_spm_gather_chain(ptr(r), ptr(d), ptr(par), ptr(out), r.size(0), k, res, float(dist_threshold), stream_ptr(r.device)),
              "pose_spm_gather_chain")
    return out


class DecodeSPM(nn.Module):
    """Drop-in for utils/spm_utils.py:203-250; `decode_batch` keeps everything on the device."""

    def __init__(self, input_size, sigma, conf_threshold, pred=True, max_people=64, sigmoid_ref=None):
        super().__init__()
        self.input_size = input_size
        self.sigma = sigma
        self.dist_threshold = (6 * sigma + 2) / 2
        self.conf_threshold = conf_threshold
        self.pred = pred
        self.max_people = max_people
        self.sigmoid_ref = sigmoid_ref

    def decode_batch(self, x):
        return spm_decode_batch(x, self.input_size, self.sigma, self.conf_threshold, self.pred, self.max_people,
                                sigmoid_ref=self.sigmoid_ref)

    def forward(self, x):
        assert x.size(0) == 1
        roots, kps, counts, total = self.decode_batch(x)
        n, tot = int(counts[0]), int(total[0])
        if tot > self.max_people:        # rare: more roots than the fixed buffer -> redo with the exact size
            roots, kps, counts, total = spm_decode_batch(x, self.input_size, self.sigma, self.conf_threshold, self.pred, tot,
                                                         sigmoid_ref=self.sigmoid_ref)
            n = tot
        if n == 0:
            e = torch.zeros((0,), dtype=torch.float32, device=x.device)
            return e, e.clone()
        return roots[0, :n].clone(), kps[0, :n].clone()


def spm_rows_to_results(kps, counts, image_sizes, image_ids, category_ids, input_size):
    """Batched tail of SPMmAPCOCO.update_state (utils/spm_utils.py:302-323): one D2H copy, then python dicts."""
    n, pmax, k, _ = kps.shape
    dev = kps.device
    w = dense(torch.as_tensor(image_sizes[0]).to(dev), "image_w", torch.int64)
    h = dense(torch.as_tensor(image_sizes[1]).to(dev), "image_h", torch.int64)
    kd, cd = dense(kps, "kps"), dense(counts, "counts", torch.int32)
    scaled = torch.empty_like(kd)            # rows >= counts[i] are never written by the kernel nor read below
    with torch.cuda.device(dev):
        check(lib().pose_spm_rescale(ptr(kd), ptr(cd), ptr(w), ptr(h), ptr(scaled), n, pmax, k, float(input_size), stream_ptr(dev)),
              "pose_spm_rescale")
    host = scaled.cpu()
    cnt = counts.cpu().tolist()
    out = []
    for i, (iid, cid) in enumerate(zip(_ids(image_ids), _ids(category_ids))):
        for p in range(cnt[i]):
            flat, score = [], np.float32(0)
            for x, y, c in host[i, p].tolist():
                if x == 0. and y == 0.:
                    flat.extend([0, 0, 0])
                    continue
                flat.extend([x, y, 1])
                score = np.float32(score + np.float32(c))      # left-to-right fp32 sum, as python sum() over fp32 tensors
            out.append({"image_id": int(iid), "category_id": int(cid), "keypoints": flat,
                        "score": float(np.float32(score / np.float32(k)))})
    return out


class SPMmAPCOCO:
    """Drop-in for utils/spm_utils.py:282-351 (batched `update_state`; `result()` runs the OKS / AP kernels of coco_eval.py)."""

    def __init__(self, json_path, input_size, sigma, conf_threshold, max_people=64, gather=False):
        """`gather=True` (not in the reference): under torch.distributed every rank's people are all-gathered in
        `update_state` (fixed-size [B,Pmax,K,3] rows + counts), so each rank's `result_list` covers the whole set."""
        self.coco = _load_coco(json_path)
        self.input_size = input_size
        self.conf_threshold = conf_threshold
        self.decoder = DecodeSPM(input_size, sigma, conf_threshold, True, max_people)
        self.result_list = []
        self.gather = gather
        self._evaluator = None
        self.stats = None

    def reset_states(self):
        self.result_list = []

    def update_state(self, target, y_pred):
        _, kps, counts, total = self.decoder.decode_batch(y_pred)
        need = total.max() if total.numel() else total.new_zeros(())
        if self.gather:
            from . import dist as pd
            if pd._active(None):                          # every rank must gather with the same Pmax
                need = need.clone()
                pd.dist.all_reduce(need, op=pd.dist.ReduceOp.MAX)
        if int(need) > self.decoder.max_people:
            self.decoder.max_people = int(need)
            _, kps, counts, total = self.decoder.decode_batch(y_pred)
        sizes, iid, cid = target['image_size'], target['image_id'], target['category_id']
        if self.gather:
            dev = kps.device
            kps, counts, iid, cid, w, h = pd.gather_spm_people(kps, counts, torch.as_tensor(iid).to(dev), torch.as_tensor(cid).to(dev),
                                                               torch.as_tensor(sizes[0]).to(dev), torch.as_tensor(sizes[1]).to(dev))
            sizes = [w, h]
        self.result_list.extend(spm_rows_to_results(kps, counts, sizes, iid, cid, self.input_size))

    def result(self):
        if not self.result_list:
            return 0
        path = os.path.join(os.getcwd(), 'results.json')
        with open(path, "w") as f:
            json.dump(self.result_list, f, indent=4)
        from .coco_eval import KeypointEval, summarize
        if self.coco is None:
            raise ValueError("result() needs the ground-truth annotations: pass json_path (or a parsed COCO dict)")
        if self._evaluator is None:
            self._evaluator = KeypointEval(self.coco)
        out = self._evaluator.evaluate(self.result_list)
        self.stats = summarize(out['precision'], out['recall'], verbose=True)
        return self.stats[1]
