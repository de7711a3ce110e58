"""Simple-Baselines utilities with the reference's names and call signatures (utils/sbp_utils.py).

Every numeric operation runs in the sm_100a kernels behind libpose_b200.so.  Differences from the
reference are additive only: batched entry points (`render_batch`, `decode_batch`) next to the
per-sample drop-ins, and one device-to-host copy per batch in `update_state` instead of ~6*K
synchronisations per sample.
"""
import json
import os

import numpy as np
import torch
from torch import nn

from . import _cabi
from ._cabi import check, dense, lib, ptr, stream_ptr


def _gauss_template(sigma):
    """float64 template of utils/sbp_utils.py:27-31 (size 6s+3, centre 3s+1)."""
    ax = np.arange(0, 6 * sigma + 3, 1, float)
    c = 3 * sigma + 1
    return np.exp(-((ax[None, :] - c) ** 2 + (ax[:, None] - c) ** 2) / (2 * sigma ** 2))


class _TemplateCache:
    """fp32 device copies of Gaussian templates, one per (sigma, device)."""

    def __init__(self):
        self._dev = {}

    def get(self, g64, sigma, device, padded=False):
        """`padded`: the layout pose_sbp_fused reads (n+1 rows of n+8 floats: 4 zero columns either side, an all-zero last row)."""
        key = (float(sigma), device.type, device.index, bool(padded))
        t = self._dev.get(key)
        if t is None:
            g = g64.astype(np.float32)
            if padded:
                g = np.pad(g, ((0, 1), (4, 4)))
            t = torch.from_numpy(np.ascontiguousarray(g)).to(device)
            self._dev[key] = t
        return t


_templates = _TemplateCache()


def template_on(device, sigma):
    g = _gauss_template(sigma)
    return _templates.get(g, sigma, device), g.shape[0]


def _kp_tensor(kp, device):
    """Keypoints as a dense CUDA tensor [.., 2], fp64 preserved (truncation happens in fp64 in the reference)."""
    if isinstance(kp, np.ndarray):
        kp = torch.from_numpy(np.ascontiguousarray(kp))
    if not isinstance(kp, torch.Tensor):
        kp = torch.as_tensor(np.asarray(kp, dtype=np.float64))
    if kp.dtype not in (torch.float32, torch.float64):
        kp = kp.to(torch.float64)
    return kp.to(device, non_blocking=True).contiguous()


################################################################################################################
# Simple Baselines Pose Estimation Utils
################################################################################################################
class SBPHeatmapGenerator:
    """Drop-in for utils/sbp_utils.py:20-53.  `__call__` keeps the per-sample NumPy contract;
    `render_batch` is the batched device entry point the training loop should use."""

    def __init__(self, output_res, num_joints, sigma=-1, device=None):
        self.output_res_h, self.output_res_w = output_res
        self.num_joints = num_joints
        if sigma < 0:
            sigma = self.output_res_h / 64
        self.sigma = sigma
        self.g = _gauss_template(sigma)
        self.device = torch.device(device) if device is not None else None

    def _dev(self):
        if self.device is None:
            if not torch.cuda.is_available():
                raise _cabi.PoseB200Error("SBPHeatmapGenerator needs a CUDA device: pose_b200 has no CPU path")
            self.device = torch.device("cuda", torch.cuda.current_device())
        return self.device

    def render_batch(self, joints, out=None):
        """joints [B,K,2] (numpy / tensor, fp32 or fp64) -> CUDA fp32 [B,K,H,W]."""
        dev = joints.device if isinstance(joints, torch.Tensor) and joints.is_cuda else self._dev()
        kp = _kp_tensor(joints, dev)
        assert kp.dim() == 3 and kp.size(-1) == 2, "joints must be [B,K,2]"
        b, k = kp.size(0), kp.size(1)
        if out is None:
            out = torch.empty((b, k, self.output_res_h, self.output_res_w), dtype=torch.float32, device=dev)
        lut = _templates.get(self.g, self.sigma, dev)
        with torch.cuda.device(dev):
            check(lib().pose_sbp_render(ptr(kp), _cabi.KP_F64 if kp.dtype == torch.float64 else _cabi.KP_F32, ptr(out),
                                        b, k, self.output_res_h, self.output_res_w, float(self.sigma), ptr(lut),
                                        self.g.shape[0], stream_ptr(dev)), "pose_sbp_render")
        return out

    def __call__(self, joints):
        """joints [K,2] -> np.float32 [K,H,W] (reference contract; runs the same kernel with B=1)."""
        kp = np.asarray(joints, dtype=np.float64).reshape(1, -1, 2)
        return self.render_batch(kp)[0].cpu().numpy()


COCO_FLIP_PAIRS = [[1, 2], [3, 4], [5, 6], [7, 8], [9, 10], [11, 12], [13, 14], [15, 16]]

_perm_cache = {}


def _flip_perm(flip_pairs, k, device):
    """[K] int32 on `device`: channel of the mirrored pass that holds joint j (pairs swap, the rest stay)."""
    key = (tuple(map(tuple, flip_pairs)), k, str(device))
    t = _perm_cache.get(key)
    if t is None:
        perm = list(range(k))
        for a, b in flip_pairs:
            perm[a], perm[b] = b, a
        t = _perm_cache[key] = torch.tensor(perm, dtype=torch.int32, device=device)
    return t


def decode_batch(heatmaps, conf_threshold, coord_scale=1.0, apply_sigmoid=False, refine=False, sigmoid_ref=None,
                 flipped=None, flip_pairs=COCO_FLIP_PAIRS, tma=None):
    """[B,K,H,W] CUDA fp32 -> [B,K,3] (x*scale, y*scale, conf); undetected rows are (-scale,-scale,-1).

    `sigmoid_ref` ("cpu" | "cuda", default `_cabi.DEFAULT_SIGMOID_REF`): with `apply_sigmoid`, argmax indices and confidences
    are bit-identical to the reference evaluated with torch.sigmoid on CPU / on CUDA tensors (they differ by an ulp, which
    decides which of two nearly equal logits is "the first maximum").

    `flipped` (not in the reference, opt-in): the maps the network produced for the horizontally mirrored images; they
    are mirrored back, left/right joints swapped (`flip_pairs`) and averaged with `heatmaps` inside the kernel.

    `tma` (default `_cabi.DEFAULT_TMA`, on): stage the maps through shared memory with bulk async copies -- the fast path;
    False selects the register-staged kernel (same results bit for bit)."""
    x = dense(heatmaps, "heatmaps")
    assert x.dim() == 4
    b, k, h, w = x.shape
    out = torch.empty((b, k, 3), dtype=torch.float32, device=x.device)
    if b == 0:
        return out
    if flipped is not None:
        xf = dense(flipped, "flipped")
        assert xf.shape == x.shape and xf.device == x.device, "flipped must match heatmaps"
        perm = _flip_perm(flip_pairs, k, x.device)
        with torch.cuda.device(x.device):
            check(lib().pose_sbp_decode_flip(ptr(x), ptr(xf), ptr(perm), ptr(out), b, k, h, w, float(conf_threshold),
                                             int(bool(apply_sigmoid)), float(coord_scale), int(bool(refine)),
                                             stream_ptr(x.device)), "pose_sbp_decode_flip")
        return out
    with torch.cuda.device(x.device):
        check(lib().pose_sbp_decode(ptr(x), ptr(out), b, k, h, w, float(conf_threshold), int(bool(apply_sigmoid)),
                                    float(coord_scale), int(bool(refine)) | (0 if (tma if tma is not None else _cabi.DEFAULT_TMA) else 2),
                                    _cabi.sigmoid_ref_code(sigmoid_ref), stream_ptr(x.device)), "pose_sbp_decode")
    return out


def nms_sbp(heatmaps, conf_threshold=0.8):
    """Drop-in for utils/sbp_utils.py:56-82: heatmaps [K,H,W] -> joints [K,3] = [x, y, confidence]."""
    return decode_batch(heatmaps[None], conf_threshold, 1.0, False)[0]


class DecodeSBP(nn.Module):
    """Drop-in for utils/sbp_utils.py:85-118 (same ctor / forward), plus `decode_batch` for B > 1."""

    def __init__(self, input_size, conf_threshold, pred=True, refine=False, sigmoid_ref=None, flip_pairs=COCO_FLIP_PAIRS):
        super().__init__()
        self.input_size = input_size[-1]
        self.conf_threshold = conf_threshold
        self.pred = pred
        self.refine = refine          # quarter-pixel refinement: NOT in the reference, default off
        self.sigmoid_ref = sigmoid_ref   # "cpu" | "cuda": which torch.sigmoid ranks near-ties (None: the package default)
        self.flip_pairs = flip_pairs  # used only when the mirrored pass is handed to forward / decode_batch (flip test)

    def decode_batch(self, x, x_flipped=None):
        return decode_batch(x, self.conf_threshold, self.input_size / x.size(-1), self.pred, self.refine, self.sigmoid_ref,
                            flipped=x_flipped, flip_pairs=self.flip_pairs)

    def forward(self, x, x_flipped=None):
        assert x.size(0) == 1
        return self.decode_batch(x, x_flipped)[0]


def backproject_packed(joints, bbox, input_size):
    """joints [B,K,3] (input scale), bbox [B,4] fp64 -> packed [B, 3K+1]: K rows (x_img, y_img, 1|0) then the score."""
    j = dense(joints, "joints")
    bb = dense(bbox.to(j.device) if isinstance(bbox, torch.Tensor) else torch.as_tensor(np.asarray(bbox)).to(j.device),
               "bbox", torch.float64)
    b, k = j.size(0), j.size(1)
    packed = torch.empty((b, 3 * k + 1), dtype=torch.float32, device=j.device)
    with torch.cuda.device(j.device):
        check(lib().pose_sbp_backproject(ptr(j), ptr(bb), ptr(packed), b, k, int(input_size[0]), int(input_size[1]),
                                         stream_ptr(j.device)), "pose_sbp_backproject")
    return packed


def backproject_rows(joints, bbox, input_size):
    """-> (rows [B,K,3] = (x_img, y_img, 1|0), score [B]) as views of the packed buffer."""
    p = backproject_packed(joints, bbox, input_size)
    k = joints.size(1)
    return p[:, :3 * k].unflatten(1, (k, 3)), p[:, 3 * k]


def _ids(v):
    return v.tolist() if isinstance(v, torch.Tensor) else [int(i) for i in v]


def packed_to_results(packed, image_ids, category_ids, pad=0, arrays=None):
    """One D2H copy, then plain-Python COCO result dicts (utils/sbp_utils.py:148-164).  `arrays` (a list) additionally receives
    (kp [B,3K] f64 as emitted, score [B] f64, image ids, category ids): what `result()` hands to the evaluator instead of
    walking the dicts again."""
    host = packed.cpu()
    kps, sc = host[:, :-1], host[:, -1]
    if arrays is not None:
        k3 = kps.double().numpy().copy()
        k3[:, 2::3] = (k3[:, 2::3] != 0)
        arrays.append((np.concatenate([k3, np.zeros((k3.shape[0], pad))], axis=1) if pad else k3, sc.double().numpy().copy(),
                       [int(i) for i in _ids(image_ids)], [int(c) for c in _ids(category_ids)]))
    out = []
    tail = [0] * pad
    for i, (iid, cid) in enumerate(zip(_ids(image_ids), _ids(category_ids))):
        flat = []
        for x, y, v in kps[i].reshape(-1, 3).tolist():
            flat.extend([x, y, 1] if v != 0 else [0, 0, 0])     # same literals as the reference emits
        out.append({"image_id": int(iid), "category_id": int(cid), "keypoints": flat + tail, "score": float(sc[i])})
    return out


class SBPmAPCOCO:
    """Drop-in for utils/sbp_utils.py:121-189.  `update_state` is batched on the device;
    `result()` evaluates OKS / AP with this package's own kernels (coco_eval.py) -- pycocotools is not needed."""

    _pad = 0

    def __init__(self, json_path, input_size, conf_threshold, gather=False):
        """`gather=True` (not in the reference): under torch.distributed every rank's rows are all-gathered in
        `update_state`, so each rank's `result_list` covers the whole validation set -- the reference lets every rank
        write its own shard to the same ./results.json (utils/sbp_utils.py:167-169)."""
        self.coco = _load_coco(json_path)
        self.input_size = input_size
        self.decoder = DecodeSBP(input_size, conf_threshold, True)
        self.result_list = []
        self.gather = gather
        self._evaluator = None
        self._arrays = []               # the arrays result_list was made from (used by result() while the two agree in length)
        self.stats = None

    def reset_states(self):
        self.result_list = []
        self._arrays = []

    def update_state(self, target, y_pred):
        joints = self.decoder.decode_batch(y_pred)                       # [B,K,3] at input scale
        packed = backproject_packed(joints, target['bbox'], self.input_size)
        iid, cid = target['image_id'], target['category_id']
        if self.gather:
            from . import dist as pd
            k = joints.size(1)
            dev = packed.device
            rows, score, iid, cid = pd.gather_rows_ragged(packed[:, :3 * k].unflatten(1, (k, 3)), packed[:, 3 * k],
                                                          torch.as_tensor(iid).to(dev), torch.as_tensor(cid).to(dev))
            packed = torch.cat([rows.flatten(1), score[:, None]], dim=1)
        self.result_list.extend(packed_to_results(packed, iid, cid, self._pad, arrays=self._arrays))

    def result(self):
        """AP at OKS 0.50 (`cocoEval.stats[1]`, utils/sbp_utils.py:189).  Writes ./results.json like the reference, then runs
        the OKS / AP kernels (coco_eval.KeypointEval) instead of pycocotools; all ten numbers are kept in `self.stats`."""
        path = os.path.join(os.getcwd(), 'results.json')
        with open(path, "w") as f:
            json.dump(self.result_list, f, indent=4)
        if not self.result_list:
            raise IndexError("list index out of range")       # what COCO.loadRes raises on an empty results file
        from .coco_eval import KeypointEval, summarize
        if self.coco is None:
            raise ValueError("result() needs the ground-truth annotations: pass json_path (or a parsed COCO dict)")
        if self._evaluator is None:
            self._evaluator = KeypointEval(self.coco)
        if self._arrays and sum(len(a[1]) for a in self._arrays) == len(self.result_list):
            # nobody edited result_list: evaluate the arrays it was built from (skips ~85 ms of dict walking per 35 k rows)
            kp = np.concatenate([a[0] for a in self._arrays])
            out = self._evaluator.evaluate_arrays(kp, np.concatenate([a[1] for a in self._arrays]),
                                                  sum((a[2] for a in self._arrays), []), sum((a[3] for a in self._arrays), []))
        else:
            out = self._evaluator.evaluate(self.result_list)
        self.stats = summarize(out['precision'], out['recall'], verbose=True)
        return self.stats[1]


def _load_coco(json_path):
    """The ground-truth annotations (`self.coco`): parsed here, no pycocotools needed."""
    if json_path is None:            # rows only (tests, ranks that never call result())
        return None
    from .coco_eval import CocoKeypointsGT
    return CocoKeypointsGT(json_path)
