"""The reference's CPU pipeline for the SBP hot path, as bench.py times it (`cpu_baseline`, `--impl reference`).

TEST / MEASUREMENT INFRASTRUCTURE ONLY (see oracle/__init__.py): never imported by the product.

One pass over a sample of images is exactly what a reference training/validation step does on the CPU:

    render   per-sample SBPHeatmapGenerator loop      utils/sbp_utils.py:33-53   (dataset/sbp_coco_dataset.py:76)
    loss     SBPLoss forward + backward               models/loss/sbp_loss.py:20-66
    decode   per-sample, per-joint nms_sbp loop       utils/sbp_utils.py:56-118
    rows     back-projection + COCO rows              utils/sbp_utils.py:131-164

The reference runs render in DataLoader worker processes (configs/sbp_coco.yaml:40, `workers: 32`) and everything else
in the training process.  `ReferencePipeline(procs=P)` gives the host the same shape of parallelism and a little more:
the per-sample Python loops of render AND decode are spread over P forked worker processes (one torch thread each, as
DataLoader workers have), the loss runs in the parent on all torch threads while the workers decode.  `procs=1` is the
plain single-process form.
"""
import os

import numpy as np
import torch
import torch.multiprocessing as mp

from . import sbp_oracle as so

_W = {}          # per-worker state (set once by _init)


def _init(sample, cfg):
    torch.set_num_threads(1)          # what torch.utils.data workers do
    _W["sample"], _W["cfg"] = sample, cfg


def _render_range(rng):
    lo, hi = rng
    kp = _W["sample"][0]
    c = _W["cfg"]
    # a torch tensor travels back through shared memory (what DataLoader workers do with collated batches), not a pipe
    return torch.from_numpy(np.stack([so.sbp_render_loop(kp[b], c["H"], c["W"], c["sigma"]) for b in range(lo, hi)]))


def _decode_range(rng):
    lo, hi = rng
    logits = _W["sample"][1]
    c = _W["cfg"]
    return torch.stack([so.sbp_decode_loop(logits[b:b + 1], c["in_w"], c["thr"], True) for b in range(lo, hi)])


def _ranges(n, parts):
    parts = max(1, min(parts, n))
    edges = [n * i // parts for i in range(parts + 1)]
    return [(edges[i], edges[i + 1]) for i in range(parts) if edges[i + 1] > edges[i]]


class ReferencePipeline:
    """sample = (kp [n,K,2] f64, logits [n,K,H,W] f32, bbox [n,4] f64, image_id [n], category_id [n]) as from
    sbp_oracle.make_config1_inputs."""

    def __init__(self, sample, H, W, sigma, in_h, in_w, thr, procs=1):
        self.sample = sample
        self.cfg = dict(H=H, W=W, sigma=sigma, in_h=in_h, in_w=in_w, thr=thr)
        self.n = int(sample[1].size(0))
        self.procs = max(1, min(int(procs), self.n))
        self.pool = None
        if self.procs > 1:
            # fork: the workers inherit the sample (no pickling of the inputs per pass); rendered maps come back as
            # shared-memory tensors, like the batches of DataLoader workers do
            self.pool = mp.get_context("fork").Pool(self.procs, initializer=_init, initargs=(sample, self.cfg))
        else:
            _W["sample"], _W["cfg"] = sample, self.cfg

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()
            self.pool = None

    def run_pass(self):
        """-> (loss, number of result rows)"""
        kp, logits, bbox, iid, cid = self.sample
        c = self.cfg
        rngs = _ranges(self.n, self.procs)
        if self.pool is not None:
            target = torch.cat(self.pool.map(_render_range, rngs))
            pending = self.pool.map_async(_decode_range, rngs)      # the workers decode while the parent runs the loss
        else:
            target = _render_range((0, self.n))
        x = logits.detach().clone().requires_grad_(True)
        loss = so.sbp_loss(x, target)
        loss.backward()
        joints = torch.cat(pending.get()) if self.pool is not None else _decode_range((0, self.n))
        rows = so.sbp_result_rows(so.sbp_backproject(joints, bbox, (c["in_h"], c["in_w"])), iid, cid)
        return float(loss.detach()), len(rows)


def usable_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1
