#!/usr/bin/env python
"""Timing of the OKS / AP evaluation (pose_b200.coco_eval.KeypointEval) on synthetic sets.

    python tools/oks_bench.py [--out gpurun_out/oks_bench.json]

Two sets: "val2017-like" (5 000 images, 0..13 people, one detection per person plus false positives) and the bench
shape of config 3 (32 768 single-person images).  Reported: host grouping time, device time of the three kernels
(CUDA events), and the whole `evaluate()` call (wall clock, includes the uploads and the precision table download).
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pose_b200.coco_eval import KeypointEval  # noqa: E402

K = 17


def synth(n_images, max_people, seed=0):
    rng = np.random.default_rng(seed)
    gts, dts = [], []
    for i in range(n_images):
        for _ in range(int(rng.integers(1, max_people + 1)) if max_people > 1 else 1):
            size = float(rng.choice([40, 80, 150, 260]))
            x0, y0 = rng.uniform(0, 640 - size), rng.uniform(0, 480 - size)
            xy = np.stack([x0 + rng.uniform(0, size, K), y0 + rng.uniform(0, size, K)], -1).round()
            v = rng.choice([0, 1, 2], K, p=[.2, .2, .6])
            flat = []
            for k in range(K):
                flat += [float(xy[k, 0]), float(xy[k, 1]), int(v[k])] if v[k] else [0, 0, 0]
            gts.append({'id': len(gts) + 1, 'image_id': i + 1, 'category_id': 1, 'keypoints': flat, 'num_keypoints': int((v > 0).sum()),
                        'bbox': [x0, y0, size, size], 'area': size * size * .5, 'iscrowd': 0})
            det = (xy + rng.normal(0, rng.choice([.01, .05, .3]) * size, (K, 2))).astype(np.float32)
            dflat = []
            for k in range(K):
                dflat += [float(det[k, 0]), float(det[k, 1]), 1]
            dts.append({'image_id': i + 1, 'category_id': 1, 'keypoints': dflat, 'score': float(rng.uniform(.1, 1))})
    return {'images': [{'id': i + 1} for i in range(n_images)], 'annotations': gts, 'categories': [{'id': 1, 'name': 'person'}]}, dts


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "oks_bench.json"))
    args = ap.parse_args()
    res = {}
    for name, n, p in (("val2017-like (5000 images, <=13 people)", 5000, 13), ("config 3 shape (32768 single-person images)", 32768, 1)):
        data, dts = synth(n, p)
        t0 = time.perf_counter()
        ev = KeypointEval(data)
        t_gt = time.perf_counter() - t0
        ev.evaluate(dts)                      # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ev._group(dts)
        t_group = time.perf_counter() - t0
        wall, kern = [], []
        for _ in range(5):
            t0 = time.perf_counter()
            out = ev.evaluate(dts)
            torch.cuda.synchronize()
            wall.append(time.perf_counter() - t0)
            kern.append(out['kernel_ms'])
        res[name] = {"ground_truths": len(data['annotations']), "detections": len(dts), "pairs": int(out['pair_off'][-1]),
                     "gt_prepare_s": t_gt, "host_group_s": t_group, "evaluate_wall_s_best": min(wall), "kernels_ms_best": min(kern),
                     "AP": float(out['stats'][0]), "AP50": float(out['stats'][1])}
        print(name, json.dumps(res[name]), flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
