"""Synthetic COCO person-keypoint ground truths and detections for the OKS / AP tests (CPU and GPU share them)."""
import numpy as np

K = 17


def person(image_id, ann_id, cx, cy, size, visible=None, iscrowd=0, category_id=1, rng=None):
    """One ground-truth annotation: joints on a fixed pattern (or random, with `rng`) inside the box."""
    visible = [2] * K if visible is None else list(visible)
    x0, y0 = cx - size / 2.0, cy - size / 2.0
    kp = []
    for k in range(K):
        if rng is None:
            x, y = x0 + (0.1 + 0.8 * ((k * 7) % K) / (K - 1)) * size, y0 + (0.1 + 0.8 * k / (K - 1)) * size
        else:
            x, y = x0 + rng.uniform(0.05, 0.95) * size, y0 + rng.uniform(0.05, 0.95) * size
        kp += [float(np.round(x)), float(np.round(y)), visible[k]] if visible[k] > 0 else [0, 0, 0]
    return {'id': ann_id, 'image_id': image_id, 'category_id': category_id, 'keypoints': kp,
            'num_keypoints': int(sum(v > 0 for v in visible)), 'bbox': [x0, y0, float(size), float(size)],
            'area': float(size) * float(size) * 0.53, 'iscrowd': iscrowd}


def make_dataset(seed, n_images, max_people, dets_per_image=None, n_cats=1, fp32_coords=True):
    """-> (gts, dts).  Mix of: crowd and unlabelled ground truths, small / medium / large people, images without
    ground truths or without detections, near-perfect to useless detections, missing joints, tied scores."""
    rng = np.random.default_rng(seed)
    gts, dts = [], []
    ann_id = 1
    for i in range(n_images):
        img = 1000 + 7 * i
        people = []
        n_p = 0 if rng.random() < 0.1 else int(rng.integers(1, max_people + 1))
        for _ in range(n_p):
            cat = int(rng.integers(1, n_cats + 1))
            size = float(rng.choice([24, 40, 70, 110, 180, 260]))
            cx, cy = rng.uniform(size / 2, 640 - size / 2), rng.uniform(size / 2, 480 - size / 2)
            r = rng.random()
            vis = [0] * K if r < 0.08 else [int(v) for v in rng.choice([0, 1, 2], K, p=[0.2, 0.2, 0.6])]
            g = person(img, ann_id, cx, cy, size, vis, iscrowd=int(rng.random() < 0.07), category_id=cat, rng=rng)
            ann_id += 1
            gts.append(g)
            people.append(g)
        cand = []
        for g in people:
            if rng.random() < 0.85:
                kp = np.array(g['keypoints'], dtype=np.float64).reshape(K, 3)
                size = g['bbox'][2]
                noise = rng.choice([0.0, 0.01, 0.03, 0.08, 0.3]) * size
                xy = np.where(kp[:, 2:3] > 0, kp[:, :2], np.array([[g['bbox'][0] + size / 2, g['bbox'][1] + size / 2]]))
                xy = xy + rng.normal(0, 1, (K, 2)) * noise
                cand.append((xy, g['category_id']))
        n_fp = int(rng.integers(0, 3))
        for _ in range(n_fp):
            cand.append((rng.uniform(0, 480, (K, 2)), int(rng.integers(1, n_cats + 1))))
        if dets_per_image is not None:
            want = int(rng.integers(dets_per_image[0], dets_per_image[1] + 1))
            while len(cand) < want:
                base = cand[int(rng.integers(0, len(cand)))] if cand and rng.random() < 0.7 else (rng.uniform(0, 480, (K, 2)), 1)
                cand.append((base[0] + rng.normal(0, 4, (K, 2)), base[1]))
        if rng.random() < 0.08:
            cand = []
        for xy, cat in cand:
            miss = rng.random(K) < 0.1
            if fp32_coords:
                xy = xy.astype(np.float32).astype(np.float64)
            flat = []
            for k in range(K):
                flat += [0, 0, 0] if miss[k] else [float(xy[k, 0]), float(xy[k, 1]), 1]
            dts.append({'image_id': img, 'category_id': cat, 'keypoints': flat, 'score': float(np.round(rng.uniform(0.05, 1.0), 2))})
    return gts, dts


def as_coco_dict(gts, extra_images=()):
    ids = sorted({g['image_id'] for g in gts} | set(extra_images))
    cats = sorted({g['category_id'] for g in gts}) or [1]
    return {'images': [{'id': i} for i in ids], 'annotations': gts, 'categories': [{'id': c, 'name': f'c{c}'} for c in cats]}
