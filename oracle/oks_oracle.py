"""CPU oracle for keypoint OKS / AP (COCO keypoint evaluation).  TEST INFRASTRUCTURE ONLY.

Parity status: **PARITY UNPINNED**.  The reference does not implement this itself: `SBPmAPCOCO.result`
(utils/sbp_utils.py:166-189), `SPMmAPCOCO.result` (utils/spm_utils.py:325-351) and `SBPmAPPIS.result` call the
third-party `pycocotools` (`COCO.loadRes`, `COCOeval(gt, dt, "keypoints")`, `.evaluate() / .accumulate() /
.summarize()`, return `stats[1]` = AP at OKS 0.50).  pycocotools is not vendored in /root/reference, its version is
not pinned there (README.md lists only "pycocotools"), and it is not installed in this image, so nothing here could be
checked against it.  This file restates the PUBLISHED algorithm of pycocotools 2.0.x `cocoeval.py`
(`computeOks`, `evaluateImg`, `accumulate`, `_summarizeKps`) and `coco.py:loadRes` (keypoints branch) as plain
loops; the known-answer tests in tests/test_oks_oracle.py anchor it on cases worked by hand (perfect detections,
a fixed shift with its closed-form OKS, a hand-built precision/recall curve, crowd / ignore / area rules).  When
pycocotools is importable, tests/test_oks_oracle.py::test_against_pycocotools compares the two and pins it.

Data model: `gts` and `dts` are lists of COCO annotation dicts.
  gt: image_id, category_id, id, keypoints [3K] (x, y, v), num_keypoints, bbox [x, y, w, h], area, iscrowd, (ignore)
  dt: image_id, category_id, keypoints [3K] (x, y, v), score        (the reference's `result_list` rows)
"""
import numpy as np

COCO_SIGMAS = np.array([.26, .25, .25, .35, .35, .79, .79, .72, .72, .62, .62, 1.07, 1.07, .87, .87, .89, .89]) / 10.0
IOU_THRS = np.linspace(.5, 0.95, int(np.round((0.95 - .5) / .05)) + 1, endpoint=True)
REC_THRS = np.linspace(.0, 1.00, int(np.round((1.00 - .0) / .01)) + 1, endpoint=True)
AREA_RNG = [[0 ** 2, 1e5 ** 2], [32 ** 2, 96 ** 2], [96 ** 2, 1e5 ** 2]]      # all, medium, large (keypoints)
MAX_DETS = [20]


def load_res(dts):
    """coco.py:loadRes, keypoints branch: area / bbox from the min/max of ALL x, y entries (zeros included), id = 1.."""
    out = []
    for n, d in enumerate(dts):
        d = dict(d)
        s = d['keypoints']
        x, y = s[0::3], s[1::3]
        x0, x1, y0, y1 = np.min(x), np.max(x), np.min(y), np.max(y)
        d['area'] = (x1 - x0) * (y1 - y0)
        d['id'] = n + 1
        d['bbox'] = [x0, y0, x1 - x0, y1 - y0]
        out.append(d)
    return out


def prepare(gts, dts, img_ids, cat_ids):
    """cocoeval.py:_prepare -- group by (image, category); keypoint GTs with no labelled joint are ignored."""
    G, D = {}, {}
    imgs, cats = set(img_ids), set(cat_ids)
    for g in gts:
        if g['image_id'] not in imgs or g['category_id'] not in cats:
            continue
        g = dict(g)
        g['ignore'] = g['ignore'] if 'ignore' in g else 0
        g['ignore'] = 'iscrowd' in g and g['iscrowd']
        g['ignore'] = (g['num_keypoints'] == 0) or g['ignore']
        G.setdefault((g['image_id'], g['category_id']), []).append(g)
    for d in dts:
        if d['image_id'] not in imgs or d['category_id'] not in cats:
            continue
        D.setdefault((d['image_id'], d['category_id']), []).append(d)
    return G, D


def compute_oks(gts, dts, sigmas=COCO_SIGMAS, max_det=MAX_DETS[-1]):
    """cocoeval.py:computeOks for one (image, category): [D', G] with the detections sorted by -score (stable), top max_det."""
    order = np.argsort([-d['score'] for d in dts], kind='mergesort')
    dts = [dts[i] for i in order][:max_det]
    if len(gts) == 0 or len(dts) == 0:
        return np.zeros((0, 0))
    ious = np.zeros((len(dts), len(gts)))
    var = (np.asarray(sigmas) * 2) ** 2
    k = len(sigmas)
    for j, gt in enumerate(gts):
        g = np.array(gt['keypoints'])
        xg, yg, vg = g[0::3], g[1::3], g[2::3]
        k1 = np.count_nonzero(vg > 0)
        bb = gt['bbox']
        x0, x1 = bb[0] - bb[2], bb[0] + bb[2] * 2
        y0, y1 = bb[1] - bb[3], bb[1] + bb[3] * 2
        for i, dt in enumerate(dts):
            d = np.array(dt['keypoints'])
            xd, yd = d[0::3], d[1::3]
            if k1 > 0:
                dx, dy = xd - xg, yd - yg
            else:       # no labelled joint: distance to the doubled box
                z = np.zeros(k)
                dx = np.max((z, x0 - xd), axis=0) + np.max((z, xd - x1), axis=0)
                dy = np.max((z, y0 - yd), axis=0) + np.max((z, yd - y1), axis=0)
            e = (dx ** 2 + dy ** 2) / var / (gt['area'] + np.spacing(1)) / 2
            if k1 > 0:
                e = e[vg > 0]
            ious[i, j] = np.sum(np.exp(-e)) / e.shape[0]
    return ious


def evaluate_img(gts, dts, ious, a_rng, max_det, iou_thrs=IOU_THRS):
    """cocoeval.py:evaluateImg for one (image, category, area range).  `ious` from compute_oks (GT in input order)."""
    if len(gts) == 0 and len(dts) == 0:
        return None
    ig = [1 if (g['ignore'] or g['area'] < a_rng[0] or g['area'] > a_rng[1]) else 0 for g in gts]
    gtind = np.argsort(ig, kind='mergesort')
    gt = [gts[i] for i in gtind]
    gt_ig = np.array([ig[i] for i in gtind])
    dtind = np.argsort([-d['score'] for d in dts], kind='mergesort')
    dt = [dts[i] for i in dtind[:max_det]]
    crowd = [int(g['iscrowd']) for g in gt]
    ious = ious[:, gtind] if len(ious) > 0 else ious
    T, Gn, Dn = len(iou_thrs), len(gt), len(dt)
    gtm, dtm, dt_ig = np.zeros((T, Gn)), np.zeros((T, Dn)), np.zeros((T, Dn))
    if len(ious) != 0:
        for ti, t in enumerate(iou_thrs):
            for di, d in enumerate(dt):
                best = min([t, 1 - 1e-10])
                m = -1
                for gi in range(Gn):
                    if gtm[ti, gi] > 0 and not crowd[gi]:
                        continue
                    if m > -1 and gt_ig[m] == 0 and gt_ig[gi] == 1:
                        break
                    if ious[di, gi] < best:
                        continue
                    best = ious[di, gi]
                    m = gi
                if m == -1:
                    continue
                dt_ig[ti, di] = gt_ig[m]
                dtm[ti, di] = gt[m]['id']
                gtm[ti, m] = d['id']
    out_rng = np.array([d['area'] < a_rng[0] or d['area'] > a_rng[1] for d in dt]).reshape((1, Dn))
    dt_ig = np.logical_or(dt_ig, np.logical_and(dtm == 0, np.repeat(out_rng, T, 0)))
    return {'dtIds': [d['id'] for d in dt], 'gtIds': [g['id'] for g in gt], 'dtMatches': dtm, 'gtMatches': gtm,
            'dtScores': [d['score'] for d in dt], 'gtIgnore': gt_ig, 'dtIgnore': dt_ig}


def accumulate(eval_imgs, n_cats, n_areas, n_imgs, iou_thrs=IOU_THRS, rec_thrs=REC_THRS, max_dets=MAX_DETS):
    """cocoeval.py:accumulate.  eval_imgs is flat, index = (cat * n_areas + area) * n_imgs + img."""
    T, R, M = len(iou_thrs), len(rec_thrs), len(max_dets)
    precision = -np.ones((T, R, n_cats, n_areas, M))
    recall = -np.ones((T, n_cats, n_areas, M))
    for k in range(n_cats):
        for a in range(n_areas):
            for m, max_det in enumerate(max_dets):
                E = [eval_imgs[(k * n_areas + a) * n_imgs + i] for i in range(n_imgs)]
                E = [e for e in E if e is not None]
                if len(E) == 0:
                    continue
                scores = np.concatenate([e['dtScores'][0:max_det] for e in E])
                inds = np.argsort(-scores, kind='mergesort')
                dtm = np.concatenate([e['dtMatches'][:, 0:max_det] for e in E], axis=1)[:, inds]
                dt_ig = np.concatenate([e['dtIgnore'][:, 0:max_det] for e in E], axis=1)[:, inds]
                gt_ig = np.concatenate([e['gtIgnore'] for e in E])
                npig = np.count_nonzero(gt_ig == 0)
                if npig == 0:
                    continue
                tps = np.logical_and(dtm, np.logical_not(dt_ig))
                fps = np.logical_and(np.logical_not(dtm), np.logical_not(dt_ig))
                tp_sum = np.cumsum(tps, axis=1).astype(dtype=float)
                fp_sum = np.cumsum(fps, axis=1).astype(dtype=float)
                for t, (tp, fp) in enumerate(zip(tp_sum, fp_sum)):
                    tp, fp = np.array(tp), np.array(fp)
                    nd = len(tp)
                    rc = tp / npig
                    pr = tp / (fp + tp + np.spacing(1))
                    q = np.zeros((R,))
                    recall[t, k, a, m] = rc[-1] if nd else 0
                    pr = pr.tolist()
                    q = q.tolist()
                    for i in range(nd - 1, 0, -1):
                        if pr[i] > pr[i - 1]:
                            pr[i - 1] = pr[i]
                    where = np.searchsorted(rc, rec_thrs, side='left')
                    try:
                        for ri, pi in enumerate(where):
                            q[ri] = pr[pi]
                    except IndexError:
                        pass
                    precision[t, :, k, a, m] = np.array(q)
    return precision, recall


def summarize(precision, recall, iou_thrs=IOU_THRS):
    """cocoeval.py:_summarizeKps -> stats[10]: AP, AP50, AP75, APm, APl, AR, AR50, AR75, ARm, ARl (maxDets 20)."""
    def one(ap, thr=None, area=0):
        s = precision if ap else recall
        if thr is not None:
            s = s[np.where(thr == iou_thrs)[0]]
        s = s[:, :, :, area, 0] if ap else s[:, :, area, 0]
        return -1 if len(s[s > -1]) == 0 else np.mean(s[s > -1])
    return np.array([one(1), one(1, .5), one(1, .75), one(1, area=1), one(1, area=2),
                     one(0), one(0, .5), one(0, .75), one(0, area=1), one(0, area=2)], dtype=np.float64)


def evaluate(gts, dts, img_ids=None, cat_ids=None, sigmas=COCO_SIGMAS):
    """The whole COCOeval("keypoints") run -> dict(stats, precision, recall, ious, eval_imgs).  stats[1] is what the
    reference's `result()` returns (utils/sbp_utils.py:189)."""
    img_ids = sorted(set(g['image_id'] for g in gts)) if img_ids is None else sorted(set(img_ids))
    cat_ids = sorted(set(g['category_id'] for g in gts)) if cat_ids is None else sorted(set(cat_ids))
    G, D = prepare(gts, load_res(dts), img_ids, cat_ids)
    ious = {(i, c): compute_oks(G.get((i, c), []), D.get((i, c), []), sigmas) for i in img_ids for c in cat_ids}
    eval_imgs = [evaluate_img(G.get((i, c), []), D.get((i, c), []), ious[i, c], a, MAX_DETS[-1])
                 for c in cat_ids for a in AREA_RNG for i in img_ids]
    precision, recall = accumulate(eval_imgs, len(cat_ids), len(AREA_RNG), len(img_ids))
    return {'stats': summarize(precision, recall), 'precision': precision, 'recall': recall, 'ious': ious,
            'eval_imgs': eval_imgs, 'img_ids': img_ids, 'cat_ids': cat_ids}
