"""COCO keypoint evaluation (OKS matching + AP/AR) on the device -- what `SBPmAPCOCO.result`
(utils/sbp_utils.py:166-189), `SPMmAPCOCO.result` (utils/spm_utils.py:325-351) and `SBPmAPPIS.result` obtain from
pycocotools in the reference: `COCO(json)`, `coco.loadRes(results)`, `COCOeval(coco, res, "keypoints")`,
`.evaluate()`, `.accumulate()`, `.summarize()`, `.stats`.

pycocotools is a third-party dependency of the reference (unpinned, not vendored, absent from this image), so this
module restates its published algorithm (PARITY UNPINNED): the host groups ground truths and
detections by (category, image) and the three kernels behind `pose_oks_matrix`, `pose_oks_match` and
`pose_ap_accumulate` do the arithmetic in fp64.  No CPU fallback: without the CUDA library `evaluate` raises.
"""
import json

import numpy as np
import torch

from ._cabi import check, lib, ptr, stream_ptr

COCO_SIGMAS = np.array([.26, .25, .25, .35, .35, .79, .79, .72, .72, .62, .62, 1.07, 1.07, .87, .87, .89, .89]) / 10.0
IOU_THRS = np.linspace(.5, 0.95, int(np.round((0.95 - .5) / .05)) + 1, endpoint=True)
REC_THRS = np.linspace(.0, 1.00, int(np.round((1.00 - .0) / .01)) + 1, endpoint=True)
AREA_RNG = np.array([[0 ** 2, 1e5 ** 2], [32 ** 2, 96 ** 2], [96 ** 2, 1e5 ** 2]], dtype=np.float64)
AREA_LBL = ['all', 'medium', 'large']
MAX_DET = 20


class CocoKeypointsGT:
    """The part of `pycocotools.coco.COCO` the metric classes use: the parsed annotation file, `getImgIds`,
    `getCatIds`.  Accepts a path to a COCO person-keypoints JSON or the already parsed dict."""

    def __init__(self, annotation_file):
        if isinstance(annotation_file, dict):
            self.dataset = annotation_file
        else:
            with open(annotation_file, 'r') as f:
                self.dataset = json.load(f)
        assert isinstance(self.dataset, dict), 'annotation file format {} not supported'.format(type(self.dataset))
        self.anns = {a['id']: a for a in self.dataset.get('annotations', [])}
        self.imgs = {im['id']: im for im in self.dataset.get('images', [])}
        self.cats = {c['id']: c for c in self.dataset.get('categories', [])}

    def getImgIds(self):
        return list(self.imgs.keys())

    def getCatIds(self):
        return [c['id'] for c in self.dataset.get('categories', [])]


def _dev(a, device, dtype):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).to(device)


class KeypointEval:
    """COCOeval(gt, dt, "keypoints") with params.imgIds / catIds = all of the ground truth's, as the reference sets them.

    The ground truth is grouped and uploaded once; `evaluate(results)` may be called once per validation epoch.
    """

    def __init__(self, gt, device=None, sigmas=COCO_SIGMAS, max_det=MAX_DET):
        self.gt = gt if isinstance(gt, CocoKeypointsGT) else CocoKeypointsGT(gt)
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.sigmas = np.asarray(sigmas, dtype=np.float64)
        self.K = len(self.sigmas)
        self.max_det = int(max_det)
        self.img_ids = sorted(set(self.gt.getImgIds()))
        self.cat_ids = sorted(set(self.gt.getCatIds()))
        self._img_index = {v: i for i, v in enumerate(self.img_ids)}
        self._cat_index = {v: i for i, v in enumerate(self.cat_ids)}
        I, C = len(self.img_ids), len(self.cat_ids)
        self.Q = I * C
        anns = [a for a in self.gt.dataset.get('annotations', [])
                if a['image_id'] in self._img_index and a['category_id'] in self._cat_index]
        q = np.array([self._cat_index[a['category_id']] * I + self._img_index[a['image_id']] for a in anns], dtype=np.int64)
        order = np.argsort(q, kind='stable')
        anns = [anns[i] for i in order]
        G = len(anns)
        self.G = G
        self.gt_ids = [a['id'] for a in anns]
        kp = np.zeros((G, self.K, 3), dtype=np.float64)
        flags = np.zeros(G, dtype=np.uint8)
        for n, a in enumerate(anns):
            k = np.asarray(a['keypoints'], dtype=np.float64)
            assert k.size == 3 * self.K, f"ground truth {a['id']}: {k.size // 3} keypoints, the OKS sigmas cover {self.K}"
            kp[n] = k.reshape(self.K, 3)
            crowd = bool(a.get('iscrowd', 0))
            flags[n] = (1 if (a['num_keypoints'] == 0 or crowd) else 0) | (2 if crowd else 0)
        self.gt_count = np.bincount(q, minlength=self.Q).astype(np.int64) if G else np.zeros(self.Q, dtype=np.int64)
        self.gt_off = np.concatenate([[0], np.cumsum(self.gt_count)]).astype(np.int32)
        d = self.device
        self._gt_kp = _dev(kp, d, np.float64)
        self._gt_bbox = _dev(np.array([a['bbox'] for a in anns], dtype=np.float64).reshape(G, 4), d, np.float64)
        self._gt_area = _dev(np.array([a['area'] for a in anns], dtype=np.float64), d, np.float64)
        self._gt_flags = _dev(flags, d, np.uint8)
        self._gt_off = _dev(self.gt_off, d, np.int32)
        self._cat_gt_off = _dev(self.gt_off[::I] if I else np.zeros(C + 1), d, np.int32)
        self._sigmas = _dev(self.sigmas, d, np.float64)
        self._area_rng = _dev(AREA_RNG, d, np.float64)
        self._iou_thrs = _dev(IOU_THRS, d, np.float64)
        self._rec_thrs = _dev(REC_THRS, d, np.float64)

    # -- host: group the detections -------------------------------------------------------------------------------
    def _group(self, results):
        """COCO-result dicts -> arrays (this walk over ~10^4..10^5 python dicts is most of an `evaluate()` call: 84 of 85 ms for
        val2017; callers that still hold the arrays the dicts were made from pass them to `evaluate_arrays` instead)."""
        if not isinstance(results, list):
            raise TypeError('results in not an array of objects')
        kp = np.array([r['keypoints'] for r in results], dtype=np.float64)
        assert kp.size == len(results) * 3 * self.K, f"results must carry {self.K} keypoints (x, y, v) each, as many as the OKS sigmas"
        return self._group_arrays(kp.reshape(len(results), 3 * self.K), np.array([r['score'] for r in results], dtype=np.float64),
                                  [r['image_id'] for r in results], [r['category_id'] for r in results])

    def _group_arrays(self, kp, score, image_ids, category_ids):
        """kp [N, 3K] (x, y, v), score [N], ids [N] -> detections sorted by (group, -score), at most max_det per group."""
        I = len(self.img_ids)
        # COCO.loadRes: every result must belong to an image of the ground truth
        img = np.array([self._img_index.get(int(i), -1) for i in image_ids], dtype=np.int64)
        assert np.all(img >= 0), 'Results do not correspond to current coco set'
        cat = np.array([self._cat_index.get(int(c), -1) for c in category_ids], dtype=np.int64)
        keep = np.nonzero(cat >= 0)[0]
        kp = np.asarray(kp, dtype=np.float64).reshape(-1, 3 * self.K)[keep]
        score = np.asarray(score, dtype=np.float64)[keep]
        q = cat[keep] * I + img[keep]
        order = np.lexsort((-score, q))                      # by group, then descending score; ties keep input order
        q, score, kp = q[order], score[order], kp[order]
        src = keep[order]
        start = np.concatenate([[0], np.cumsum(np.bincount(q, minlength=self.Q))])[:-1] if len(q) else np.zeros(self.Q, dtype=np.int64)
        rank = np.arange(len(q)) - start[q] if len(q) else np.zeros(0, dtype=np.int64)
        top = rank < self.max_det                             # COCOeval keeps the max_det best of every group
        return q[top], score[top], kp[top].reshape(-1, self.K, 3), src[top]

    # -- device ---------------------------------------------------------------------------------------------------
    def evaluate(self, results):
        """-> dict(stats [10], precision [T,R,C,A], recall [T,C,A], oks (flat), det_* ...).  stats[1] is AP at OKS 0.50."""
        return self._evaluate_grouped(*self._group(results))

    def evaluate_arrays(self, kp, score, image_ids, category_ids):
        """Same from arrays (kp [N, 3K] or [N, K, 3], score [N], ids [N]): no python dicts are walked."""
        return self._evaluate_grouped(*self._group_arrays(np.asarray(kp, dtype=np.float64).reshape(len(score), -1), score, image_ids, category_ids))

    def _evaluate_grouped(self, q, score, kp, src):
        d = self.device
        I, C, Q, G, K = len(self.img_ids), len(self.cat_ids), self.Q, self.G, self.K
        A, T, R = len(AREA_RNG), len(IOU_THRS), len(REC_THRS)
        D = len(q)
        det_count = np.bincount(q, minlength=Q).astype(np.int64) if D else np.zeros(Q, dtype=np.int64)
        det_off = np.concatenate([[0], np.cumsum(det_count)]).astype(np.int32)
        pair_off = np.concatenate([[0], np.cumsum(det_count * self.gt_count)]).astype(np.int64)
        n_pairs = int(pair_off[-1])
        cat_of = q // I if I else q
        order = np.lexsort((-score, cat_of)).astype(np.int64)          # by category, then descending score, stable

        det_kp = _dev(kp, d, np.float64)
        t_det_off, t_pair_off = _dev(det_off, d, np.int32), _dev(pair_off, d, np.int64)
        t_cat_det_off = _dev(det_off[::I] if I else np.zeros(C + 1), d, np.int32)
        t_order = _dev(order, d, np.int64)
        oks = torch.empty(max(n_pairs, 1), dtype=torch.float64, device=d)
        det_area = torch.empty(max(D, 1), dtype=torch.float64, device=d)
        dt_match = torch.empty((A, T, max(D, 1)), dtype=torch.int32, device=d)
        dt_ignore = torch.empty((A, T, max(D, 1)), dtype=torch.uint8, device=d)
        gt_ignore = torch.empty((A, max(G, 1)), dtype=torch.uint8, device=d)
        precision = torch.empty((T, R, C, A), dtype=torch.float64, device=d)
        recall = torch.empty((T, C, A), dtype=torch.float64, device=d)
        L = lib()
        ws_m = int(L.pose_oks_match_workspace_bytes(A, T, G))
        ws_a = int(L.pose_ap_accumulate_workspace_bytes(A, T, D))
        ws = torch.empty(max(ws_m, ws_a, 8), dtype=torch.uint8, device=d)
        with torch.cuda.device(d):
            st = stream_ptr(d)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            check(L.pose_oks_matrix(ptr(det_kp), ptr(self._gt_kp), ptr(self._gt_bbox), ptr(self._gt_area), ptr(t_det_off),
                                    ptr(self._gt_off), ptr(t_pair_off), ptr(self._sigmas), ptr(oks), ptr(det_area),
                                    Q, D, G, n_pairs, K, st), "pose_oks_matrix")
            check(L.pose_oks_match(ptr(oks), ptr(t_pair_off), ptr(t_det_off), ptr(self._gt_off), ptr(det_area), ptr(self._gt_area),
                                   ptr(self._gt_flags), ptr(self._area_rng), ptr(self._iou_thrs), Q, A, T, D, G,
                                   ptr(dt_match), ptr(dt_ignore), ptr(gt_ignore), ptr(ws), ws.numel(), st), "pose_oks_match")
            check(L.pose_ap_accumulate(ptr(t_order), ptr(dt_match), ptr(dt_ignore), ptr(gt_ignore), ptr(t_cat_det_off),
                                       ptr(self._cat_gt_off), ptr(self._rec_thrs), C, A, T, R, D, G, ptr(precision), ptr(recall),
                                       ptr(ws), ws.numel(), st), "pose_ap_accumulate")
            ev1.record()
        prec, rec = precision.cpu().numpy()[..., None], recall.cpu().numpy()[..., None]     # trailing max_det axis, as COCOeval
        return {'stats': summarize(prec, rec), 'precision': prec, 'recall': rec, 'kernel_ms': ev0.elapsed_time(ev1),
                'oks': oks[:n_pairs], 'pair_off': pair_off, 'det_off': det_off, 'det_src': src, 'det_area': det_area[:D],
                'dt_match': dt_match[:, :, :D], 'dt_ignore': dt_ignore[:, :, :D], 'gt_ignore': gt_ignore[:, :G]}


def summarize(precision, recall, verbose=False):
    """COCOeval._summarizeKps: [AP, AP50, AP75, APm, APl, AR, AR50, AR75, ARm, ARl] at max_det 20."""
    def one(ap, thr=None, area=0):
        s = precision if ap else recall
        if thr is not None:
            s = s[np.where(thr == IOU_THRS)[0]]
        s = s[:, :, :, area, 0] if ap else s[:, :, area, 0]
        v = -1 if len(s[s > -1]) == 0 else np.mean(s[s > -1])
        if verbose:
            rng = '{:0.2f}:{:0.2f}'.format(IOU_THRS[0], IOU_THRS[-1]) if thr is None else '{:0.2f}'.format(thr)
            print(' {:<18} {} @[ IoU={:<9} | area={:>6s} | maxDets={:>3d} ] = {:0.3f}'.format(
                'Average Precision' if ap else 'Average Recall', '(AP)' if ap else '(AR)', rng, AREA_LBL[area], MAX_DET, v))
        return v
    return np.array([one(1), one(1, .5), one(1, .75), one(1, area=1), one(1, area=2),
                     one(0), one(0, .5), one(0, .75), one(0, area=1), one(0, area=2)], dtype=np.float64)
