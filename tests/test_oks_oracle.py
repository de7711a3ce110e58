"""Known-answer tests for the OKS / AP oracle (oracle/oks_oracle.py) -- CPU only.

pycocotools is the third-party evaluator the reference calls (utils/sbp_utils.py:175-189); it is absent here, so the
oracle is anchored on cases worked by hand from the published COCO keypoint-evaluation rules, and compared with
pycocotools itself whenever that package can be imported (skipped otherwise: PARITY UNPINNED)."""
import json
import math

import numpy as np
import pytest

from oracle import oks_oracle as oo
from oks_cases import make_dataset, person


def test_numpy_sum_order_is_what_the_kernel_restates():
    """np.sum over <= 17 doubles = 8 running sums + fixed tree + tail (csrc/oks_kernels.cuh numpy_sum)."""
    def restated(a):
        n = len(a)
        if n < 8:
            r = 0.0
            for v in a:
                r += v
            return r
        r = [float(v) for v in a[:8]]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] += a[i + j]
            i += 8
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
        for v in a[i:]:
            res += v
        return res
    rng = np.random.default_rng(0)
    for n in range(1, 33):
        for _ in range(50):
            a = np.exp(-rng.uniform(0, 5, n))
            assert np.sum(a) == restated(a), n


def test_perfect_detections_score_one():
    """Same idea as the reference's own check of the evaluation call (test_coco_keypoints_map.py:25-66: the ground truth
    submitted as results must score 1)."""
    gts = [person(1, 10, cx=100, cy=100, size=80), person(2, 11, cx=300, cy=200, size=120)]
    dts = [{'image_id': g['image_id'], 'category_id': 1, 'keypoints': list(g['keypoints']), 'score': 0.9} for g in gts]
    out = oo.evaluate(gts, dts)
    for (i, c), m in out['ious'].items():
        assert m.shape == (1, 1) and m[0, 0] == 1.0
    assert out['stats'][0] == pytest.approx(1.0) and out['stats'][1] == pytest.approx(1.0) and out['stats'][5] == pytest.approx(1.0)


def test_fixed_shift_closed_form():
    """every labelled joint displaced by (3, 4): OKS = mean_k exp(-25 / (2 (2 sigma_k)^2 area))."""
    g = person(1, 10, cx=100, cy=100, size=80, visible=[2] * 9 + [0] * 8)
    kp = np.array(g['keypoints'], dtype=np.float64).reshape(17, 3)
    kp[:, 0] += 3
    kp[:, 1] += 4
    d = {'image_id': 1, 'category_id': 1, 'keypoints': kp.reshape(-1).tolist(), 'score': 1.0}
    want = np.mean([math.exp(-25.0 / (2 * s) ** 2 / g['area'] / 2) for s in oo.COCO_SIGMAS[:9]])
    got = oo.compute_oks([g], [d])[0, 0]
    assert got == pytest.approx(want, rel=1e-14)


def test_unlabelled_gt_uses_doubled_box():
    g = person(1, 10, cx=100, cy=100, size=40, visible=[0] * 17)     # bbox = [80, 80, 40, 40] -> doubled box [40, 160]^2
    assert g['num_keypoints'] == 0
    inside = {'image_id': 1, 'category_id': 1, 'keypoints': [100, 100, 1] * 17, 'score': 1.0}
    outside = {'image_id': 1, 'category_id': 1, 'keypoints': [170, 100, 1] * 17, 'score': 1.0}     # 10 px right of the box
    m = oo.compute_oks([g], [inside, outside])
    assert m[0, 0] == 1.0
    want = np.mean([math.exp(-100.0 / (2 * s) ** 2 / g['area'] / 2) for s in oo.COCO_SIGMAS])
    assert m[1, 0] == pytest.approx(want, rel=1e-14)


def test_hand_built_pr_curve():
    """Three images, one person each; detections (score, hit): (.9, TP) (.8, FP) (.7, TP) (.6, TP at OKS .5 only).
    At OKS 0.5: precision envelope over recall = 1, 1/3: 1.0 ; 2/3: .75 ; 1: .75 -> AP50 = (34*1 + 67*.75) / 101."""
    gts = [person(i, 10 + i, cx=100, cy=100, size=100) for i in (1, 2, 3)]

    def det(img, score, shift):
        kp = np.array(gts[img - 1]['keypoints'], dtype=np.float64).reshape(17, 3)
        kp[:, 0] += shift
        return {'image_id': img, 'category_id': 1, 'keypoints': kp.reshape(-1).tolist(), 'score': score}
    far = 500.0
    # the last detection is shifted so that its OKS lands between 0.5 and 0.95: a hit at the low thresholds only
    dts = [det(1, .9, 0.0), det(1, .8, far), det(2, .7, 0.0), det(3, .6, 8.0)]
    out = oo.evaluate(gts, dts)
    oks3 = out['ious'][(3, 1)][0, 0]
    assert 0.5 < oks3 < 0.95
    want50 = (34 * 1.0 + 67 * 0.75) / 101
    assert out['stats'][1] == pytest.approx(want50, rel=1e-12)
    assert out['stats'][6] == pytest.approx(1.0)                    # AR50: all three found
    # at thresholds above oks3 the last detection is a false positive: recall 2/3, AP = 34*1 + 33*.75 (rc 2/3 reached by det 3 of 4)
    t_hi = int(np.searchsorted(oo.IOU_THRS, oks3, side='left'))
    p_hi = out['precision'][t_hi, :, 0, 0, 0]
    assert np.allclose(p_hi[:34], 1.0, rtol=1e-15) and np.allclose(p_hi[34:67], 2 / 3, rtol=1e-15) and p_hi[67:].tolist() == [0.0] * 34


def test_crowd_ignore_and_area_rules():
    small = person(1, 1, cx=50, cy=50, size=20)                       # area 400 < 32^2: ignored for medium / large
    crowd = person(1, 2, cx=200, cy=200, size=100, iscrowd=1)
    gts = [small, crowd]
    d_small = {'image_id': 1, 'category_id': 1, 'keypoints': list(small['keypoints']), 'score': .9}
    d_c1 = {'image_id': 1, 'category_id': 1, 'keypoints': list(crowd['keypoints']), 'score': .8}
    d_c2 = {'image_id': 1, 'category_id': 1, 'keypoints': list(crowd['keypoints']), 'score': .7}
    G, D = oo.prepare(gts, oo.load_res([d_small, d_c1, d_c2]), [1], [1])
    ious = oo.compute_oks(G[1, 1], D[1, 1])
    e = oo.evaluate_img(G[1, 1], D[1, 1], ious, oo.AREA_RNG[0], 20)
    assert e['gtIgnore'].tolist() == [0, 1]
    assert e['dtMatches'][0].tolist() == [1, 2, 2]                    # the crowd GT is matched twice
    assert e['dtIgnore'][0].tolist() == [False, True, True]
    e_med = oo.evaluate_img(G[1, 1], D[1, 1], ious, oo.AREA_RNG[1], 20)
    assert e_med['gtIgnore'].tolist() == [1, 1]
    assert e_med['dtIgnore'][0].all()                                  # matched to ignored GTs -> not counted
    out = oo.evaluate(gts, [d_small, d_c1, d_c2])
    assert out['stats'][1] == pytest.approx(1.0) and out['stats'][3] == -1     # no counted medium GT at all


def test_max_det_and_score_ties_are_stable():
    gts, dts = make_dataset(seed=3, n_images=4, max_people=3, dets_per_image=(24, 30))
    for d in dts[::3]:
        d['score'] = 0.5                                               # many exact ties
    out = oo.evaluate(gts, dts)
    for e in out['eval_imgs']:
        if e is not None:
            assert len(e['dtIds']) <= 20
            s = e['dtScores']
            assert all(s[i] >= s[i + 1] for i in range(len(s) - 1))
            ids = e['dtIds']
            assert all(ids[i] < ids[i + 1] for i in range(len(s) - 1) if s[i] == s[i + 1])


def test_against_pycocotools(tmp_path):
    cocoeval = pytest.importorskip("pycocotools.cocoeval", reason="pycocotools absent: OKS/AP parity unpinned")
    from pycocotools.coco import COCO
    gts, dts = make_dataset(seed=11, n_images=40, max_people=6)
    ann = tmp_path / "gt.json"
    ann.write_text(json.dumps({'images': [{'id': i} for i in sorted({g['image_id'] for g in gts})], 'annotations': gts,
                               'categories': [{'id': 1, 'name': 'person'}]}))
    res = tmp_path / "dt.json"
    res.write_text(json.dumps(dts))
    gt = COCO(str(ann))
    ev = cocoeval.COCOeval(gt, gt.loadRes(str(res)), "keypoints")
    ev.evaluate()
    ev.accumulate()
    ev.summarize()
    out = oo.evaluate(gts, dts)
    assert np.allclose(out['stats'], ev.stats, rtol=1e-12, atol=0)
    assert np.array_equal(out['precision'], ev.eval['precision'])
