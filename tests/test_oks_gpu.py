"""GPU parity of the OKS / AP kernels (pose_oks_matrix, pose_oks_match, pose_ap_accumulate, through the C ABI) against
oracle/oks_oracle.py.  OKS values within 1e-13 relative (fp64; only exp() may differ by an ulp), matches, ignore flags
bit-exact, precision / recall tables bit-exact given equal matches, the ten summary numbers within 1e-12."""
import numpy as np
import pytest
import torch

from oks_cases import as_coco_dict, make_dataset, person

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pb():
    import __graft_entry__ as ge
    ge.build()
    import pose_b200
    return pose_b200


def _compare(pb, gts, dts, extra_images=()):
    from oracle import oks_oracle as oo
    from pose_b200.coco_eval import KeypointEval
    data = as_coco_dict(gts, list(extra_images) + [d['image_id'] for d in dts])      # images without ground truth exist too
    ev = KeypointEval(data)
    n0 = pb.launch_count()
    got = ev.evaluate(dts)
    assert pb.launch_count() - n0 >= 1
    want = oo.evaluate(gts, dts, img_ids=[im['id'] for im in data['images']], cat_ids=[c['id'] for c in data['categories']])
    I, C, A = len(ev.img_ids), len(ev.cat_ids), len(oo.AREA_RNG)
    assert ev.img_ids == want['img_ids'] and ev.cat_ids == want['cat_ids']
    oks = got['oks'].cpu().numpy()
    dtm, dti, gti = got['dt_match'].cpu().numpy(), got['dt_ignore'].cpu().numpy(), got['gt_ignore'].cpu().numpy()
    gt_ids = np.array(ev.gt_ids + [0])
    n_pairs = 0
    for c, cat in enumerate(ev.cat_ids):
        for i, img in enumerate(ev.img_ids):
            q = c * I + i
            w = want['ious'][img, cat]
            lo, hi = got['pair_off'][q], got['pair_off'][q + 1]
            assert hi - lo == w.size, (img, cat)
            n_pairs += w.size
            if w.size:
                assert np.allclose(oks[lo:hi], w.reshape(-1), rtol=1e-13, atol=0), np.abs(oks[lo:hi] - w.reshape(-1)).max()
            d0, d1 = got['det_off'][q], got['det_off'][q + 1]
            g0, g1 = ev.gt_off[q], ev.gt_off[q + 1]
            for a in range(A):
                e = want['eval_imgs'][(c * A + a) * I + i]
                if e is None:
                    assert d0 == d1 and g0 == g1
                    continue
                assert d1 - d0 == len(e['dtIds']) and g1 - g0 == len(e['gtIds'])
                # detections in the same order (score-sorted, stable): matched ground-truth ids, ignore flags
                m = dtm[a, :, d0:d1]
                assert np.array_equal(np.where(m > 0, gt_ids[m - 1], 0), e['dtMatches'].astype(np.int64)), (img, cat, a)
                assert np.array_equal(dti[a, :, d0:d1].astype(bool), e['dtIgnore'].astype(bool)), (img, cat, a)
                ig = dict(zip(e['gtIds'], e['gtIgnore'].tolist()))
                assert [ig[g] for g in ev.gt_ids[g0:g1]] == gti[a, g0:g1].tolist()
    assert n_pairs == oks.size
    assert np.array_equal(got['precision'], want['precision'])
    assert np.array_equal(got['recall'], want['recall'])
    assert np.allclose(got['stats'], want['stats'], rtol=1e-12, atol=0)
    return got, want


@pytest.mark.parametrize("seed,n_images,max_people,n_cats,dpi", [
    (1, 30, 5, 1, None), (2, 60, 8, 1, None), (3, 12, 4, 1, (24, 40)), (4, 40, 6, 3, None), (5, 1, 1, 1, None), (6, 300, 10, 2, None)])
def test_matches_oracle_on_mixed_sets(pb, seed, n_images, max_people, n_cats, dpi):
    gts, dts = make_dataset(seed, n_images, max_people, dets_per_image=dpi, n_cats=n_cats)
    if not gts:
        gts = [person(1000, 1, 100, 100, 80)]
    got, want = _compare(pb, gts, dts, extra_images=[999999])
    assert -1 <= got['stats'][1] <= 1


def test_ground_truth_as_results_like_the_reference_script(pb):
    """The reference's own check of its evaluation call (test_coco_keypoints_map.py:25-66): every ground-truth annotation
    is submitted as a result with score 0.9.  With labelled, non-crowd people only, every number is 1 (medium / large
    where such people exist); with crowd / unlabelled annotations mixed in, the oracle decides."""
    rng = np.random.default_rng(5)
    gts = []
    for i in range(200):
        for _ in range(int(rng.integers(1, 5))):
            size = float(rng.choice([50, 90, 150, 240]))
            gts.append(person(10 + i, len(gts) + 1, rng.uniform(size / 2, 600), rng.uniform(size / 2, 400), size,
                              [int(v) for v in rng.choice([1, 2], 17)], rng=rng))
    dts = [{'image_id': g['image_id'], 'category_id': g['category_id'], 'keypoints': g['keypoints'], 'score': float(0.9)} for g in gts]
    got, _ = _compare(pb, gts, dts)
    assert np.allclose(got['stats'], 1.0, rtol=1e-12)
    gts2, _ = make_dataset(9, 80, 6)
    dts2 = [{'image_id': g['image_id'], 'category_id': g['category_id'], 'keypoints': g['keypoints'], 'score': float(0.9)} for g in gts2]
    _compare(pb, gts2, dts2)


def test_score_ties_and_detection_src_order(pb):
    gts, dts = make_dataset(7, 25, 6, dets_per_image=(5, 25))
    for d in dts[::2]:
        d['score'] = 0.5
    got, want = _compare(pb, gts, dts)
    # within a group equal scores keep the order of the results list (COCO ids are positions + 1)
    off, src = got['det_off'], got['det_src']
    sc = np.array([d['score'] for d in dts])
    for q in range(len(off) - 1):
        s, r = sc[src[off[q]:off[q + 1]]], src[off[q]:off[q + 1]]
        assert all(s[i] > s[i + 1] or (s[i] == s[i + 1] and r[i] < r[i + 1]) for i in range(len(s) - 1))


def test_no_detections_for_the_only_category_and_foreign_category(pb):
    gts = [person(1, 1, 100, 100, 80), person(2, 2, 200, 200, 120)]
    dts = [{'image_id': 1, 'category_id': 77, 'keypoints': list(gts[0]['keypoints']), 'score': 1.0}]     # dropped: unknown category
    got, want = _compare(pb, gts, dts)
    assert got['stats'][1] == 0.0 and got['stats'][6] == 0.0


def test_result_for_unknown_image_asserts_like_loadres(pb):
    from pose_b200.coco_eval import KeypointEval
    ev = KeypointEval(as_coco_dict([person(1, 1, 100, 100, 80)]))
    with pytest.raises(AssertionError):
        ev.evaluate([{'image_id': 5, 'category_id': 1, 'keypoints': [0] * 51, 'score': 1.0}])


def test_perfect_detections_at_bench_scale(pb):
    """32 768 single-person images, detections = ground truth: every OKS is exactly 1 and AP = AR = 1 (size-independent property)."""
    n = 32768
    gts = [person(i + 1, i + 1, 100 + (i % 37), 120 + (i % 11), 60 + (i % 90)) for i in range(n)]
    dts = [{'image_id': g['image_id'], 'category_id': 1, 'keypoints': list(g['keypoints']), 'score': 1.0 - (i % 1000) * 1e-4}
           for i, g in enumerate(gts)]
    from pose_b200.coco_eval import KeypointEval
    out = KeypointEval(as_coco_dict(gts)).evaluate(dts)
    assert bool((out['oks'] == 1.0).all()) and out['oks'].numel() == n
    assert np.allclose(out['stats'][[0, 1, 2, 5, 6, 7]], 1.0, rtol=1e-12)


def test_metric_class_result_end_to_end(pb, tmp_path, monkeypatch):
    """SBPmAPCOCO(json).update_state(...).result(): logits rendered from the ground-truth joints decode back to the
    joints' pixels (x4 grid), so AP50 is 1; the same rows through the oracle give the same ten numbers."""
    import json
    from oracle import oks_oracle as oo
    monkeypatch.chdir(tmp_path)
    rng = np.random.default_rng(0)
    B, K, H, W = 24, 17, 64, 48
    kp = np.stack([rng.uniform(4, W - 4, (B, K)), rng.uniform(4, H - 4, (B, K))], axis=-1)
    bbox = np.stack([rng.uniform(0, 300, B), rng.uniform(0, 300, B), rng.uniform(120, 300, B), rng.uniform(160, 400, B)], axis=-1)
    gts = []
    for b in range(B):
        xs = np.floor(kp[b, :, 0]) * 4 * (bbox[b, 2] / 192) + bbox[b, 0]
        ys = np.floor(kp[b, :, 1]) * 4 * (bbox[b, 3] / 256) + bbox[b, 1]
        flat = []
        for k in range(K):
            flat += [float(xs[k]), float(ys[k]), 2]
        gts.append({'id': b + 1, 'image_id': 500 + b, 'category_id': 1, 'keypoints': flat, 'num_keypoints': K,
                    'bbox': bbox[b].tolist(), 'area': float(bbox[b, 2] * bbox[b, 3]), 'iscrowd': 0})
    ann = tmp_path / "person_keypoints.json"
    ann.write_text(json.dumps(as_coco_dict(gts)))
    dev = torch.device("cuda", 0)
    target = torch.from_numpy(np.stack([pb.SBPHeatmapGenerator([H, W], K, 2)(kp[b]) for b in range(B)])).to(dev)
    logits = torch.logit(target.clamp(1e-4, 1 - 1e-4))
    m = pb.SBPmAPCOCO(str(ann), [256, 192], 0.25)
    assert sorted(m.coco.getImgIds()) == [500 + b for b in range(B)] and m.coco.getCatIds() == [1]
    m.update_state({'bbox': torch.from_numpy(bbox), 'image_id': torch.arange(500, 500 + B), 'category_id': torch.ones(B, dtype=torch.int64)},
                   logits)
    ap50 = m.result()
    assert ap50 == pytest.approx(1.0, rel=1e-12)
    assert json.loads((tmp_path / "results.json").read_text()) == m.result_list
    want = oo.evaluate(gts, m.result_list)['stats']
    assert np.allclose(m.stats, want, rtol=1e-12, atol=0)
    # result() evaluated the arrays result_list was built from (no dict walk); the dict path gives the same bits, and an edited
    # result_list switches to it
    assert m._arrays and np.array_equal(m.stats, m._evaluator.evaluate(m.result_list)['stats'])
    fast = m.stats.copy()
    m.result_list.pop()
    m.result()
    assert not np.array_equal(m.stats, fast)
    m.reset_states()
    with pytest.raises(IndexError):
        m.result()
