"""Comparison helpers shared by the CPU and GPU parity tests.

Tolerances (BASELINE.json north_star): indices / peak picks bit-exact; confidences,
coordinates, loss and gradients within `rel` relative in fp32 (1e-5 for CUDA-vs-oracle,
1e-6 for oracle-vs-reference where only ulp-level CPU vectorisation differences exist).
"""
import numpy as np


def to_np(a):
    try:
        import torch
        if isinstance(a, torch.Tensor):
            return a.detach().cpu().numpy()
    except ImportError:
        pass
    return np.asarray(a)


def close(a, b, rel):
    a, b = float(a), float(b)
    return abs(a - b) <= rel * abs(b)


def allclose(a, b, rel, atol_frac=0.1):
    """elementwise |a-b| <= rel*|b| + atol_frac*rel*max|b|."""
    a, b = to_np(a).astype(np.float64), to_np(b).astype(np.float64)
    if a.shape != b.shape:
        return False
    if b.size == 0:
        return True
    return bool(np.all(np.abs(a - b) <= rel * np.abs(b) + atol_frac * rel * np.abs(b).max()))


def assert_joints(got, want, rel):
    """[..., 3] rows (x, y, conf): x, y bit-exact (they are integer pixel indices times a ratio), conf within rel."""
    got, want = to_np(got), to_np(want)
    assert got.shape == want.shape, (got.shape, want.shape)
    bad = np.argwhere(got[..., :2] != want[..., :2])
    assert bad.size == 0, f"{len(bad)} coordinate mismatches, first at {bad[0]}: got {got[tuple(bad[0][:-1])]} want {want[tuple(bad[0][:-1])]}"
    assert np.array_equal(got[..., 2] < 0, want[..., 2] < 0)
    assert np.all(np.abs(got[..., 2].astype(np.float64) - want[..., 2]) <= rel * np.abs(want[..., 2])), \
        float(np.abs(got[..., 2].astype(np.float64) - want[..., 2]).max())


def assert_rows(got, want, rel):
    """COCO-result rows: ids exact, keypoint zero-pattern exact, coordinates and score within rel."""
    assert len(got) == len(want), (len(got), len(want))
    for r, w in zip(got, want):
        assert r["image_id"] == w["image_id"] and r["category_id"] == w["category_id"]
        a, b = np.array(r["keypoints"], dtype=np.float64), np.array(w["keypoints"], dtype=np.float64)
        assert a.shape == b.shape
        assert np.array_equal(a == 0, b == 0), (r, w)
        # coordinates: relative to the value, plus 0.1*rel of the row's coordinate range (kx = disp*z + x cancels near 0)
        assert np.all(np.abs(a - b) <= rel * np.abs(b) + 0.1 * rel * np.abs(b).max() + 1e-12), (r, w)
        assert abs(r["score"] - w["score"]) <= rel * abs(w["score"]) + 1e-12, (r["score"], w["score"])


def assert_spm_people(got_roots, got_kps, want_roots, want_kps, rel):
    """roots [N,3] (x,y exact; conf rel), kps [N,K,3] (zero pattern exact, values rel) -- or empties."""
    gr, gk, wr, wk = map(to_np, (got_roots, got_kps, want_roots, want_kps))
    assert gr.shape == wr.shape, (gr.shape, wr.shape)
    assert gk.shape == wk.shape, (gk.shape, wk.shape)
    if wr.size == 0:
        return
    assert np.array_equal(gr[:, :2], wr[:, :2]), (gr, wr)
    assert np.all(np.abs(gr[:, 2].astype(np.float64) - wr[:, 2]) <= rel * np.abs(wr[:, 2]))
    zg, zw = np.all(gk == 0, axis=-1), np.all(wk == 0, axis=-1)
    assert np.array_equal(zg, zw)
    assert allclose(gk, wk, rel)


def chained_field(res=64, k=5, seed=3):
    """A displacement field that encodes a 2-person, 5-joint skeleton HIERARCHICALLY (joint k relative to parents[k]): every
    joint's (dx, dy)/z is written at its parent's pixel (3x3 neighbourhood), the way the SPM paper's hierarchical SPR is trained."""
    rng = np.random.default_rng(seed)
    parents = [-1, 0, 1, -1, 3]
    z = np.sqrt(2.0 * res * res)
    disp = np.zeros((2 * k, res, res), np.float32)
    roots, joints = [], []
    for centre in ((16, 20), (44, 40)):
        pos = {-1: centre}
        for j in range(k):
            px, py = pos[parents[j]]
            jx, jy = px + int(rng.integers(5, 9)), py + int(rng.integers(-8, -4))
            pos[j] = (jx, jy)
            disp[2 * j, py - 1:py + 2, px - 1:px + 2] = (jx - px) / z
            disp[2 * j + 1, py - 1:py + 2, px - 1:px + 2] = (jy - py) / z
        roots.append([float(centre[0]), float(centre[1]), 0.9])
        joints.append([pos[j] for j in range(k)])
    import torch
    return torch.tensor(roots), torch.from_numpy(disp), parents, np.array(joints, np.float64)
