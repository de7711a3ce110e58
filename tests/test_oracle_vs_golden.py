"""Pin the oracle to the golden vectors the UNMODIFIED reference produced (oracle/make_golden.py).

Runs everywhere (CPU only).  Integer / index outputs must be bit-exact; float outputs are
compared bit-exactly too where the oracle performs the same CPU torch / numpy arithmetic
as the reference, and at 1e-6 where only summation order may differ.
"""
import numpy as np
import pytest
import torch

from conftest import golden_rows, load_golden
from helpers import allclose, assert_joints, assert_rows, assert_spm_people, close
from oracle import cases
from oracle import sbp_oracle as so
from oracle import spm_oracle as po


REL = 1e-6


@pytest.mark.parametrize("name", list(cases.SBP_SHAPES))
def test_sbp_render_loss_decode_rows(name):
    g = load_golden("sbp_" + name)
    kp, logits, bbox, iid, cid, meta = cases.sbp_case(name)
    assert cases.digest(kp, logits, bbox) == str(g["inputs_sha"]), "input regeneration drifted"
    k, h, w, sigma, in_size = meta["k"], meta["h"], meta["w"], meta["sigma"], meta["input_size"]

    # template + render: bit exact, both forms
    assert np.array_equal(so.gauss_template(so.resolve_sigma(sigma, h)), g["template"])
    t_vec = so.sbp_render(kp, h, w, sigma)
    assert t_vec.dtype == np.float32 and np.array_equal(t_vec, g["target"])
    t_loop = np.stack([so.sbp_render_loop(kp[b], h, w, sigma) for b in range(kp.shape[0])])
    assert np.array_equal(t_loop, g["target"])

    # loss + gradient (same CPU torch ops -> bit exact)
    loss, grad = so.sbp_loss_and_grad(logits, torch.from_numpy(g["target"]))
    assert close(loss, g["loss"], REL)          # sum order depends on the torch thread count
    assert allclose(grad.numpy(), g["dlogits"], REL)   # vector/scalar sigmoid paths differ by an ulp with chunking
    # fp64 closed form agrees with the reference to fp32 accuracy
    l64, g64 = so.sbp_loss_closed_form_f64(logits, torch.from_numpy(g["target"]))
    assert abs(float(l64) - float(g["loss"])) <= 2e-6 * abs(float(l64))
    scale = float(g64.abs().max())
    assert float((g64 - torch.from_numpy(g["dlogits"]).double()).abs().max()) <= 2e-6 * scale

    # decode
    for thr in (0.25, 0.99):
        want = g[f"joints_pred_thr{thr}"]
        got = so.sbp_decode(logits, in_size[1], thr, True).numpy()
        assert_joints(got, want, REL)
        got_loop = np.stack([so.sbp_decode_loop(logits[b:b + 1], in_size[1], thr, True).numpy()
                             for b in range(logits.size(0))])
        assert_joints(got_loop, want, REL)
    got = so.sbp_decode(torch.from_numpy(g["target"]), in_size[1], 0.99, False).numpy()
    assert np.array_equal(got, g["joints_target_thr0.99"])

    # update_state rows (COCO and PIS)
    joints = so.sbp_decode(logits, in_size[1], 0.25, True)
    img = so.sbp_backproject(joints, bbox, in_size)
    assert_rows(so.sbp_result_rows(img, iid, cid), golden_rows(g, "rows_coco"), REL)
    assert_rows(so.sbp_result_rows(img, iid, cid, pad=18), golden_rows(g, "rows_pis"), REL)


def test_sbp_render_decode_roundtrip_property():
    """decode(render(kp), thr .99, pred=False) == (4*int(x), 4*int(y), 1) visible / (-4,-4,-1) invisible.

    The implicit known-answer test in dataset/sbp_coco_dataset.py:299-320.
    """
    kp, *_ = so.make_config1_inputs(16)
    t = torch.from_numpy(so.sbp_render(kp, 64, 48, 2))
    j = so.sbp_decode(t, 192, 0.99, False).numpy()
    vis = kp[..., 0] >= 0
    assert np.array_equal(j[vis][:, 0], 4.0 * np.trunc(kp[vis][:, 0]))
    assert np.array_equal(j[vis][:, 1], 4.0 * np.trunc(kp[vis][:, 1]))
    assert np.all(j[vis][:, 2] == 1.0)
    assert np.all(j[~vis] == np.array([-4.0, -4.0, -1.0], dtype=np.float32))


def test_sbp_adversarial_decode():
    g = load_golden("sbp_adversarial")
    maps = cases.sbp_adversarial_maps()
    assert cases.digest(maps) == str(g["inputs_sha"])
    for thr in (0.25, 0.5):
        assert_joints(so.sbp_decode(maps, 192, thr, True), g[f"joints_pred_thr{thr}"], REL)
    assert np.array_equal(so.sbp_decode(maps, 192, 0.99, False).numpy(), g["joints_raw_thr0.99"])
    # known answers: constant map -> index 0; last-element maximum; all-below-threshold -> (-4,-4,-1)
    j = g["joints_pred_thr0.25"]
    assert np.all(j[1] == np.array([-4, -4, -1], dtype=np.float32))
    assert np.all(j[4][:, :2] == np.array([47 * 4, 63 * 4], dtype=np.float32))
    assert np.all(j[5][:, :2] == 0)
    assert np.all(j[2][:, 2] == 1.0)


def test_sbp_config1_scalars():
    g = load_golden("sbp_config1")
    kp, logits, bbox, iid, cid = so.make_config1_inputs(32)
    assert cases.digest(kp, logits, bbox) == str(g["inputs_sha"])
    t = so.sbp_render(kp, 64, 48, 2)
    assert float(t.astype(np.float64).sum()) == float(g["target_sum"])
    assert int((t > 0).sum()) == int(g["target_nnz"])
    loss, grad = so.sbp_loss_and_grad(logits, torch.from_numpy(t))
    assert close(loss, g["loss"], REL)
    assert allclose(grad[:2].numpy(), g["grad_slice"], REL)
    assert_joints(so.sbp_decode(logits, 192, 0.25, True), g["joints"], REL)


@pytest.mark.parametrize("name,n", [("small", 6), ("coco", 4)])
def test_spm_render_loss_decode_rows(name, n):
    g = load_golden("spm_" + name)
    people, target, logits, meta = cases.spm_case(name, n)
    assert cases.digest(logits, *[a for p in people for a in p]) == str(g["inputs_sha"])
    k, res, sigma, in_size = meta["k"], meta["res"], meta["sigma"], meta["input_size"]
    assert np.array_equal(target, g["target"])

    tt = torch.from_numpy(g["target"])
    loss, grad = po.spm_loss_and_grad(logits, tt)
    assert close(loss, g["loss"], REL)
    if "dlogits" in g:
        assert allclose(grad.numpy(), g["dlogits"], REL)
    else:
        assert allclose(grad[:1, :, 40:88, 40:88].numpy(), g["grad_slice"], REL)
    l64, g64 = po.spm_loss_closed_form_f64(logits, tt)
    assert abs(float(l64) - float(g["loss"])) <= 2e-6 * abs(float(l64))
    assert float((g64 - grad.double()).abs().max()) <= 2e-6 * float(g64.abs().max())

    for tag, src, pred, thr in (("pred", logits, True, 0.5), ("target", tt, False, 0.99)):
        for b in range(n):
            r, kj = po.spm_decode(src[b:b + 1], in_size, sigma, thr, pred)
            assert_spm_people(r, kj, g[f"roots_{tag}_{b}"], g[f"kps_{tag}_{b}"], REL)

    rows = po.spm_result_rows(logits, [torch.from_numpy(g["image_w"]), torch.from_numpy(g["image_h"])],
                              torch.arange(n) + 7, torch.ones(n, dtype=torch.int64), in_size, sigma, 0.5, True)
    assert_rows(rows, golden_rows(g, "rows"), REL)
