"""GPU parity tests, SPM path (config 4): render, loss fwd/bwd, root NMS + displacement decode, COCO rows.

Peak picks (root x, y, count, order) are bit-exact; rendered targets are bit-exact; loss, gradients, confidences and
joint coordinates within REL = 1e-5 relative (fp32).
"""
import numpy as np
import pytest
import torch

from conftest import golden_rows, load_golden
from helpers import allclose, assert_rows, assert_spm_people, chained_field as _chained_field, close
from oracle import cases
from oracle import spm_oracle as po

pytestmark = pytest.mark.gpu
REL = 1e-5


@pytest.fixture(scope="module")
def pb():
    import pose_b200
    assert torch.cuda.is_available(), "GPU tests need a CUDA device (no CPU fallback exists)"
    pose_b200.lib()
    return pose_b200


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda", 0)


@pytest.mark.parametrize("name,n", [("small", 6), ("coco", 4)])
def test_render_bit_exact(pb, dev, name, n):
    g = load_golden("spm_" + name)
    people, target, logits, meta = cases.spm_case(name, n)
    c, j, cnt = cases.pack_people(people)
    got = pb.spm_render_batch(c, j, cnt, meta["res"], meta["sigma"]).cpu().numpy()
    assert got.dtype == np.float32 and got.shape == g["target"].shape
    assert np.array_equal(got, g["target"])
    # padding beyond counts[i] must be ignored whatever it holds
    c2, j2, _ = cases.pack_people(people, pmax=c.shape[1] + 3)
    c2[:, c.shape[1]:] = 50
    j2[:, c.shape[1]:] = 60
    assert np.array_equal(pb.spm_render_batch(c2, j2, cnt, meta["res"], meta["sigma"]).cpu().numpy(), g["target"])


def test_render_many_persons_and_empty_images(pb, dev):
    """More persons than one shared-memory pass (64), overlapping boxes (displacements add up in person order),
    images without persons, and persons whose centre is (0, 0) (skipped)."""
    rng = np.random.default_rng(3)
    k, res, sigma = 3, 64, 1
    people = []
    for p in (70, 0, 130, 1):
        c = rng.integers(2, res - 2, size=(p, 1, 2), dtype=np.int64)
        j = np.clip(c + rng.integers(-20, 21, size=(p, k, 2), dtype=np.int64), 0, res - 1)
        if p:
            c[0] = 0                              # centre (0,0): skipped by all three generators
            j[p // 2, 1] = 0                      # joint (0,0): skipped
        people.append((c, j))
    want = np.stack([po.spm_render(c, j, res, sigma) for c, j in people])
    pmax = 130
    cc = np.zeros((4, pmax, 2), dtype=np.int64)
    jj = np.zeros((4, pmax, k, 2), dtype=np.int64)
    cnt = np.array([70, 0, 130, 1], dtype=np.int32)
    for i, (c, j) in enumerate(people):
        cc[i, :c.shape[0]] = c[:, 0]
        jj[i, :c.shape[0]] = j
    got = pb.spm_render_batch(cc, jj, cnt, res, sigma).cpu().numpy()
    assert np.array_equal(got, want)


def test_generator_dropins(pb, dev):
    """The reference's three-object protocol (dataset/spm_coco_dataset.py:77-86), per image, numpy in / numpy out."""
    g = load_golden("spm_small")
    people, target, logits, meta = cases.spm_case("small", 6)
    k, res, sigma = meta["k"], meta["res"], meta["sigma"]
    hg, mg, dg = pb.SPMHeatmapGenerator(res, 1, sigma), pb.SPMMaskGenerator(res, sigma), pb.SPMDisplacementGenerator(res, k)
    for i, (c, j) in enumerate(people[:3]):
        hm = hg(c)
        masks = mg(c)
        disp = dg(j, masks)
        assert hm.shape == (1, res, res) and disp.shape == (2 * k, res, res)
        assert np.array_equal(np.asarray(masks), po.spm_masks(c, res, sigma))
        assert np.array_equal(np.concatenate([hm, disp], axis=0), g["target"][i])


@pytest.mark.parametrize("res,k,n", [(20, 2, 3), (36, 1, 5), (132, 3, 2), (256, 2, 1)])
def test_dense_loss_odd_sizes_and_mask_bits(pb, dev, res, k, n):
    """Dense-target loss at map sizes whose planes are not whole 16 KB units / whose quads do not fill the mask words: the root
    mask travels as bits (spm_root_mask_kernel); target values <= 0, NaN and -0.0 in the root plane are mask 0 like the
    reference's torch.where(t > 0); loss / gradients against the fp64 closed form, with and without dlogits; empty batch."""
    rng = np.random.default_rng(res)
    people = []
    for p in rng.integers(0, 5, size=n):
        c = rng.integers(0, res, size=(p, 1, 2), dtype=np.int64)
        j = np.clip(c + rng.integers(-9, 10, size=(p, k, 2), dtype=np.int64), 0, res - 1)
        people.append((c, j))
    target = torch.from_numpy(np.stack([po.spm_render(c, j, res, 1) for c, j in people]))
    target[0, 0, 0, :3] = torch.tensor([-0.0, -1.0, 1e-30])                     # mask 0, 0, 1
    x = torch.randn(target.shape, generator=torch.Generator().manual_seed(res)) * 2
    l64, g64 = po.spm_loss_closed_form_f64(x, target)
    r = pb.spm_loss_fused(x.to(dev), target.to(dev))
    assert close(r["loss"].item(), float(l64), REL) and allclose(r["dlogits"], g64, REL, atol_frac=1.0)
    r0 = pb.spm_loss_fused(x.to(dev), target.to(dev), want_grad=False)
    assert r0["dlogits"] is None and r0["loss"].item() == r["loss"].item()
    assert torch.equal(r["loss_num"], r0["loss_num"])
    e = pb.spm_loss_fused(torch.zeros((0, 1 + 2 * k, res, res), device=dev), torch.zeros((0, 1 + 2 * k, res, res), device=dev))
    assert e["loss"].item() == 0.0


@pytest.mark.parametrize("name,n", [("small", 6), ("coco", 4)])
def test_loss_and_grad(pb, dev, name, n):
    g = load_golden("spm_" + name)
    people, target, logits, meta = cases.spm_case(name, n)
    tt = torch.from_numpy(g["target"])
    l64, g64 = po.spm_loss_closed_form_f64(logits, tt)
    x = logits.to(dev).requires_grad_(True)
    loss = pb.SPMLoss()(x, tt.to(dev))
    assert loss.dim() == 0 and loss.requires_grad
    loss.backward()
    assert close(loss.item(), float(g["loss"]), REL), (loss.item(), float(g["loss"]))
    assert close(loss.item(), float(l64), REL)
    # displacement gradients carry (1 - tanh^2): with |tanh| ~ 0.999 the fp32 subtraction cancels three digits, in the
    # reference as much as here, so the absolute floor is 1e-5 of the gradient scale (atol_frac=1), not 1e-6
    assert allclose(x.grad, g64, REL, atol_frac=1.0)
    if "dlogits" in g:
        assert allclose(x.grad, g["dlogits"], REL, atol_frac=1.0)
    else:
        assert allclose(x.grad[:1, :, 40:88, 40:88], g["grad_slice"], REL, atol_frac=1.0)
        assert close(x.grad.double().abs().sum().item(), float(g["grad_abs_sum"]), REL)
    x2 = logits.to(dev).requires_grad_(True)
    (pb.SPMLoss()(x2, tt.to(dev)) * 2.5).backward()
    assert allclose(x2.grad, 2.5 * g64, REL, atol_frac=1.0)
    with torch.no_grad():
        assert close(pb.SPMLoss()(logits.to(dev), tt.to(dev)).item(), float(g["loss"]), REL)
    # random logits against a target (exercises |d| >= 1 SmoothL1 branch and saturated activations)
    gen = torch.Generator().manual_seed(5)
    xr = torch.randn(logits.shape, generator=gen) * 4
    lr, gr = po.spm_loss_closed_form_f64(xr, tt)
    x3 = xr.to(dev).requires_grad_(True)
    l3 = pb.SPMLoss()(x3, tt.to(dev))
    l3.backward()
    assert close(l3.item(), float(lr), REL) and allclose(x3.grad, gr, REL, atol_frac=1.0)


@pytest.mark.parametrize("name,n", [("small", 6), ("coco", 4)])
def test_decode_matches_reference(pb, dev, name, n):
    g = load_golden("spm_" + name)
    people, target, logits, meta = cases.spm_case(name, n)
    in_size, sigma = meta["input_size"], meta["sigma"]
    tt = torch.from_numpy(g["target"])
    for tag, src, pred, thr in (("pred", logits, True, 0.5), ("target", tt, False, 0.99)):
        dec = pb.DecodeSPM(in_size, sigma, thr, pred)
        for b in range(n):
            r, kj = dec(src[b:b + 1].to(dev))
            assert_spm_people(r, kj, g[f"roots_{tag}_{b}"], g[f"kps_{tag}_{b}"], REL)
        # batched form agrees with the per-image drop-in
        roots, kps, counts, total = dec.decode_batch(src.to(dev))
        for b in range(n):
            c = int(counts[b])
            assert c == g[f"roots_{tag}_{b}"].shape[0] and int(total[b]) == c
            assert_spm_people(roots[b, :c], kps[b, :c], g[f"roots_{tag}_{b}"], g[f"kps_{tag}_{b}"], REL)
    m = pb.SPMmAPCOCO(None, in_size, sigma, 0.5)
    m.update_state({"image_size": [torch.from_numpy(g["image_w"]), torch.from_numpy(g["image_h"])],
                    "image_id": torch.arange(n) + 7, "category_id": torch.ones(n, dtype=torch.int64)}, logits.to(dev))
    assert_rows(m.result_list, golden_rows(g, "rows"), REL)


def test_nms_rules_empty_overflow_and_gather(pb, dev):
    # empty decode returns two shape-[0] tensors (utils/spm_utils.py:120-121, :180-181)
    x = torch.full((1, 3, 16, 16), -9.0, device=dev)
    r, k = pb.DecodeSPM(64, 1, 0.5, True)(x)
    assert tuple(r.shape) == (0,) and tuple(k.shape) == (0,)
    assert pb.SPMmAPCOCO(None, 64, 1, 0.5).result() == 0
    # radius rule: distance exactly == threshold is suppressed (strict >), sqrt(17) survives; ties -> row-major
    h = torch.zeros(1, 16, 16)
    h[0, 5, 5], h[0, 5, 9], h[0, 9, 6] = 0.9, 0.8, 0.7
    got = pb.nms_spm(h.to(dev), 0.5, 4.0).cpu()
    assert torch.equal(got, po.spm_nms(h, 0.5, 4.0)) and got.shape[0] == 2
    h2 = torch.zeros(1, 16, 16)
    h2[0, 2, 12] = h2[0, 2, 3] = h2[0, 10, 3] = 0.75
    assert torch.equal(pb.nms_spm(h2.to(dev), 0.5, 4.0).cpu(), po.spm_nms(h2, 0.5, 4.0))
    # more roots than the fixed buffer: counts are capped, totals reported, the drop-in retries with the exact size
    dense_map = torch.full((1, 3, 32, 32), 0.0)
    gen = torch.Generator().manual_seed(1)
    dense_map[0, 0] = torch.rand(32, 32, generator=gen) * 0.5 + 0.5
    roots, kps, counts, total = pb.spm_decode_batch(dense_map.to(dev), 32, 1, 0.5, pred=False, max_people=4)
    want = po.spm_nms(dense_map[0, 0:1], 0.5, 4.0)
    assert int(counts[0]) == 4 and int(total[0]) == want.shape[0] > 4
    assert torch.equal(roots[0, :4, :2].cpu(), want[:4, :2])
    r, _ = pb.DecodeSPM(32, 1, 0.5, False, max_people=4)(dense_map.to(dev))
    assert torch.equal(r.cpu(), want)
    # dense map with more candidates than the kernel's candidate list holds (4096 > 2048): bitmap fallback, same picks;
    # sigmoid-activated variant of the same map goes through the pred=True path
    dense64 = torch.zeros((1, 3, 64, 64))
    dense64[0, 0] = torch.rand(64, 64, generator=gen) * 0.45 + 0.55
    want64 = po.spm_nms(dense64[0, 0:1], 0.5, 4.0)
    r64, _ = pb.DecodeSPM(64, 1, 0.5, False, max_people=8)(dense64.to(dev))
    assert want64.shape[0] > 8 and torch.equal(r64.cpu(), want64)
    lg64 = dense64.clone()
    lg64[0, 0] = torch.logit(dense64[0, 0])
    rl, kl = pb.DecodeSPM(64, 1, 0.5, True, max_people=300)(lg64.to(dev))
    wl, wkl = po.spm_decode(lg64, 64, 1, 0.5, True)
    assert_spm_people(rl, kl, wl, wkl, REL)
    # get_spm_keypoints drop-in
    people, target, logits, meta = cases.spm_case("small", 2)
    tt = torch.from_numpy(target)
    roots = po.spm_nms(tt[0, 0:1], 0.99, 4.0)
    want_k = po.spm_keypoints(roots, tt[0, 1:], 4.0)
    got_k = pb.get_spm_keypoints(roots.to(dev), tt[0, 1:].to(dev), 4.0)
    assert allclose(got_k, want_k, REL) and torch.equal(got_k.cpu() == 0, want_k == 0)


def test_hierarchical_chain_opt_in(pb, dev):
    """SURVEY 8 f-4, PARITY UNPINNED (not in the reference): get_spm_keypoints_chained against the oracle's restatement; with
    every parent = -1 it is the pinned single hop, bit for bit."""
    people, target, logits, meta = cases.spm_case("coco", 4, seed=31)
    tt = torch.from_numpy(target)
    k = tt.shape[1] // 2
    for b in range(4):
        roots = po.spm_nms(tt[b, 0:1], 0.99, 4.0)
        if roots.dim() != 2 or roots.shape[0] == 0:
            continue
        single = pb.get_spm_keypoints(roots.to(dev), tt[b, 1:].to(dev), 4.0)
        assert torch.equal(pb.get_spm_keypoints_chained(roots.to(dev), tt[b, 1:].to(dev), [-1] * k, 4.0), single)
        # a COCO-like tree on the (single-hop encoded) field: compared with the oracle's restatement of the same chain
        parents = [-1, 0, 0, 1, 2, -1, -1, 5, 6, 7, 8, -1, -1, 11, 12, 13, 14][:k]
        want = po.spm_keypoints_chained(roots, tt[b, 1:], parents, 4.0)
        got = pb.get_spm_keypoints_chained(roots.to(dev), tt[b, 1:].to(dev), parents, 4.0)
        assert allclose(got, want, REL) and torch.equal(got.cpu() == 0, want == 0)
    r, disp, parents, joints = _chained_field()
    got = pb.get_spm_keypoints_chained(r.to(dev), disp.to(dev), parents, 4.0)
    want = po.spm_keypoints_chained(r, disp, parents, 4.0)
    assert torch.equal(got.cpu(), want)
    assert np.allclose(got[..., :2].cpu().numpy(), joints, atol=1e-3)
    cyc = pb.get_spm_keypoints_chained(r.to(dev), disp.to(dev), [1, 0, -1, -1, 3], 4.0)
    assert torch.equal(cyc.cpu(), po.spm_keypoints_chained(r, disp, [1, 0, -1, -1, 3], 4.0))


def test_config4_batch_decode_against_oracle(pb, dev):
    """Config 4 at a larger batch: 64 multi-person images, every image's picks against the oracle."""
    people, target, logits, meta = cases.spm_case("coco", 64, seed=777)
    c, j, cnt = cases.pack_people(people)
    t = pb.spm_render_batch(c, j, cnt, 128, 1)
    assert np.array_equal(t.cpu().numpy(), target)
    roots, kps, counts, total = pb.spm_decode_batch(logits.to(dev), 512, 1, 0.5, True, max_people=32)
    for b in range(0, 64, 4):
        wr, wk = po.spm_decode(logits[b:b + 1], 512, 1, 0.5, True)
        n = int(counts[b])
        assert_spm_people(roots[b, :n], kps[b, :n], wr, wk, REL)
    l64, _ = po.spm_loss_closed_form_f64(logits[:8], torch.from_numpy(target[:8]))
    assert close(pb.spm_loss_fused(logits[:8].to(dev), torch.from_numpy(target[:8]).to(dev), want_grad=False)["loss"].item(),
                 float(l64), REL)


# --------------------------------------------------------------------------- fused render + loss (+grad): persons in, no dense target
@pytest.mark.parametrize("name,n", [("small", 6), ("coco", 4)])
def test_fused_render_loss_grad(pb, dev, name, n):
    """pose_spm_fused: the target it renders in registers is bit-identical to the golden (reference-generated) target,
    loss / gradients match the golden reference values and the dense-target kernel."""
    g = load_golden("spm_" + name)
    people, target, logits, meta = cases.spm_case(name, n)
    c, j, cnt = cases.pack_people(people)
    x = logits.to(dev)
    r = pb.spm_fused(x, c, j, cnt, meta["sigma"], want_grad=True, want_target=True)
    assert np.array_equal(r["target"].cpu().numpy(), g["target"])
    assert close(r["loss"].item(), float(g["loss"]), REL)
    l64, g64 = po.spm_loss_closed_form_f64(logits, torch.from_numpy(g["target"]))
    assert close(r["loss"].item(), float(l64), REL) and allclose(r["dlogits"], g64, REL, atol_frac=1.0)
    # same arithmetic per element as the dense-target kernel: gradients bit-identical, loss equal up to summation order
    d = pb.spm_loss_fused(x, torch.from_numpy(g["target"]).to(dev))
    assert torch.equal(r["dlogits"], d["dlogits"])
    assert close(r["loss"].item(), d["loss"].item(), 1e-6)
    # without grad / target outputs, and padding beyond counts[i] ignored
    c2, j2, _ = cases.pack_people(people, pmax=c.shape[1] + 3)
    c2[:, c.shape[1]:] = 9
    j2[:, c.shape[1]:] = 11
    r2 = pb.spm_fused(x, c2, j2, cnt, meta["sigma"], want_grad=False)
    assert r2["dlogits"] is None and r2["target"] is None and r2["loss"].item() == r["loss"].item()
    # SPMLoss drop-in with the persons hand-off, autograd with a scaled upstream gradient
    x3 = logits.to(dev).requires_grad_(True)
    loss = pb.SPMLoss(sigma=meta["sigma"])(x3, {"centers": torch.from_numpy(c), "joints": torch.from_numpy(j), "counts": torch.from_numpy(cnt)})
    assert loss.dim() == 0 and loss.requires_grad
    (loss * 2.5).backward()
    assert close(loss.item(), float(g["loss"]), REL) and allclose(x3.grad, 2.5 * g64, REL, atol_frac=1.0)


def test_fused_crowded_edges_and_nan(pb, dev):
    """64 persons per image (the fused kernel's limit) with overlapping boxes, centres on the border / off the map / (0,0),
    dropped joints, an empty image, random logits (saturated activations, |d| >= 1 SmoothL1 branch); non-integer sigma
    (box and Gaussian support differ); more than 64 persons falls back to render + dense loss; NaN logits propagate."""
    rng = np.random.default_rng(17)
    # res 64 / 32 / 36: linear warp mapping; 256 (8 tile columns) and 384 (12: not a power of two): the tiled mapping
    for k, res, sigma, ps in ((3, 64, 1, (64, 0, 37, 1)), (2, 32, 1.5, (5, 9, 0, 2)), (2, 36, 0.5, (3, 1, 4, 2)),
                              (2, 256, 2.5, (9, 0, 3)), (1, 384, 1, (6, 2))):
        people = []
        for p in ps:
            c = rng.integers(-3, res + 3, size=(p, 1, 2), dtype=np.int64)
            j = np.clip(c + rng.integers(-20, 21, size=(p, k, 2), dtype=np.int64), 0, res - 1)
            if p > 1:
                c[0] = 0
                j[p // 2, k - 1] = 0
            people.append((c, j))
        want_t = np.stack([po.spm_render(c, j, res, sigma) for c, j in people])
        c, j, cnt = cases.pack_people(people)
        xr = torch.randn((len(ps), 1 + 2 * k, res, res), generator=torch.Generator().manual_seed(3)) * 4
        r = pb.spm_fused(xr.to(dev), c, j, cnt, sigma, want_grad=True, want_target=True)
        assert np.array_equal(r["target"].cpu().numpy(), want_t), (k, res, sigma)
        assert np.array_equal(pb.spm_render_batch(c, j, cnt, res, sigma).cpu().numpy(), want_t)
        l64, g64 = po.spm_loss_closed_form_f64(xr, torch.from_numpy(want_t))
        assert close(r["loss"].item(), float(l64), REL) and allclose(r["dlogits"], g64, REL, atol_frac=1.0)
    # > 64 persons: SPMLoss renders and uses the dense kernel; the raw entry point refuses
    k, res = 2, 64
    c = rng.integers(2, res - 2, size=(1, 70, 2), dtype=np.int64)
    j = np.clip(c[:, :, None, :] + rng.integers(-9, 10, size=(1, 70, k, 2), dtype=np.int64), 1, res - 2)
    cnt = np.array([70], dtype=np.int32)
    xr = torch.randn((1, 1 + 2 * k, res, res), generator=torch.Generator().manual_seed(4))
    want_t = po.spm_render(c[0][:, None], j[0], res, 1)[None]
    l64, _ = po.spm_loss_closed_form_f64(xr, torch.from_numpy(want_t))
    assert close(pb.SPMLoss(sigma=1)(xr.to(dev), (c, j, cnt)).item(), float(l64), REL)
    with pytest.raises(pb.PoseB200Error):
        pb.spm_fused(xr.to(dev), c, j, cnt, 1)
    # a NaN logit anywhere (even under a zero mask) makes the loss NaN, as tanh(nan)*0 does in the reference
    xn = xr.clone()
    xn[0, 3, 1, 1] = float("nan")
    assert torch.isnan(pb.spm_fused(xn.to(dev), c[:, :8], j[:, :8], np.array([8], dtype=np.int32), 1, want_grad=False)["loss"])
    # empty batch
    e = pb.spm_fused(torch.zeros((0, 5, 32, 32), device=dev), np.zeros((0, 1, 2), np.int64), np.zeros((0, 1, 2, 2), np.int64),
                     np.zeros((0,), np.int32), 1)
    assert e["loss"].item() == 0.0 and e["dlogits"].shape[0] == 0


@pytest.mark.parametrize("res", [128, 64])
def test_fused_readonly_screen_nan_and_inf(pb, dev, res):
    """The read-only form screens EVERY quad for NaN (covered or not) and, at R=128, does so before the covered-quad list exists:
    NaN anywhere -> NaN loss in both forms; +-inf under a zero mask adds nothing (sigmoid / tanh of +-inf times 0), also when
    +inf and -inf share a quad (their sum is NaN, the per-element screen is not); results equal the writing form and the oracle."""
    rng = np.random.default_rng(5)
    k, sigma = 3, 1
    people = []
    for p in (4, 0, 7):
        c = rng.integers(8, res - 8, size=(p, 1, 2), dtype=np.int64)
        j = np.clip(c + rng.integers(-12, 13, size=(p, k, 2), dtype=np.int64), 0, res - 1)
        people.append((c, j))
    target = np.stack([po.spm_render(c, j, res, sigma) for c, j in people])
    c, j, cnt = cases.pack_people(people)
    x0 = torch.randn((3, 1 + 2 * k, res, res), generator=torch.Generator().manual_seed(9)) * 3
    cx, cy = int(people[0][0][0, 0, 0]), int(people[0][0][0, 0, 1])          # a root centre of image 0: covered, mask set
    free = np.argwhere(target[0, 0] == 0)                                       # uncovered pixels of image 0's root plane
    fy, fx = (int(v) for v in free[len(free) // 2])
    fx &= ~3                                                                    # first element of its quad
    assert target[0, 0, fy, fx] == 0 and target[0, 0, fy, fx + 1] == 0

    def both(x):
        g = pb.spm_fused(x.to(dev), c, j, cnt, sigma, want_grad=True)
        r = pb.spm_fused(x.to(dev), c, j, cnt, sigma, want_grad=False)
        return g["loss"].item(), r["loss"].item()
    lg, lr = both(x0)
    l64, _ = po.spm_loss_closed_form_f64(x0, torch.from_numpy(target))
    assert lg == lr and close(lr, float(l64), REL)
    # +-inf under a zero mask: same quad, root plane and a displacement plane; the empty image too
    xi = x0.clone()
    xi[0, 0, fy, fx], xi[0, 0, fy, fx + 1] = float("inf"), float("-inf")
    xi[0, 2, fy, fx], xi[0, 2, fy, fx + 1] = float("inf"), float("-inf")
    xi[1, 4, 5, 8] = float("inf")
    li64, _ = po.spm_loss_closed_form_f64(xi, torch.from_numpy(target))
    lgi, lri = both(xi)
    assert np.isfinite(lri) and lgi == lri and close(lri, float(li64), REL)
    # NaN: uncovered pixel / covered pixel, root plane / displacement plane, first / last image
    for (n, ch, y, x) in ((0, 0, fy, fx), (0, 3, fy, fx + 1), (0, 0, cy, cx), (0, 5, cy, cx), (1, 6, res - 1, res - 1), (2, 1, 0, 0)):
        xn = x0.clone()
        xn[n, ch, y, x] = float("nan")
        lgn, lrn = both(xn)
        assert np.isnan(lgn) and np.isnan(lrn), (n, ch, y, x)


def test_fused_config4_batch(pb, dev):
    """Config 4 at N=64: fused kernel against the dense path on the same rendered target (every image, every plane)."""
    people, target, logits, meta = cases.spm_case("coco", 64, seed=777)
    c, j, cnt = cases.pack_people(people)
    x = logits.to(dev)
    r = pb.spm_fused(x, c, j, cnt, 1, want_grad=True, want_target=True)
    assert np.array_equal(r["target"].cpu().numpy(), target)
    d = pb.spm_loss_fused(x, torch.from_numpy(target).to(dev))
    assert torch.equal(r["dlogits"], d["dlogits"]) and close(r["loss"].item(), d["loss"].item(), 1e-6)
    l64, _ = po.spm_loss_closed_form_f64(logits, torch.from_numpy(target))
    assert close(r["loss"].item(), float(l64), REL)
