"""GPU parity tests, Simple Baselines path: CUDA kernels (through the C ABI / drop-in classes) vs the golden vectors
produced by the unmodified reference and vs the oracle.

Tolerances (BASELINE.json north_star): argmax indices bit-exact incl. first-index tie-breaking; rendered targets
bit-exact; loss, gradients, confidences within REL = 1e-5 relative (fp32).
"""
import ctypes

import numpy as np
import pytest
import torch

from conftest import golden_rows, load_golden
from helpers import allclose, assert_joints, assert_rows, close
from oracle import cases
from oracle import sbp_oracle as so

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


def _sigmoid_ref_eval(pb, x, ref):
    y = torch.empty_like(x)
    C = pb._cabi
    C.check(pb.lib().pose_sigmoid_ref_eval(C.ptr(x), C.ptr(y), x.numel(), C.sigmoid_ref_code(ref), C.stream_ptr(x.device)))
    return y


def _bits(t):
    return t.view(torch.int32)


def test_device_restatements_of_torch_sigmoid_are_bit_exact(pb, dev):
    """The two functions that rank near-ties ARE torch.sigmoid: the CUDA flavour against torch on this GPU for ALL 2^32
    inputs, the CPU flavour (Sleef expf_u10 + IEEE add / divide) against torch on the host for 2^27 random bit patterns
    plus every float in [-20, 20) at stride 64 (the C restatement was compared exhaustively at build time)."""
    chunk = 1 << 28
    for start in range(0, 1 << 32, chunk):
        x = (torch.arange(chunk, device=dev, dtype=torch.int64) + start).to(torch.int32).view(torch.float32)   # wraps to the bit pattern
        got, want = _sigmoid_ref_eval(pb, x, "cuda"), torch.sigmoid(x)
        nan = torch.isnan(want) & torch.isnan(got)
        bad = (_bits(got) != _bits(want)) & ~nan
        assert int(bad.sum()) == 0, (start, x[bad][:4], got[bad][:4], want[bad][:4])
        del x, got, want, nan, bad
    g = torch.Generator().manual_seed(5)
    rnd = torch.randint(-(1 << 31), (1 << 31) - 1, (1 << 27,), generator=g, dtype=torch.int64).to(torch.int32).view(torch.float32)
    lo, hi = np.float32(-20).view(np.int32), np.float32(-1e-30).view(np.int32)
    neg = torch.arange(int(hi), int(lo), 64, dtype=torch.int64).to(torch.int32).view(torch.float32)       # negative floats, descending magnitude
    xs = torch.cat([rnd, neg, -neg]).contiguous()
    want = so.torch_sigmoid_vector_body(xs)              # (chunk / vector tails go through glibc expf in ATen: avoided)
    got = _sigmoid_ref_eval(pb, xs.to(dev), "cpu").cpu()
    nan = torch.isnan(want) & torch.isnan(got)
    bad = (_bits(got) != _bits(want)) & ~nan
    assert int(bad.sum()) == 0, (xs[bad][:4], got[bad][:4], want[bad][:4])


@pytest.mark.parametrize("ref", ["cpu", "cuda"])
def test_candidate_window_excludes_nothing_that_could_win(pb, dev, ref):
    """sigmoid_window_lo: no input below the window of m reaches sigmoid_ref(m) (every 61st float m, 128 probes each)."""
    v = torch.ones(1, dtype=torch.int64, device=dev)
    C = pb._cabi
    C.check(pb.lib().pose_sigmoid_window_check(C.ptr(v), C.sigmoid_ref_code(ref), C.stream_ptr(dev)))
    assert int(v.item()) == 0


def test_decode_confidence_is_the_reference_sigmoid(pb, dev):
    """conf = torch.sigmoid(logit) bit for bit (CPU flavour vs the host, CUDA flavour vs this GPU); saturates to exactly 1."""
    probe = torch.tensor([-20.0, -8.0, -2.5, -0.3, 0.0, 0.7, 3.0, 9.0], device=dev)
    maps = torch.full((1, 8, 64, 48), -50.0, device=dev)
    maps[0, torch.arange(8), 5, 7] = probe
    j = pb.decode_batch(maps, -1.0, 1.0, True, sigmoid_ref="cpu")[0, :, 2].cpu()
    assert torch.equal(j, so.torch_sigmoid_vector_body(probe.cpu()))
    j = pb.decode_batch(maps, -1.0, 1.0, True, sigmoid_ref="cuda")[0, :, 2]
    assert torch.equal(j, torch.sigmoid(probe))
    sat = torch.tensor([17.0, 30.0, 88.0, 1e4], device=dev).reshape(1, 4, 1, 1).expand(1, 4, 4, 4).contiguous()
    assert torch.all(pb.decode_batch(sat, 0.5, 1.0, True)[0, :, 2] == 1.0)


@pytest.mark.parametrize("name", list(cases.SBP_SHAPES))
def test_render_bit_exact(pb, dev, name):
    g = load_golden("sbp_" + name)
    kp, logits, bbox, iid, cid, meta = cases.sbp_case(name)
    gen = pb.SBPHeatmapGenerator([meta["h"], meta["w"]], meta["k"], meta["sigma"])
    assert np.array_equal(gen.g, g["template"])
    got = gen.render_batch(kp).cpu().numpy()
    assert got.dtype == np.float32 and np.array_equal(got, g["target"])
    # per-sample drop-in contract (numpy in, numpy out)
    one = gen(kp[1])
    assert isinstance(one, np.ndarray) and np.array_equal(one, g["target"][1])
    # fp32 keypoints: same kernel, truncation happens on the fp32 value
    kp32 = kp.astype(np.float32)
    want32 = so.sbp_render(kp32.astype(np.float64), meta["h"], meta["w"], meta["sigma"])
    assert np.array_equal(gen.render_batch(torch.from_numpy(kp32).to(dev)).cpu().numpy(), want32)


@pytest.mark.parametrize("name", list(cases.SBP_SHAPES))
def test_loss_and_grad_dense_and_fused(pb, dev, name):
    g = load_golden("sbp_" + name)
    kp, logits, bbox, iid, cid, meta = cases.sbp_case(name)
    target = torch.from_numpy(g["target"]).to(dev)
    want_loss, want_grad = float(g["loss"]), g["dlogits"]
    l64, g64 = so.sbp_loss_closed_form_f64(logits, torch.from_numpy(g["target"]))

    for tgt, kw in ((target, {}), (torch.from_numpy(kp).to(dev), {"sigma": meta["sigma"]})):
        x = logits.to(dev).requires_grad_(True)
        loss = pb.SBPLoss(**kw)(x, tgt)
        assert loss.dim() == 0 and loss.dtype == torch.float32 and loss.device == x.device and loss.requires_grad
        loss.backward()
        assert close(loss.item(), want_loss, REL), (loss.item(), want_loss)
        assert close(loss.item(), float(l64), REL)
        assert allclose(x.grad, want_grad, REL)
        assert allclose(x.grad, g64, REL)

    # arbitrary upstream gradient (e.g. loss / accumulate_grad_batches) is applied without a host sync
    x = logits.to(dev).requires_grad_(True)
    (pb.SBPLoss()(x, target) * 0.37).backward()
    assert allclose(x.grad, 0.37 * g64, REL)
    # no-grad (validation) path returns the same value
    with torch.no_grad():
        assert close(pb.SBPLoss()(logits.to(dev), target).item(), want_loss, REL)


@pytest.mark.parametrize("name", list(cases.SBP_SHAPES))
def test_decode_matches_reference(pb, dev, name):
    g = load_golden("sbp_" + name)
    kp, logits, bbox, iid, cid, meta = cases.sbp_case(name)
    in_size = meta["input_size"]
    x = logits.to(dev)
    for thr in (0.25, 0.99):
        dec = pb.DecodeSBP(list(in_size), thr, True, sigmoid_ref="cpu")
        got = dec.decode_batch(x).cpu().numpy()
        assert np.array_equal(got, g[f"joints_pred_thr{thr}"])        # indices AND confidences bit for bit
        dec = pb.DecodeSBP(list(in_size), thr, True, sigmoid_ref="cuda")
        assert torch.equal(dec.decode_batch(x).cpu(), so.sbp_decode(x, in_size[1], thr, True))     # the reference's ops on this GPU
    dec = pb.DecodeSBP(list(in_size), 0.99, False)
    got = dec.decode_batch(torch.from_numpy(g["target"]).to(dev)).cpu().numpy()
    assert np.array_equal(got, g["joints_target_thr0.99"])        # pred=False: no activation -> bit exact
    # reference call shape: batch of one -> [K,3]
    one = pb.DecodeSBP(list(in_size), 0.25, True)(x[2:3])
    assert tuple(one.shape) == (meta["k"], 3)
    assert_joints(one, g["joints_pred_thr0.25"][2], REL)
    with pytest.raises(AssertionError):
        pb.DecodeSBP(list(in_size), 0.25, True)(x[:2])
    # nms_sbp drop-in on an activated map
    h = torch.sigmoid(logits[0]).to(dev)
    want = so.sbp_decode(torch.sigmoid(logits[0:1]), h.size(-1), 0.8, False)[0]
    assert_joints(pb.nms_sbp(h, 0.8), want, REL)


def test_decode_adversarial_ties_plateaus_borders(pb, dev):
    g = load_golden("sbp_adversarial")
    maps = cases.sbp_adversarial_maps().to(dev)
    for thr in (0.25, 0.5):
        assert np.array_equal(pb.decode_batch(maps, thr, 4.0, True, sigmoid_ref="cpu").cpu().numpy(), g[f"joints_pred_thr{thr}"])
        assert torch.equal(pb.decode_batch(maps, thr, 4.0, True, sigmoid_ref="cuda").cpu(), so.sbp_decode(maps, 192, thr, True))
    got = pb.decode_batch(maps, 0.99, 4.0, False).cpu().numpy()
    assert np.array_equal(got, g["joints_raw_thr0.99"])


def test_decode_near_ties_follow_the_reference_sigmoid_to_the_last_bit(pb, dev):
    """Top logits 1-4 ulp apart at the magnitudes where fp32 sigmoid merges neighbours (0.5 ... 16.5): the reference's pick
    (golden, CPU tensors) is reproduced bit for bit, and so is the reference's pick on CUDA tensors (its ops run here)."""
    g = load_golden("sbp_adversarial")
    near = cases.sbp_neartie_maps()
    assert cases.digest(near) == str(g["neartie_sha"])
    x = near.to(dev)
    got = pb.decode_batch(x, 0.25, 4.0, True, sigmoid_ref="cpu").cpu().numpy()
    assert np.array_equal(got, g["joints_neartie_thr0.25"])
    assert torch.equal(pb.decode_batch(x, 0.25, 4.0, True, sigmoid_ref="cuda").cpu(), so.sbp_decode(x, 192, 0.25, True))
    # the same maps through every fused variant that decodes
    kp = np.full((near.size(0), near.size(1), 2), 10.0)
    for grad in (True, False):
        for ref in ("cpu", "cuda"):
            r = pb.sbp_fused(x, keypoints=kp, sigma=2, want_grad=grad, decode=True, conf_threshold=0.25, coord_scale=4.0, sigmoid_ref=ref)
            assert torch.equal(r["joints"], pb.decode_batch(x, 0.25, 4.0, True, sigmoid_ref=ref))


def test_decode_fp32_threshold_and_tied_maps(pb, dev):
    x = torch.zeros(1, 2, 8, 8, device=dev)
    x[0, 0, 3, 4] = float(np.float32(0.99))
    x[0, 1, 3, 4] = float(np.nextafter(np.float32(0.99), np.float32(2)))
    j = pb.decode_batch(x, 0.99, 4.0, False).cpu().numpy()
    assert np.array_equal(j[0, 0], np.array([-4, -4, -1], dtype=np.float32))
    assert np.array_equal(j[0, 1, :2], np.array([16, 12], dtype=np.float32))
    # random and heavily tied data (7 distinct values per map: every lane holds several candidate vectors -> second look)
    gen = torch.Generator(device=dev).manual_seed(3)
    a = torch.randn(64, 17, 64, 48, device=dev, generator=gen) * 4
    b = torch.randint(-3, 4, (64, 17, 64, 48), device=dev, generator=gen).float() * 6.0
    for t in (a, b):
        for pred in (True, False):
            assert torch.equal(pb.decode_batch(t, 0.25, 4.0, pred, sigmoid_ref="cpu").cpu(), so.sbp_decode(t.cpu(), 192, 0.25, pred))
            assert torch.equal(pb.decode_batch(t, 0.25, 4.0, pred, sigmoid_ref="cuda").cpu(), so.sbp_decode(t, 192, 0.25, pred))


@pytest.mark.parametrize("name", ["coco", "hires"])
def test_fused_render_loss_grad_decode_single_pass(pb, dev, name):
    """One launch: keypoints + logits -> loss, dlogits, rendered target, joints -- all equal to the staged results."""
    g = load_golden("sbp_" + name)
    kp, logits, bbox, iid, cid, meta = cases.sbp_case(name)
    x = logits.to(dev)
    scale = meta["input_size"][1] / meta["w"]
    before = pb.launch_count()
    r = pb.sbp_fused(x, keypoints=kp, sigma=meta["sigma"], want_grad=True, decode=True, conf_threshold=0.25,
                     coord_scale=scale, want_target=True)
    assert pb.launch_count() - before == 2           # fused kernel + the deterministic partial-sum finalize
    assert np.array_equal(r["target"].cpu().numpy(), g["target"])
    assert close(r["loss"].item(), float(g["loss"]), REL)
    assert allclose(r["dlogits"], g["dlogits"], REL)
    assert_joints(r["joints"], g["joints_pred_thr0.25"], REL)
    assert torch.equal(r["joints"], pb.decode_batch(x, 0.25, scale, True))
    # un-normalised numerators reproduce the loss: (5 S_pos + S_neg) / (2 K B)
    num = r["loss_num"].cpu().numpy()
    assert close((5 * num[0] + num[1]) / (2 * meta["k"] * x.size(0)), float(g["loss"]), REL)
    # the same call can also back-project (bbox given): the epilogue launch writes the packed COCO rows
    rb = pb.sbp_fused(x, keypoints=kp, sigma=meta["sigma"], want_grad=True, decode=True, conf_threshold=0.25, coord_scale=scale,
                      bbox=bbox, input_size=meta["input_size"])
    assert torch.equal(rb["packed"], pb.backproject_packed(r["joints"], bbox, meta["input_size"]))
    assert_rows(pb.packed_to_results(rb["packed"], iid, cid), golden_rows(g, "rows_coco"), REL)
    # run-to-run determinism (fixed reduction order, no float atomics)
    r2 = pb.sbp_fused(x, keypoints=kp, sigma=meta["sigma"], want_grad=True, decode=True, conf_threshold=0.25, coord_scale=scale)
    assert torch.equal(r["loss"], r2["loss"]) and torch.equal(r["dlogits"], r2["dlogits"])


@pytest.mark.parametrize("name", ["coco", "hires", "sigma3", "sigma1"])
def test_tma_staged_kernel_matches_ldg_kernel(pb, dev, name):
    """The bulk-async (TMA) staged kernels and the register-staged ones do the same arithmetic per element: dlogits and joints are
    bit-identical; the fp32 partial sums of a map are grouped over different threads, so the loss agrees to fp32 rounding."""
    g = load_golden("sbp_" + name)
    kp, logits, bbox, iid, cid, meta = cases.sbp_case(name)
    x = logits.to(dev)
    scale = meta["input_size"][1] / meta["w"]
    for grad in (True, False):
        for dec in (True, False):
            a = pb.sbp_fused(x, keypoints=kp, sigma=meta["sigma"], want_grad=grad, decode=dec, conf_threshold=0.25, coord_scale=scale, tma=False)
            b = pb.sbp_fused(x, keypoints=kp, sigma=meta["sigma"], want_grad=grad, decode=dec, conf_threshold=0.25, coord_scale=scale, tma=True)
            assert close(b["loss"].item(), a["loss"].item(), 1e-6) and allclose(b["loss_num"], a["loss_num"], 1e-6)
            if grad:
                assert torch.equal(a["dlogits"], b["dlogits"])
            if dec:
                assert torch.equal(a["joints"], b["joints"])
    assert close(b["loss"].item(), float(g["loss"]), REL)


@pytest.mark.parametrize("name", list(cases.SBP_SHAPES))
def test_update_state_rows(pb, dev, name):
    g = load_golden("sbp_" + name)
    kp, logits, bbox, iid, cid, meta = cases.sbp_case(name)
    tgt = {"bbox": bbox, "image_id": iid, "category_id": cid}
    m = pb.SBPmAPCOCO(None, list(meta["input_size"]), 0.25)
    m.update_state(tgt, logits.to(dev))
    assert_rows(m.result_list, golden_rows(g, "rows_coco"), REL)
    m.reset_states()
    assert m.result_list == []
    p = pb.SBPmAPPIS(None, list(meta["input_size"]), 0.25)
    p.update_state({k: (v.to(dev) if k == "bbox" else v) for k, v in tgt.items()}, logits.to(dev))
    assert_rows(p.result_list, golden_rows(g, "rows_pis"), REL)


def test_config1_batch32_matches_reference(pb, dev):
    """BASELINE.json configs[0]: B=32 synthetic batch, every stage against what the reference produced."""
    g = load_golden("sbp_config1")
    kp, logits, bbox, iid, cid = so.make_config1_inputs(32)
    assert cases.digest(kp, logits, bbox) == str(g["inputs_sha"])
    t = pb.SBPHeatmapGenerator([64, 48], 17, 2).render_batch(kp)
    assert float(t.double().sum().item()) == float(g["target_sum"]) and int((t > 0).sum()) == int(g["target_nnz"])
    x = logits.to(dev).requires_grad_(True)
    loss = pb.SBPLoss(sigma=2)(x, torch.from_numpy(kp).to(dev))
    loss.backward()
    assert close(loss.item(), float(g["loss"]), REL)
    assert allclose(x.grad[:2], g["grad_slice"], REL)
    assert close(x.grad.double().abs().sum().item(), float(g["grad_abs_sum"]), REL)
    assert_joints(pb.DecodeSBP([256, 192], 0.25, True).decode_batch(x.detach()), g["joints"], REL)


@pytest.mark.parametrize("sigma", [0.7, 1.1, 1.5, 2.4, 2.5, 4.0])
@pytest.mark.parametrize("hw", [(32, 24), (30, 27), (64, 48), (21, 36), (10, 18), (12, 26)])
def test_fused_padded_template_windows(pb, dev, sigma, hw):
    """The fused kernels fetch targets from a zero-padded shared-memory template with clamped indices (no range tests):
    windows narrower than the template (sigma = 0.7, 1.1, 2.4: 6s+3 is not an integer), half-to-even corners (1.5, 2.5),
    windows clipped by every edge, joints beyond the map, W % 4 != 0 (scalar kernel; with H*W % 4 == 0 the training kernel
    keeps 128-bit vectors that wrap rows) -- target bit-exact, loss/grad 1e-5."""
    h, w = hw
    k = 24
    rng = np.random.default_rng(int(sigma * 10) * 100 + h)
    b = 3
    kp = np.stack([rng.uniform(-2, w + 3, (b, k)), rng.uniform(-2, h + 3, (b, k))], -1)
    kp[0, :8] = [[0, 0], [w - 1, 0], [0, h - 1], [w - 1, h - 1], [w - 0.5, h / 2], [w / 2, h + 2.0], [0.99, 3.2], [3.0, 0.5]]
    kp[1, :3] = -1.0
    logits = torch.randn(b, k, h, w, generator=torch.Generator().manual_seed(h * w)) * 3
    want_t = so.sbp_render(kp, h, w, sigma)
    wl, wg = so.sbp_loss_closed_form_f64(logits, torch.from_numpy(want_t))
    x = logits.to(dev)
    r = pb.sbp_fused(x, keypoints=kp, sigma=sigma, want_grad=True, decode=True, conf_threshold=0.25, coord_scale=4.0, want_target=True)
    assert np.array_equal(r["target"].cpu().numpy(), want_t)
    assert close(r["loss"].item(), float(wl), REL) and allclose(r["dlogits"], wg, REL)
    assert_joints(r["joints"], so.sbp_decode(logits, 4 * w, 0.25, True), REL)
    for grad, dec in ((True, False), (False, True), (False, False)):        # the read-only variants take the branch-free path
        v = pb.sbp_fused(x, keypoints=kp, sigma=sigma, want_grad=grad, decode=dec, conf_threshold=0.25, coord_scale=4.0,
                         want_target=not grad)
        assert close(v["loss"].item(), float(wl), REL)
        if not grad:
            assert np.array_equal(v["target"].cpu().numpy(), want_t)
        if grad:
            assert torch.equal(v["dlogits"], r["dlogits"])
        if dec:
            assert torch.equal(v["joints"], r["joints"])
    if w % 4 == 0:
        t = pb.sbp_fused(x, keypoints=kp, sigma=sigma, want_grad=True, decode=True, conf_threshold=0.25, coord_scale=4.0, tma=True)
        assert torch.equal(t["dlogits"], r["dlogits"]) and close(t["loss"].item(), float(wl), REL)
    # no target out: where the shape allows it these go through the bulk-staged kernels -- the read-only ones in their two-phase
    # form (whole map as zero target, then the window's vectors dealt out to the map's threads)
    for dec in (False, True):
        v = pb.sbp_fused(x, keypoints=kp, sigma=sigma, want_grad=False, decode=dec, conf_threshold=0.25, coord_scale=4.0)
        assert close(v["loss"].item(), float(wl), REL), (v["loss"].item(), float(wl))
        if dec:
            assert torch.equal(v["joints"], r["joints"])


def test_odd_shapes_scalar_path_and_empty_batch(pb, dev):
    """H*W not a multiple of 4 -> scalar kernels; unaligned views; N == 0."""
    k, h, w, sigma = 3, 7, 9, 1
    kp, logits, *_ = so.make_config1_inputs(5, k, h, w, seed=9, torch_seed=9)
    logits = logits * 3
    want_t = so.sbp_render(kp, h, w, sigma)
    assert np.array_equal(pb.SBPHeatmapGenerator([h, w], k, sigma).render_batch(kp).cpu().numpy(), want_t)
    wl, wg = so.sbp_loss_closed_form_f64(logits, torch.from_numpy(want_t))
    x = logits.to(dev).requires_grad_(True)
    loss = pb.SBPLoss(sigma=sigma)(x, torch.from_numpy(kp).to(dev))
    loss.backward()
    assert close(loss.item(), float(wl), REL) and allclose(x.grad, wg, REL)
    assert_joints(pb.decode_batch(logits.to(dev), 0.25, 4.0, True), so.sbp_decode(logits, 4 * w, 0.25, True), REL)
    # a 4-byte-offset view is not 16-byte aligned: dense() keeps it on the device and the kernels still agree
    big = torch.randn(2 * 17 * 64 * 48 + 1, device=dev)
    view = big[1:].view(2, 17, 64, 48)
    assert torch.equal(pb.decode_batch(view, 0.25, 4.0, True), pb.decode_batch(view.clone(), 0.25, 4.0, True))
    empty = torch.zeros(0, 17, 64, 48, device=dev)
    assert tuple(pb.decode_batch(empty, 0.25, 4.0, True).shape) == (0, 17, 3)


def test_quarter_pixel_refinement_unpinned(pb, dev):
    """PARITY UNPINNED: the reference has no refinement; checked against our restatement of the published rule."""
    kp, logits, *_ = so.make_config1_inputs(8, 17, 64, 48, seed=21, torch_seed=21)
    x = so.realistic_logits(kp, 64, 48).to(dev)
    base = pb.decode_batch(x, 0.25, 1.0, True)
    ref = so.sbp_refine_quarter(base.cpu(), torch.sigmoid(x.cpu()))
    got = pb.decode_batch(x, 0.25, 1.0, True, refine=True).cpu()
    assert torch.equal(got[..., 2], base.cpu()[..., 2])
    # sign decisions agree wherever the neighbours differ by more than sigmoid rounding noise
    assert (got[..., :2] - ref[..., :2]).abs().max() <= 0.5
    assert ((got[..., :2] - ref[..., :2]).abs() > 0).float().mean() < 0.02


def test_no_cpu_fallback(pb):
    with pytest.raises(pb.PoseB200Error):
        pb.decode_batch(torch.zeros(1, 1, 8, 8), 0.5)
    with pytest.raises(pb.PoseB200Error):
        pb.SBPLoss()(torch.zeros(1, 1, 8, 8), torch.zeros(1, 1, 8, 8))


def test_cabi_argument_errors(pb, dev):
    L, C = pb.lib(), pb._cabi
    x = torch.zeros(1, 1, 8, 8, device=dev)
    j = torch.zeros(1, 1, 3, device=dev)
    assert L.pose_sbp_decode(None, C.ptr(j), 1, 1, 8, 8, 0.5, 1, 1.0, 0, 1, C.stream_ptr(dev)) == -1
    assert b"NULL" in L.pose_b200_last_error()
    assert L.pose_sbp_decode(C.ptr(x), C.ptr(j), 1, 0, 8, 8, 0.5, 1, 1.0, 0, 1, C.stream_ptr(dev)) == -1
    assert L.pose_sbp_decode(C.ptr(x), C.ptr(j), 1, 1, 8, 8, 0.5, 1, 1.0, 0, 7, C.stream_ptr(dev)) == -1
    assert b"sigmoid_ref" in L.pose_b200_last_error()
    loss = torch.zeros((), device=dev)
    ws = torch.zeros(16, dtype=torch.uint8, device=dev)
    rc = L.pose_sbp_fused(C.ptr(x), C.ptr(x), None, 0, 1.0, None, 0, None, None, C.ptr(loss), None, None, 0.0, 1.0,
                          1, 1, 8, 8, 5.0, 1.0, 0.5, 0, None, None, 0, 0, None, C.ptr(ws), ws.numel(), C.stream_ptr(dev))
    assert rc == -3      # workspace too small
    big_ws = torch.zeros(int(L.pose_sbp_fused_workspace_bytes(1, 1)), dtype=torch.uint8, device=dev)
    bb = torch.zeros(1, 4, dtype=torch.float64, device=dev)
    rc = L.pose_sbp_fused(C.ptr(x), C.ptr(x), None, 0, 1.0, None, 0, None, None, C.ptr(loss), None, C.ptr(j), 0.0, 1.0,
                          1, 1, 8, 8, 5.0, 1.0, 0.5, 4, C.ptr(bb), None, 256, 192, None, C.ptr(big_ws), big_ws.numel(), C.stream_ptr(dev))
    assert rc == -1 and b"bbox and packed_out" in L.pose_b200_last_error()


# ----------------------------------------------------------------------------------------------- full size (config 2)

@pytest.fixture(scope="module")
def big(dev):
    """BASELINE.json configs[1] shapes: B=4096, K=17, 64x48, generated on the device."""
    b, k, h, w = 4096, 17, 64, 48
    gen = torch.Generator(device=dev).manual_seed(0)
    logits = torch.randn(b, k, h, w, device=dev, generator=gen) * 3
    kp = torch.stack([torch.rand(b, k, device=dev, generator=gen, dtype=torch.float64) * w,
                      torch.rand(b, k, device=dev, generator=gen, dtype=torch.float64) * h], dim=-1)
    vis = torch.rand(b, k, device=dev, generator=gen) < 0.85
    kp[~vis] = -1.0
    return logits, kp, vis


def test_full_size_render_decode_roundtrip(pb, dev, big):
    """decode(render(kp), thr .99, pred=False) == (4*int(x), 4*int(y), 1) for visible, (-4,-4,-1) for invisible joints."""
    logits, kp, vis = big
    t = pb.SBPHeatmapGenerator([64, 48], 17, 2).render_batch(kp)
    j = pb.decode_batch(t, 0.99, 4.0, False)
    want = torch.where(vis[..., None], torch.stack([4 * kp[..., 0].floor(), 4 * kp[..., 1].floor(), torch.ones_like(kp[..., 0])], -1),
                       torch.tensor([-4.0, -4.0, -1.0], device=dev, dtype=torch.float64)).float()
    assert torch.equal(j, want)
    # linearity-style checksum: every visible interior joint contributes the same template mass
    interior = vis & (kp[..., 0] >= 8) & (kp[..., 0] < 40) & (kp[..., 1] >= 8) & (kp[..., 1] < 56)
    mass = t.double().sum(dim=(2, 3))
    g = so.gauss_template(2).astype(np.float32).astype(np.float64).sum()
    assert torch.all((mass[interior] - g).abs() < 1e-9) and torch.all(mass[~vis] == 0)


def test_full_size_fused_equals_staged_and_oracle_subset(pb, dev, big):
    logits, kp, vis = big
    b = logits.size(0)
    fused = pb.sbp_fused(logits, keypoints=kp, sigma=2, want_grad=True, decode=True, conf_threshold=0.25, coord_scale=4.0)
    t = pb.SBPHeatmapGenerator([64, 48], 17, 2).render_batch(kp)
    dense_ = pb.sbp_fused(logits, target=t, want_grad=True)
    assert close(fused["loss"].item(), dense_["loss"].item(), 1e-6)
    assert allclose(fused["dlogits"][:64], dense_["dlogits"][:64], 1e-6)
    assert torch.equal(fused["joints"], pb.decode_batch(logits, 0.25, 4.0, True))
    # oracle on a 128-sample subset: loss numerators are additive over samples
    sub = slice(1000, 1128)
    ls, gs = so.sbp_loss_closed_form_f64(logits[sub].cpu(), torch.from_numpy(so.sbp_render(kp[sub].cpu().numpy(), 64, 48, 2)))
    r = pb.sbp_fused(logits[sub], keypoints=kp[sub], sigma=2, want_grad=True, global_batch=b)
    assert close(r["loss"].item() * b / 128, float(ls), REL)
    assert allclose(r["dlogits"] * (b / 128), gs, REL)
    assert allclose(fused["dlogits"][sub] * (b / 128), gs, REL)
    assert_joints(fused["joints"][sub], so.sbp_decode(logits[sub].cpu(), 192, 0.25, True), REL)
    # additivity of the un-normalised numerators over shards (what the multi-GPU all-reduce relies on)
    # (same kernel variant as `fused`: the per-map fp32 partial sums are then bit-identical, only the fp64 grouping differs)
    parts = [pb.sbp_fused(logits[i:i + 1024], keypoints=kp[i:i + 1024], sigma=2, want_grad=True, decode=True, conf_threshold=0.25,
                          coord_scale=4.0)["loss_num"] for i in range(0, b, 1024)]
    tot = torch.stack(parts).sum(0)
    assert allclose(tot, fused["loss_num"], 1e-12)
    ldg = pb.sbp_fused(logits, keypoints=kp, sigma=2, want_grad=True, decode=True, conf_threshold=0.25, coord_scale=4.0, tma=False)
    assert torch.equal(ldg["dlogits"], fused["dlogits"]) and torch.equal(ldg["joints"], fused["joints"])
    assert allclose(ldg["loss_num"], fused["loss_num"], 1e-7)         # register-staged kernel: other grouping of the fp32 partial sums
    other = pb.sbp_fused(logits, keypoints=kp, sigma=2, want_grad=False)["loss_num"]      # another variant: fp32 rounding differs
    assert allclose(other, fused["loss_num"], 1e-7)


def _mismatch_count(pb, x, ref, thr=0.25, chunk=1024):
    """Maps whose decoded row (x, y, conf -- all three bit for bit) differs from the reference's ops on the same logits:
    ref "cpu" = torch.sigmoid on the host, one [1,K,H,W] call per sample as DecodeSBP.forward makes them; "cuda" = on this GPU."""
    got = pb.decode_batch(x, thr, 4.0, True, sigmoid_ref=ref).cpu()
    bad = 0
    for i in range(0, x.size(0), chunk):
        src = x[i:i + chunk].cpu() if ref == "cpu" else x[i:i + chunk]
        want = so.sbp_decode(src, 4 * x.size(-1), thr, True)
        bad += int((_bits(got[i:i + chunk]) != _bits(want)).any(-1).sum())
    return bad


def _record(name, value):
    import json
    import os
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(d):
        path = os.path.join(d, "argmax_mismatch.json")
        rec = json.load(open(path)) if os.path.exists(path) else {}
        rec[name] = value
        json.dump(rec, open(path, "w"), indent=1)


def _big_neartie(dev, b=4096, k=17, h=64, w=48, seed=17):
    """On-device version of cases.sbp_neartie_maps for a full batch: every map has two contenders d = 1..4 ulp apart at a
    magnitude from NEARTIE_MAGNITUDES, the earlier pixel holding the smaller value on even maps."""
    gen = torch.Generator(device=dev).manual_seed(seed)
    x = torch.randn(b * k, h * w, device=dev, generator=gen) * 0.5 - 6.0
    t = torch.arange(b * k, device=dev)
    mags = torch.tensor(cases.NEARTIE_MAGNITUDES, device=dev)[t % len(cases.NEARTIE_MAGNITUDES)]
    lo = mags.clone()
    for d in range(1, 5):
        lo = torch.where((t % 4) + 1 >= d, torch.nextafter(lo, torch.full_like(lo, -1e30)), lo)
    p = torch.rand(b * k, 2, device=dev, generator=gen)
    p1 = (p[:, 0] * (h * w - 2)).long()
    p2 = p1 + 1 + (p[:, 1] * (h * w - 1 - p1).float()).long().clamp(max=h * w - 2)
    p2 = p2.clamp(max=h * w - 1)
    even = (t % 2 == 0)
    x[t, p1] = torch.where(even, lo, mags)
    x[t, p2] = torch.where(even, mags, lo)
    return x.view(b, k, h, w)


@pytest.mark.parametrize("kind", ["randn3", "realistic", "neartie"])
def test_full_config_argmax_mismatches_are_zero(pb, dev, big, kind):
    """BASELINE.json configs[1], all 69 632 maps: the decoded rows equal the reference's (torch.sigmoid -> first row-major
    maximum) bit for bit -- against the host (ATen CPU sigmoid) and against the same ops on this GPU (ATen CUDA sigmoid)."""
    logits, kp, vis = big
    if kind == "randn3":
        x = logits
    elif kind == "realistic":
        t = pb.SBPHeatmapGenerator([64, 48], 17, 2).render_batch(kp)
        g = torch.Generator(device=dev).manual_seed(2)
        x = torch.logit((t + 0.05 * torch.rand(t.shape, device=dev, generator=g)).clamp(1e-4, 1 - 1e-4))
    else:
        x = _big_neartie(dev)
    for ref in ("cuda", "cpu"):
        bad = _mismatch_count(pb, x, ref)
        _record(f"B4096_{kind}_{ref}", {"maps": x.size(0) * x.size(1), "mismatches": bad})
        assert bad == 0, (kind, ref, bad)
    # the fused kernel's decode is the same code: identical rows
    f = pb.sbp_fused(x, keypoints=kp, sigma=2, want_grad=False, decode=True, conf_threshold=0.25, coord_scale=4.0)
    assert torch.equal(f["joints"], pb.decode_batch(x, 0.25, 4.0, True))


def test_config3_shard_size_properties(pb, dev):
    """BASELINE.json configs[2]: B=32768 over 2 GPUs = 16384 images per rank (3.4 GB per tensor: byte offsets beyond 2^31).
    Size-independent properties: render -> decode round trip, fused decode == decode kernel, loss numerators additive over
    4096-image pieces, dlogits of the last piece identical to that piece run alone."""
    b, k, h, w = 16384, 17, 64, 48
    free, _ = torch.cuda.mem_get_info(dev)
    if free < 16 << 30:
        pytest.skip("needs 16 GB of free device memory")
    gen = torch.Generator(device=dev).manual_seed(3)
    logits = torch.randn(b, k, h, w, device=dev, generator=gen) * 3
    kp = torch.stack([torch.rand(b, k, device=dev, generator=gen, dtype=torch.float64) * w,
                      torch.rand(b, k, device=dev, generator=gen, dtype=torch.float64) * h], dim=-1)
    vis = torch.rand(b, k, device=dev, generator=gen) < 0.85
    kp[~vis] = -1.0
    t = pb.SBPHeatmapGenerator([h, w], k, 2).render_batch(kp)
    j = pb.decode_batch(t, 0.99, 4.0, False)
    want = torch.where(vis[..., None], torch.stack([4 * kp[..., 0].floor(), 4 * kp[..., 1].floor(), torch.ones_like(kp[..., 0])], -1),
                       torch.tensor([-4.0, -4.0, -1.0], device=dev, dtype=torch.float64)).float()
    assert torch.equal(j, want)
    del t, j, want
    fused = pb.sbp_fused(logits, keypoints=kp, sigma=2, want_grad=True, decode=True, conf_threshold=0.25, coord_scale=4.0)
    assert torch.equal(fused["joints"], pb.decode_batch(logits, 0.25, 4.0, True))
    for ref in ("cuda", "cpu"):          # all 278 528 maps against the reference's ops
        bad = _mismatch_count(pb, logits, ref)
        _record(f"B16384_randn3_{ref}", {"maps": b * k, "mismatches": bad})
        assert bad == 0, (ref, bad)
    parts = [pb.sbp_fused(logits[i:i + 4096], keypoints=kp[i:i + 4096], sigma=2, want_grad=True, decode=True, conf_threshold=0.25,
                          coord_scale=4.0, global_batch=b) for i in range(0, b, 4096)]
    assert allclose(torch.stack([p["loss_num"] for p in parts]).sum(0), fused["loss_num"], 1e-12)
    assert torch.equal(parts[-1]["dlogits"], fused["dlogits"][-4096:]) and torch.equal(parts[-1]["joints"], fused["joints"][-4096:])
    val = pb.sbp_fused(logits, keypoints=kp, sigma=2, want_grad=False, decode=True, conf_threshold=0.25, coord_scale=4.0)
    assert torch.equal(val["joints"], fused["joints"]) and allclose(val["loss_num"], fused["loss_num"], 1e-7)


@pytest.mark.parametrize("shape", [(6, 17, 64, 48), (3, 11, 64, 48), (2, 17, 33, 27), (2, 5, 16, 20)])
def test_flip_test_decode_unpinned(pb, dev, shape):
    """PARITY UNPINNED: the reference has no flip test; checked against our restatement of the published rule
    (oracle.sbp_oracle.sbp_flip_average) -- indices bit-exact, confidence 1e-5, including the refine option."""
    b, k, h, w = shape
    pairs = [[a, a + 1] for a in range(1, k - 1, 2)]
    rng = np.random.default_rng(b * 1000 + k)
    kp = np.stack([rng.uniform(3, w - 3, (b, k)), rng.uniform(3, h - 3, (b, k))], -1)
    tgt = pb.SBPHeatmapGenerator([h, w], k, 2).render_batch(kp)
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.logit((tgt + 0.05 * torch.rand(tgt.shape, device=dev, generator=g)).clamp(1e-4, 1 - 1e-4))
    # the mirrored pass: mirror + swap of a slightly different prediction
    perm = list(range(k))
    for a, c in pairs:
        perm[a], perm[c] = c, a
    x2 = torch.logit((tgt * 0.9 + 0.05 * torch.rand(tgt.shape, device=dev, generator=g)).clamp(1e-4, 1 - 1e-4))
    xf = x2[:, perm].flip(-1).contiguous()
    for pred in (True, False):
        a_in, f_in = (x, xf) if pred else (torch.sigmoid(x), torch.sigmoid(xf))
        heat = so.sbp_flip_average(a_in.cpu(), f_in.cpu(), pairs, pred)
        want = so.sbp_decode(heat, 4 * w, 0.25, pred=False)
        got = pb.decode_batch(a_in, 0.25, 4.0, pred, flipped=f_in, flip_pairs=pairs)
        assert_joints(got, want, 1e-5)
        base = so.sbp_decode(heat, w, 0.25, pred=False)
        want_r = so.sbp_refine_quarter(base, heat)
        got_r = pb.decode_batch(a_in, 0.25, 1.0, pred, refine=True, flipped=f_in, flip_pairs=pairs)
        assert_joints(got_r, want_r, 1e-5)
    # exact ties across the two passes (pred=False, dyadic values): first row-major index wins
    t = torch.zeros(1, k, h, w, device=dev)
    tf = torch.zeros(1, k, h, w, device=dev)
    t[0, 0, 5, 7] = 0.5; tf[0, 0, 5, w - 1 - 7] = 0.5            # both passes agree -> 0.5
    t[0, 0, 2, 3] = 1.0                                             # one pass only -> also 0.5, earlier index
    got = pb.decode_batch(t, 0.25, 1.0, False, flipped=tf, flip_pairs=pairs).cpu()
    assert got[0, 0].tolist() == [3.0, 2.0, 0.5]
    mod = pb.DecodeSBP([4 * h, 4 * w], 0.25, True, flip_pairs=pairs)
    assert torch.equal(mod(x[:1], xf[:1]), pb.decode_batch(x[:1], 0.25, 4.0, True, flipped=xf[:1], flip_pairs=pairs)[0])
