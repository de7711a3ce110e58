"""Head fusion: `sbp_head_fused` = Conv2d(C, K, 1, bias=False) (models/detector/sbp.py:35-37) + SBPLoss
(models/loss/sbp_loss.py:20-66) + DecodeSBP (utils/sbp_utils.py:97-118) in one kernel, against
  * an fp64 convolution (the yardstick) next to torch's own fp32 conv2d (what the reference computes),
  * the oracle's loss / gradient / decode evaluated on the reference's logits.
Tolerances: loss 1e-5 relative; dlogits 1e-5 of the gradient's max-norm; logits no worse than 4x the fp32 conv2d's own error
(+ 1e-6 of the largest logit); joints identical wherever the reference's top two logits are not within that error.
"""
import numpy as np
import pytest
import torch

from oracle import sbp_oracle as so

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pb():
    import pose_b200
    pose_b200.lib()
    return pose_b200


def _inputs(b, c, k, h, w, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    feats = torch.randn((b, c, h, w), generator=g, device="cuda").relu_() * scale          # post-ReLU, like the deconv stack's output
    weight = torch.randn((k, c), generator=g, device="cuda") * (2.0 / c) ** 0.5
    rng = np.random.default_rng(1234 + seed)
    kp = np.stack([rng.uniform(0, w, (b, k)), rng.uniform(0, h, (b, k))], axis=-1)
    kp[rng.uniform(size=(b, k)) >= 0.85] = -1.0
    return feats, weight, kp


def _reference(feats, weight):
    k, c = weight.shape
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ref32 = torch.nn.functional.conv2d(feats, weight.view(k, c, 1, 1))                      # the reference's head, fp32
    ref64 = torch.einsum("kc,bchw->bkhw", weight.double(), feats.double())
    return ref32, ref64


@pytest.mark.parametrize("b,c,k,h,w,sigma", [(3, 512, 17, 64, 48, 2), (2, 512, 11, 64, 48, 2), (5, 256, 17, 32, 32, 1), (2, 128, 17, 96, 72, 3)])
def test_logits_loss_grad_decode(pb, b, c, k, h, w, sigma):
    feats, weight, kp = _inputs(b, c, k, h, w)
    ref32, ref64 = _reference(feats, weight)
    r = pb.sbp_head_fused(feats, weight, kp, sigma=sigma, want_grad=True, decode=True, conf_threshold=0.25, coord_scale=4.0, want_logits=True)
    torch.cuda.synchronize()
    # ---- the contraction
    err = (r["logits"].double() - ref64).abs().max().item()
    err32 = (ref32.double() - ref64).abs().max().item()
    top = ref64.abs().max().item()
    assert err <= 4 * err32 + 1e-6 * top, (err, err32, top)
    # ---- loss and gradient against the oracle on the reference's (fp32 conv2d) logits, and the fp64 closed form on the fp64 logits
    tgt = so.sbp_render(kp, h, w, sigma)
    want_loss, want_grad = so.sbp_loss_and_grad(ref32.cpu(), torch.from_numpy(tgt))
    assert abs(r["loss"].item() - want_loss.item()) <= 1e-5 * abs(want_loss.item()), (r["loss"].item(), want_loss.item())
    l64, g64 = so.sbp_loss_closed_form_f64(ref64.cpu(), torch.from_numpy(tgt))
    assert abs(r["loss"].item() - float(l64)) <= 1e-5 * float(l64)
    gerr = (r["dlogits"].cpu().double() - g64).abs().max().item()
    assert gerr <= 1e-5 * g64.abs().max().item(), (gerr, g64.abs().max().item())
    assert (r["dlogits"].cpu() - want_grad).abs().max().item() <= 2e-5 * want_grad.abs().max().item()
    # ---- decode: the oracle on the reference's logits; a different pick is allowed only between logits closer than the two
    #      contractions' errors (there the reference's own answer depends on its summation order)
    want_j = so.sbp_decode(ref32.cpu(), w * 4, 0.25, True)
    got_j = r["joints"].cpu()
    flat = ref64.reshape(b * k, -1)
    tol = 2 * (err + err32)
    for m in range(b * k):
        gx, gy = got_j.reshape(-1, 3)[m, :2].tolist()
        wx, wy = want_j.reshape(-1, 3)[m, :2].tolist()
        if (gx, gy) == (wx, wy):
            continue
        assert gx >= 0 and wx >= 0, (m, gx, gy, wx, wy)
        gi, wi = int(gy / 4) * w + int(gx / 4), int(wy / 4) * w + int(wx / 4)
        assert abs(flat[m, gi].item() - flat[m, wi].item()) <= tol, (m, gx, gy, wx, wy)
    conf_g, conf_w = got_j[..., 2], want_j[..., 2]
    assert torch.equal(conf_g < 0, conf_w < 0) or (torch.sigmoid(flat.max(1).values.float().cpu()) - 0.25).abs().min() < 1e-5
    both = (conf_g >= 0) & (conf_w >= 0)
    assert (conf_g[both] - conf_w[both]).abs().max().item() <= 1e-5


def test_variants_agree_and_residual_matters(pb):
    """The shared-memory-residual variant computes the same sums in the same order: bit-identical logits.  Without the feature
    residual the logits carry the 2^-11 truncation of TF32 -- the reason the residual exists."""
    feats, weight, kp = _inputs(4, 512, 17, 64, 48, seed=3)
    _, ref64 = _reference(feats, weight)
    a = pb.sbp_head_fused(feats, weight, kp, sigma=2, want_logits=True, decode=True, coord_scale=4.0)
    s = pb.sbp_head_fused(feats, weight, kp, sigma=2, want_logits=True, decode=True, coord_scale=4.0, tuning=pb.head_tuning(residual_in_smem=True))
    n = pb.sbp_head_fused(feats, weight, kp, sigma=2, want_logits=True, residual=False)
    torch.cuda.synchronize()
    assert torch.equal(a["logits"], s["logits"]) and torch.equal(a["dlogits"], s["dlogits"]) and torch.equal(a["joints"], s["joints"])
    assert a["loss"].item() == s["loss"].item()
    e_full = (a["logits"].double() - ref64).abs().max().item()
    e_trunc = (n["logits"].double() - ref64).abs().max().item()
    assert e_trunc > 50 * e_full, (e_trunc, e_full)


def test_back_projection_rows_equal_unfused_pipeline(pb):
    """packed rows (SBPmAPCOCO.update_state arithmetic) from the fused head == the unfused product path run on the fused kernel's own logits."""
    b, c, k, h, w = 6, 512, 17, 64, 48
    feats, weight, kp = _inputs(b, c, k, h, w, seed=5)
    rng = np.random.default_rng(9)
    bbox = np.stack([rng.uniform(0, 400, b), rng.uniform(0, 400, b), rng.uniform(40, 300, b), rng.uniform(60, 400, b)], axis=-1)
    r = pb.sbp_head_fused(feats, weight, kp, sigma=2, want_grad=False, want_logits=True, conf_threshold=0.25, coord_scale=4.0, bbox=bbox,
                          input_size=(256, 192))
    u = pb.sbp_fused(r["logits"], keypoints=kp, sigma=2, want_grad=True, decode=True, conf_threshold=0.25, coord_scale=4.0, bbox=bbox,
                     input_size=(256, 192))
    torch.cuda.synchronize()
    assert torch.equal(r["joints"][..., :2], u["joints"][..., :2])
    assert (r["joints"][..., 2] - u["joints"][..., 2]).abs().max().item() <= 1e-6
    assert torch.allclose(r["packed"], u["packed"], rtol=1e-6, atol=1e-6)
    assert abs(r["loss"].item() - u["loss"].item()) <= 1e-6 * abs(u["loss"].item())


def test_decode_only_without_keypoints(pb):
    """Inference form (head -> DecodeSBP, no target): the same joints as the training-form call."""
    feats, weight, kp = _inputs(3, 512, 17, 64, 48, seed=13)
    a = pb.sbp_head_fused(feats, weight, kp, sigma=2, decode=True, conf_threshold=0.25, coord_scale=4.0)
    b = pb.sbp_head_fused(feats, weight, None, conf_threshold=0.25, coord_scale=4.0)
    torch.cuda.synchronize()
    assert b["dlogits"] is None and torch.equal(a["joints"], b["joints"])
    ref32, _ = _reference(feats, weight)
    want = so.sbp_decode(ref32.cpu(), 192, 0.25, True)
    assert torch.equal(b["joints"].cpu()[..., :2], want[..., :2])


def test_many_images_per_cta_and_determinism(pb):
    """More images than SMs (several images per persistent CTA, both accumulator buffers, every ring wrap) and run-to-run identical results."""
    b, c, k, h, w = 300, 64, 17, 16, 16
    feats, weight, kp = _inputs(b, c, k, h, w, seed=7)
    ref32, ref64 = _reference(feats, weight)
    r1 = pb.sbp_head_fused(feats, weight, kp, sigma=1, want_logits=True, decode=True)
    r2 = pb.sbp_head_fused(feats, weight, kp, sigma=1, want_logits=True, decode=True)
    torch.cuda.synchronize()
    for key in ("logits", "dlogits", "joints", "loss"):
        assert torch.equal(r1[key], r2[key]), key
    err = (r1["logits"].double() - ref64).abs().max().item()
    err32 = (ref32.double() - ref64).abs().max().item()
    assert err <= 4 * err32 + 1e-6 * ref64.abs().max().item()
    tgt = so.sbp_render(kp, h, w, 1)
    l64, _ = so.sbp_loss_closed_form_f64(ref64.cpu(), torch.from_numpy(tgt))
    assert abs(r1["loss"].item() - float(l64)) <= 1e-5 * float(l64)


def test_autograd_matches_conv_plus_sbploss(pb):
    """HeadFusedSBPLoss(features, weight, kp).backward() == the reference's training step head(features) -> SBPLoss -> backward()."""
    b, c, k, h, w = 4, 128, 17, 64, 48
    feats, weight, kp = _inputs(b, c, k, h, w, seed=11)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    f1 = feats.clone().requires_grad_(True)
    w1 = weight.clone().view(k, c, 1, 1).requires_grad_(True)
    loss1 = pb.HeadFusedSBPLoss(sigma=2)(f1, w1, kp)
    loss1.backward()
    f2 = feats.clone().requires_grad_(True)
    w2 = weight.clone().view(k, c, 1, 1).requires_grad_(True)
    tgt = torch.from_numpy(so.sbp_render(kp, h, w, 2)).cuda()
    loss2 = so.sbp_loss(torch.nn.functional.conv2d(f2, w2), tgt)            # the reference's op chain, on the device
    loss2.backward()
    torch.cuda.synchronize()
    assert abs(loss1.item() - loss2.item()) <= 1e-5 * abs(loss2.item())
    assert (w1.grad - w2.grad).abs().max().item() <= 2e-5 * w2.grad.abs().max().item()
    assert (f1.grad - f2.grad).abs().max().item() <= 2e-5 * f2.grad.abs().max().item()


def test_rejects_unsupported_shapes(pb):
    feats, weight, kp = _inputs(1, 48, 17, 64, 48)                    # C not a multiple of 32
    with pytest.raises(pb.PoseB200Error):
        pb.sbp_head_fused(feats, weight, kp, sigma=2)
    feats, weight, kp = _inputs(1, 64, 17, 10, 10)                    # H*W not a multiple of 128
    with pytest.raises(pb.PoseB200Error):
        pb.sbp_head_fused(feats, weight, kp, sigma=2)
    feats, weight, kp = _inputs(1, 64, 20, 16, 16)                    # more joints than the epilogue holds
    with pytest.raises(pb.PoseB200Error):
        pb.sbp_head_fused(feats, weight, kp, sigma=2)
    with pytest.raises(pb.PoseB200Error):
        pb.sbp_head_fused(feats.cpu(), weight, kp, sigma=2)
