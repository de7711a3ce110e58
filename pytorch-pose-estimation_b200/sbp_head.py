"""Head fusion: the detector's last layer fused with the heat-map loss / gradient / decode.

The reference's SBP head is `nn.Conv2d(512, num_keypoints, 1, 1, bias=False)` (models/detector/sbp.py:35-37) and its output goes
straight into `SBPLoss` (models/loss/sbp_loss.py:20-66) in training and into `DecodeSBP` (utils/sbp_utils.py:97-118) in
validation.  `sbp_head_fused(features, weight, keypoints, ...)` computes the convolution on the tensor cores (tcgen05, TF32 with
both operands split so the logits carry fp32-level error) and applies loss / dL/dlogits / argmax to the accumulators as they
leave TMEM: the logits never exist in HBM.

`HeadFusedSBPLoss` is the training-loop form: `loss = crit(features, head.weight, keypoints)`; backward() turns the saved
dL/dlogits into the gradients of the features and of the weight with two library GEMMs (torch.matmul -- the conv's own backward,
outside the hot path).
"""
import torch
from torch import nn

from . import _cabi
from ._cabi import check, dense, lib, ptr, stream_ptr
from .sbp_loss import scale_grad_
from .sbp_utils import _gauss_template, _kp_tensor, _templates


def head_tuning(raw_stages=0, residual_stages=0, residual_in_smem=False):
    """Pack the `tuning` word of pose_sbp_head_fused (0 = library defaults: residuals through tensor memory)."""
    return (int(raw_stages) & 0xff) | ((int(residual_stages) & 0xff) << 8) | ((1 << 24) if residual_in_smem else 0)


def sbp_head_fused(features, weight, keypoints=None, sigma=-1, want_grad=True, decode=False, conf_threshold=0.25, coord_scale=1.0,
                   lambda_positive=5.0, lambda_negative=1.0, global_batch=None, bbox=None, input_size=None, want_logits=False,
                   residual=True, sigmoid_ref=None, tuning=0, out=None):
    """features [B,C,H,W] fp32 (NCHW), weight [K,C] or [K,C,1,1] fp32, keypoints [B,K,2] in heat-map pixels.

    Returns dict(loss, loss_num, dlogits, logits, joints, packed) like `sbp_fused`; `logits` only with `want_logits`
    (tests).  `residual=False` drops the feature residual (plain TF32 features; diagnostics).  `keypoints=None` is the inference
    form (the head followed by DecodeSBP, inference_sbp.py:57-58,73): only `joints` / `packed` are meaningful then."""
    x = dense(features, "features")
    assert x.dim() == 4, "features must be [B,C,H,W]"
    b, c, h, w = x.shape
    wt = dense(weight, "weight").reshape(weight.shape[0], -1)
    k = wt.shape[0]
    assert wt.shape[1] == c, "weight must be [K,C] / [K,C,1,1]"
    dev = x.device
    sig = float(h / 64 if sigma < 0 else sigma)
    kp = lut = None
    lut_n, kp_dtype = 0, _cabi.KP_F64
    if keypoints is not None:
        kp = _kp_tensor(keypoints, dev)
        assert tuple(kp.shape) == (b, k, 2), "keypoints must be [B,K,2]"
        g = _gauss_template(sig)
        lut, lut_n = _templates.get(g, sig, dev), g.shape[0]
        kp_dtype = _cabi.KP_F64 if kp.dtype == torch.float64 else _cabi.KP_F32
    else:                                    # inference: head -> decode, no target
        want_grad, decode = False, True
    out = out or {}
    flags = 0
    dlogits = logits = joints = packed = bb = None
    if want_grad:
        flags |= _cabi.F_GRAD
        dlogits = out.get("dlogits")
        if dlogits is None:
            dlogits = torch.empty((b, k, h, w), dtype=torch.float32, device=dev)
    if want_logits:
        flags |= _cabi.F_HEAD_LOGITS_OUT
        logits = torch.empty((b, k, h, w), dtype=torch.float32, device=dev)
    if not residual:
        flags |= _cabi.F_HEAD_NO_RESIDUAL
    if decode or bbox is not None:
        flags |= _cabi.F_DECODE
        if _cabi.sigmoid_ref_code(sigmoid_ref) == _cabi.SIGMOID_ATEN_CUDA:
            flags |= _cabi.F_SIGMOID_CUDA
        joints = out.get("joints")
        if joints is None:
            joints = torch.empty((b, k, 3), dtype=torch.float32, device=dev)
    in_h = in_w = 0
    if bbox is not None:
        bb = dense(bbox.to(dev) if isinstance(bbox, torch.Tensor) else torch.as_tensor(bbox).to(dev), "bbox", torch.float64)
        in_h, in_w = int(input_size[0]), int(input_size[1])
        packed = out.get("packed")
        if packed is None:
            packed = torch.empty((b, 3 * k + 1), dtype=torch.float32, device=dev)
    loss = out.get("loss")
    if loss is None:
        loss = torch.empty((), dtype=torch.float32, device=dev)
    num = out.get("loss_num")
    if num is None:
        num = torch.empty((2,), dtype=torch.float64, device=dev)
    nbytes = int(lib().pose_sbp_head_workspace_bytes(b, k, c))
    ws = _cabi.workspace(dev, nbytes)
    inv_norm = 1.0 / (2.0 * k * (global_batch if global_batch is not None else b)) if b > 0 else 0.0
    with torch.cuda.device(dev):
        check(lib().pose_sbp_head_fused(ptr(x), ptr(wt), ptr(kp), kp_dtype, sig, ptr(lut), lut_n, ptr(dlogits), ptr(logits), ptr(loss),
                                        ptr(num), ptr(joints), float(conf_threshold), float(coord_scale), b, c, k, h, w,
                                        float(lambda_positive), float(lambda_negative), inv_norm, flags, ptr(bb), ptr(packed), in_h, in_w,
                                        int(tuning), ptr(ws), ws.numel(), stream_ptr(dev)), "pose_sbp_head_fused")
    return dict(loss=loss, loss_num=num, dlogits=dlogits, logits=logits, joints=joints, packed=packed)


class _HeadFusedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, weight, keypoints, sigma, lp, ln, global_batch):
        need = bool(ctx.needs_input_grad[0] or ctx.needs_input_grad[1])
        r = sbp_head_fused(features, weight, keypoints, sigma, want_grad=need, lambda_positive=lp, lambda_negative=ln,
                           global_batch=global_batch)
        ctx.dlogits = r["dlogits"]
        ctx.save_for_backward(features, weight)
        return r["loss"]

    @staticmethod
    def backward(ctx, grad_output):
        d = ctx.dlogits
        if d is None:
            raise RuntimeError("HeadFusedSBPLoss: backward() called but the forward pass ran without requires_grad")
        ctx.dlogits = None
        features, weight = ctx.saved_tensors
        d = scale_grad_(d, grad_output)
        b, k, h, w = d.shape
        c = features.shape[1]
        d2 = d.view(b, k, h * w)
        gf = gw = None
        if ctx.needs_input_grad[0]:      # dX[b] = W^T dlogits[b]
            gf = torch.matmul(weight.reshape(k, c).t().unsqueeze(0), d2).view(b, c, h, w)
        if ctx.needs_input_grad[1]:      # dW = sum_b dlogits[b] X[b]^T
            gw = torch.einsum("bkp,bcp->kc", d2, features.reshape(b, c, h * w)).view(weight.shape)
        return gf, gw, None, None, None, None, None


class HeadFusedSBPLoss(nn.Module):
    """`SBPLoss()(head(features), target)` of the reference's training step (module/sbp_detector.py:24 with
    models/detector/sbp.py:47) as ONE call on the head's input: forward(features, weight, keypoints[B,K,2]) -> 0-dim loss."""

    def __init__(self, sigma=-1, global_batch=None):
        super().__init__()
        self.lambda_positive = 5
        self.lambda_negative = 1
        self.sigma = sigma
        self.global_batch = global_batch

    def forward(self, features, weight, keypoints):
        return _HeadFusedFn.apply(features, weight, keypoints, self.sigma, float(self.lambda_positive), float(self.lambda_negative),
                                  self.global_batch)
