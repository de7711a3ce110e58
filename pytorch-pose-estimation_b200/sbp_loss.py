"""SBPLoss with the reference's signature (models/loss/sbp_loss.py:9-66) on the fused sm_100a kernel.

forward(input, target) accepts either the reference's dense target [B,K,H,W] or -- the fused path --
the keypoints [B,K,2] themselves, in which case the Gaussian target is rendered in registers and
never exists in HBM.  The gradient w.r.t. the logits is produced by the same pass that computes the
loss; backward() only applies grad_output (a no-op launch when it is 1).
"""
import ctypes
import os

import torch
from torch import nn

from . import _cabi
from ._cabi import check, dense, lib, ptr, stream_ptr
from .sbp_utils import _gauss_template, _kp_tensor, _templates


DEFAULT_TMA = _cabi.DEFAULT_TMA


def sbp_fused(logits, target=None, keypoints=None, sigma=-1, want_grad=True, decode=False, conf_threshold=0.25,
              coord_scale=1.0, want_target=False, lambda_positive=5.0, lambda_negative=1.0, global_batch=None,
              bbox=None, input_size=None, out=None, tma=None, exchange=None, sigmoid_ref=None):
    """One pass over the logits: loss (+dlogits) (+rendered target) (+decoded joints).

    Returns dict(loss 0-dim fp32, loss_num fp64[2] = un-normalised (S_pos, S_neg), dlogits, target, joints, packed);
    entries not requested are None.  `global_batch` (default: local B) sets the 1/(2*K*B) normalisation so
    image shards on several GPUs produce gradients of the global-batch loss.  With `bbox` [B,4] fp64 and
    `input_size` (H_in, W_in) the same call also back-projects the decoded joints into `packed` [B,3K+1]
    (SBPmAPCOCO.update_state arithmetic).  `out` may carry preallocated `dlogits`, `joints`, `packed`, `loss`,
    `loss_num` tensors (e.g. views into a communication buffer).  `exchange` (a `dist.PeerExchange`) makes the
    epilogue store the rows / loss numerators / ids into every rank's receive buffer over NVLink.  `sigmoid_ref` as in
    `decode_batch`.
    """
    x = dense(logits, "input")
    assert x.dim() == 4, "input must be [B,K,H,W]"
    b, k, h, w = x.shape
    dev = x.device
    if (target is None) == (keypoints is None):
        raise ValueError("pass exactly one of target / keypoints")
    flags = 0
    t_in = kp = lut = None
    lut_n, kp_dtype, sig = 0, 0, 1.0
    if target is not None:
        t_in = dense(target, "target")
        assert t_in.shape == x.shape, "target shape must equal input shape"
    else:
        sig = float(h / 64 if sigma < 0 else sigma)
        kp = _kp_tensor(keypoints, dev)
        assert tuple(kp.shape) == (b, k, 2), "keypoints must be [B,K,2]"
        g = _gauss_template(sig)
        lut, lut_n = _templates.get(g, sig, dev, padded=True), g.shape[0]
        kp_dtype = _cabi.KP_F64 if kp.dtype == torch.float64 else _cabi.KP_F32
    out = out or {}
    dlogits = t_out = joints = packed = bb = None
    if want_grad:
        flags |= _cabi.F_GRAD
        dlogits = out.get("dlogits")
        if dlogits is None:
            dlogits = torch.empty_like(x)
    if want_target and kp is not None:
        flags |= _cabi.F_TARGET_OUT
        t_out = torch.empty_like(x)
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
        if packed is None and exchange is None:
            packed = torch.empty((b, 3 * k + 1), dtype=torch.float32, device=dev)
    if tma if tma is not None else DEFAULT_TMA:
        flags |= _cabi.F_TMA           # honoured by the library only where it applies (render mode, aligned, no target out)
    loss = out.get("loss")
    if loss is None:
        loss = torch.empty((), dtype=torch.float32, device=dev)
    num = out.get("loss_num")
    if num is None:
        num = torch.empty((2,), dtype=torch.float64, device=dev)
    nbytes = int(lib().pose_sbp_fused_workspace_bytes(b, k))
    ws = _cabi.workspace(dev, nbytes)
    inv_norm = 1.0 / (2.0 * k * (global_batch if global_batch is not None else b)) if b > 0 else 0.0
    with torch.cuda.device(dev):
        check(lib().pose_sbp_fused(ptr(x), ptr(t_in), ptr(kp), kp_dtype, sig, ptr(lut), lut_n, ptr(dlogits), ptr(t_out),
                                   ptr(loss), ptr(num), ptr(joints), float(conf_threshold), float(coord_scale),
                                   b, k, h, w, float(lambda_positive), float(lambda_negative), inv_norm, flags,
                                   ptr(bb), ptr(packed), in_h, in_w,
                                   ctypes.byref(exchange.desc) if exchange is not None else None,
                                   ptr(ws), ws.numel(), stream_ptr(dev)), "pose_sbp_fused")
    return dict(loss=loss, loss_num=num, dlogits=dlogits, target=t_out, joints=joints, packed=packed)


def scale_grad_(dlogits, grad_output):
    """dlogits *= grad_output (device scalar) without a host sync; the kernel exits early when it is 1."""
    g = dense(grad_output.reshape(1), "grad_output")
    with torch.cuda.device(dlogits.device):
        check(lib().pose_scale_grad(ptr(dlogits), ptr(g), dlogits.numel(), stream_ptr(dlogits.device)), "pose_scale_grad")
    return dlogits


class _SBPLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, is_kp, sigma, lp, ln, global_batch):
        need = bool(ctx.needs_input_grad[0])     # grad mode is off inside Function.forward; this is the caller's view
        r = sbp_fused(logits, None if is_kp else target, target if is_kp else None, sigma, want_grad=need,
                      lambda_positive=lp, lambda_negative=ln, global_batch=global_batch)
        ctx.dlogits = r["dlogits"]
        ctx.in_dtype, ctx.in_shape = logits.dtype, logits.shape
        return r["loss"]

    @staticmethod
    def backward(ctx, grad_output):
        d = ctx.dlogits
        if d is None:
            raise RuntimeError("SBPLoss: backward() called but the forward pass ran without requires_grad")
        ctx.dlogits = None                      # single use: the buffer is scaled in place and handed to autograd
        d = scale_grad_(d, grad_output)
        if d.dtype != ctx.in_dtype:
            d = d.to(ctx.in_dtype)
        return d.view(ctx.in_shape), None, None, None, None, None, None


class SBPLoss(nn.Module):
    """Simple Baseline Pose-Estimation loss: drop-in for models/loss/sbp_loss.py:9-66 (stateless, no parameters).

    `forward(input, target)`: input [B,K,H,W] logits; target either [B,K,H,W] heat maps (reference contract)
    or [B,K,2] keypoints in heat-map pixels (fused render+loss; `sigma` as in SBPHeatmapGenerator).
    """

    def __init__(self, sigma=-1, global_batch=None):
        super().__init__()
        self.lambda_positive = 5
        self.lambda_negative = 1
        self.sigma = sigma
        self.global_batch = global_batch

    def forward(self, input, target):
        is_kp = target.dim() == 3 and target.size(-1) == 2
        if not is_kp and not target.is_cuda and input.is_cuda:
            target = target.to(input.device, non_blocking=True)     # the reference moves the target too (:36-39)
        return _SBPLossFn.apply(input, target, is_kp, self.sigma, float(self.lambda_positive),
                                float(self.lambda_negative), self.global_batch)
