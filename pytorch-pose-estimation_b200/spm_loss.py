"""SPMLoss with the reference's signature (models/loss/spm_loss.py:9-105) on one fused sm_100a kernel."""
import torch
from torch import nn

from . import _cabi
from ._cabi import check, dense, lib, ptr, stream_ptr
from .sbp_loss import scale_grad_


def spm_loss_fused(logits, target, want_grad=True, lambda_root=1.0, lambda_disp=0.1, global_batch=None):
    """loss (+dlogits) in one pass.  Returns dict(loss, loss_num fp64[2] = (S_root, S_disp), dlogits)."""
    x = dense(logits, "input")
    t = dense(target, "target")
    assert x.dim() == 4 and x.shape == t.shape and x.size(2) == x.size(3) and x.size(1) % 2 == 1
    n, c, r, _ = x.shape
    k = (c - 1) // 2
    dev = x.device
    dlogits = torch.empty_like(x) if want_grad else None
    loss = torch.empty((), dtype=torch.float32, device=dev)
    num = torch.empty((2,), dtype=torch.float64, device=dev)
    ws = _cabi.workspace(dev, int(lib().pose_spm_loss_workspace_bytes(n, k, r)))
    inv_norm = 1.0 / (global_batch if global_batch is not None else n) if n > 0 else 0.0
    with torch.cuda.device(dev):
        check(lib().pose_spm_loss(ptr(x), ptr(t), ptr(dlogits), ptr(loss), ptr(num), n, k, r, float(lambda_root),
                                  float(lambda_disp), inv_norm, int(want_grad), ptr(ws), ws.numel(), stream_ptr(dev)),
              "pose_spm_loss")
    return dict(loss=loss, loss_num=num, dlogits=dlogits)


def spm_fused(logits, centers, joints, counts, sigma=-1, want_grad=True, want_target=False, lambda_root=1.0, lambda_disp=0.1,
              global_batch=None):
    """Render + loss (+dlogits) in ONE pass from the persons themselves: the [N,1+2K,R,R] target never touches HBM.

    centers [N,Pmax,2] i64, joints [N,Pmax,K,2] i64, counts [N] i32 (as for `spm_render_batch`; Pmax <= 64).
    Returns dict(loss, loss_num fp64[2] = (S_root, S_disp), dlogits, target) -- `target` only with want_target
    (bit-identical to `spm_render_batch`).
    """
    from .spm_utils import _i64, _render_args
    x = dense(logits, "input")
    assert x.dim() == 4 and x.size(2) == x.size(3) and x.size(1) % 2 == 1
    n, c, r, _ = x.shape
    k = (c - 1) // 2
    dev = x.device
    cen, jnt, cnt, sig, lut, lut_n = _render_args(centers, joints, counts, r, sigma, dev)
    assert cen.size(0) == n and jnt.size(2) == k, "persons do not match the logits' batch / joint count"
    dlogits = torch.empty_like(x) if want_grad else None
    target = torch.empty_like(x) if want_target else None
    loss = torch.empty((), dtype=torch.float32, device=dev)
    num = torch.empty((2,), dtype=torch.float64, device=dev)
    ws = _cabi.workspace(dev, int(lib().pose_spm_fused_workspace_bytes(n, k, r)))
    inv_norm = 1.0 / (global_batch if global_batch is not None else n) if n > 0 else 0.0
    flags = (_cabi.F_GRAD if want_grad else 0) | (_cabi.F_TARGET_OUT if want_target else 0)
    with torch.cuda.device(dev):
        check(lib().pose_spm_fused(ptr(x), ptr(cen), ptr(jnt), ptr(cnt), ptr(dlogits), ptr(target), ptr(loss), ptr(num),
                                   n, cen.size(1), k, r, sig, ptr(lut), lut_n, float(lambda_root), float(lambda_disp), inv_norm,
                                   flags, ptr(ws), ws.numel(), stream_ptr(dev)), "pose_spm_fused")
    return dict(loss=loss, loss_num=num, dlogits=dlogits, target=target)


MAX_FUSED_PERSONS = 64      # pose_spm_fused keeps one 64-bit person mask per map row


class _SPMLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, lr, ld, global_batch, persons=None):
        need = bool(ctx.needs_input_grad[0])     # grad mode is off inside Function.forward; this is the caller's view
        if persons is not None:
            r = spm_fused(logits, persons[0], persons[1], persons[2], persons[3], need, False, lr, ld, global_batch)
        else:
            r = spm_loss_fused(logits, target, need, lr, ld, global_batch)
        ctx.dlogits = r["dlogits"]
        ctx.in_dtype, ctx.in_shape = logits.dtype, logits.shape
        return r["loss"]

    @staticmethod
    def backward(ctx, grad_output):
        d = ctx.dlogits
        if d is None:
            raise RuntimeError("SPMLoss: backward() called but the forward pass ran without requires_grad")
        ctx.dlogits = None
        d = scale_grad_(d, grad_output)
        if d.dtype != ctx.in_dtype:
            d = d.to(ctx.in_dtype)
        return d.view(ctx.in_shape), None, None, None, None, None


class SPMLoss(nn.Module):
    """SPM loss: drop-in for models/loss/spm_loss.py:9-105 (root sigmoid-MSE + 0.1 x SmoothL1 of tanh displacements,
    both masked by the root target's support, / batch)."""

    def __init__(self, global_batch=None, sigma=-1):
        super().__init__()
        self.lambda_root = 1
        self.lambda_root_negative = 1
        self.lambda_disp = 0.1
        self.lambda_disp_negative = 1
        self.global_batch = global_batch
        self.sigma = sigma               # only used by the persons hand-off below (the reference's loss never renders)

    def forward(self, input, target):
        """`target`: the dense [B,1+2K,R,R] tensor (reference signature), or the persons it would be rendered from --
        a dict / tuple of (centers [B,Pmax,2] i64, joints [B,Pmax,K,2] i64, counts [B] i32): the keypoints-only hand-off
        (SURVEY 8 f-1), rendered in registers by the fused kernel.  More than 64 persons per image: render, then dense loss."""
        if isinstance(target, dict):
            target = (target["centers"], target["joints"], target["counts"])
        if isinstance(target, (tuple, list)):
            centers, joints, counts = target
            if centers.shape[1] > MAX_FUSED_PERSONS:
                from .spm_utils import spm_render_batch
                target = spm_render_batch(centers, joints, counts, input.size(-1), self.sigma, device=input.device)
            else:
                return _SPMLossFn.apply(input, None, float(self.lambda_root), float(self.lambda_disp), self.global_batch,
                                        (centers, joints, counts, self.sigma))
        if not target.is_cuda and input.is_cuda:
            target = target.to(input.device, non_blocking=True)
        return _SPMLossFn.apply(input, target, float(self.lambda_root), float(self.lambda_disp), self.global_batch)
