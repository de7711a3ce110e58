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
    ws = _cabi.workspace(dev, int(lib().pose_spm_loss_workspace_bytes()))
    inv_norm = 1.0 / (global_batch if global_batch is not None else n) if n > 0 else 0.0
    with torch.cuda.device(dev):
        check(lib().pose_spm_loss(ptr(x), ptr(t), ptr(dlogits), ptr(loss), ptr(num), n, k, r, float(lambda_root),
                                  float(lambda_disp), inv_norm, int(want_grad), ptr(ws), ws.numel(), stream_ptr(dev)),
              "pose_spm_loss")
    return dict(loss=loss, loss_num=num, dlogits=dlogits)


class _SPMLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, lr, ld, global_batch):
        need = bool(ctx.needs_input_grad[0])     # grad mode is off inside Function.forward; this is the caller's view
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
        return d.view(ctx.in_shape), None, None, None, None


class SPMLoss(nn.Module):
    """SPM loss: drop-in for models/loss/spm_loss.py:9-105 (root sigmoid-MSE + 0.1 x SmoothL1 of tanh displacements,
    both masked by the root target's support, / batch)."""

    def __init__(self, global_batch=None):
        super().__init__()
        self.lambda_root = 1
        self.lambda_root_negative = 1
        self.lambda_disp = 0.1
        self.lambda_disp_negative = 1
        self.global_batch = global_batch

    def forward(self, input, target):
        if not target.is_cuda and input.is_cuda:
            target = target.to(input.device, non_blocking=True)
        return _SPMLossFn.apply(input, target, float(self.lambda_root), float(self.lambda_disp), self.global_batch)
