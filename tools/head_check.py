"""Head-fusion diagnostics on the GPU: accuracy of the tcgen05 3xTF32 contraction against an fp64 convolution (next to torch's fp32
conv2d), loss / gradient / decode against fp64 closed forms evaluated with torch on the fp64 logits (the parity tests proper,
against the oracle, are tests/test_head_gpu.py), timing against conv2d + the fused loss kernel.  Prints JSON lines."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pose_b200 as pb                                     # noqa: E402


def make_inputs(b, c, k, h, w, seed=0, dev="cuda"):
    g = torch.Generator(device=dev).manual_seed(seed)
    feats = torch.randn((b, c, h, w), generator=g, device=dev).relu_()            # post-ReLU features, like the reference's deconv stack
    weight = torch.randn((k, c), generator=g, device=dev) * (2.0 / c) ** 0.5
    rng = np.random.default_rng(1234 + seed)
    kp = np.stack([rng.uniform(0, w, (b, k)), rng.uniform(0, h, (b, k))], axis=-1)
    kp[rng.uniform(size=(b, k)) >= 0.85] = -1.0
    return feats, weight, kp


def accuracy(b, c, k, h, w, tuning=0):
    feats, weight, kp = make_inputs(b, c, k, h, w)
    ref64 = torch.einsum("kc,bchw->bkhw", weight.double(), feats.double())
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    conv32 = torch.nn.functional.conv2d(feats, weight.view(k, c, 1, 1))
    out = {"shape": [b, c, k, h, w], "tuning": tuning}
    for name, residual in (("3xtf32", True), ("tf32_features_raw", False)):
        r = pb.sbp_head_fused(feats, weight, kp, sigma=2, want_grad=True, decode=True, conf_threshold=0.25, coord_scale=4.0, want_logits=True,
                              residual=residual, tuning=tuning)
        torch.cuda.synchronize()
        err = (r["logits"].double() - ref64).abs()
        out[name] = {"max_abs_err": float(err.max()), "mean_abs_err": float(err.mean()), "ref_absmax": float(ref64.abs().max())}
        if residual:
            res = r
    e32 = (conv32.double() - ref64).abs()
    out["torch_conv2d_fp32"] = {"max_abs_err": float(e32.max()), "mean_abs_err": float(e32.mean())}
    # emulations of the no-residual result: features truncated vs rounded to tf32 (which one does the tensor core do?)
    ft = (feats.view(torch.int32) & -8192).view(torch.float32)
    emu_trunc = torch.einsum("kc,bchw->bkhw", weight.double(), ft.double())
    r0 = pb.sbp_head_fused(feats, weight, kp, sigma=2, want_logits=True, residual=False, tuning=tuning)
    out["no_residual_vs_truncated_features"] = float((r0["logits"].double() - emu_trunc).abs().max())
    # loss / grad / decode against fp64 closed forms on the fp64 logits (target rendered by the render kernel, bit-exact vs the oracle)
    tgt = pb.SBPHeatmapGenerator([h, w], k, 2).render_batch(kp).double()
    sg = torch.sigmoid(ref64)
    pos = tgt > 0
    wl = (5.0 * ((sg - tgt)[pos] ** 2).sum() + ((sg - tgt)[~pos] ** 2).sum() + 5.0 * (tgt[~pos] ** 2).sum()) / (2 * k * b)
    wg = torch.where(pos, 5.0, 1.0) * 2.0 * (sg - tgt) * sg * (1 - sg) / (2 * k * b)
    out["loss"] = {"got": float(res["loss"]), "want": float(wl), "rel": abs(float(res["loss"]) - float(wl)) / abs(float(wl))}
    gerr = (res["dlogits"].double() - wg).abs().max()
    out["grad"] = {"max_abs_err": float(gerr), "rel_to_maxnorm": float(gerr / wg.abs().max())}
    flat = ref64.reshape(b * k, -1)
    idx = flat.argmax(1)
    conf = torch.sigmoid(flat.gather(1, idx[:, None]))[:, 0]
    wx = torch.where(conf > 0.25, (idx % w).double() * 4.0, torch.full_like(conf, -4.0))
    wy = torch.where(conf > 0.25, (idx // w).double() * 4.0, torch.full_like(conf, -4.0))
    gj = res["joints"].reshape(b * k, 3).double()
    out["decode"] = {"coord_mismatch": int(((gj[:, 0] != wx) | (gj[:, 1] != wy)).sum()), "maps": b * k,
                     "conf_max_err": float((gj[:, 2] - torch.where(conf > 0.25, conf, torch.full_like(conf, -1.0))).abs().max())}
    return out


def timing(b, c, k, h, w, iters, tuning=0):
    feats, weight, kp = make_inputs(b, c, k, h, w)
    kpt = torch.from_numpy(kp).cuda()
    out = {"shape": [b, c, k, h, w], "tuning": tuning, "feature_bytes": feats.numel() * 4}
    dl = torch.empty((b, k, h, w), device="cuda")
    jo = torch.empty((b, k, 3), device="cuda")

    def run(fn, n):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    algo_train = feats.numel() * 4 + dl.numel() * 4
    algo_val = feats.numel() * 4
    t = run(lambda: pb.sbp_head_fused(feats, weight, kpt, sigma=2, want_grad=True, decode=True, conf_threshold=0.25, coord_scale=4.0,
                                      out={"dlogits": dl, "joints": jo}, tuning=tuning), iters)
    out["head_fused_grad_decode_ms"] = t
    out["head_fused_grad_decode_GBps"] = algo_train / t / 1e6
    t = run(lambda: pb.sbp_head_fused(feats, weight, kpt, sigma=2, want_grad=False, decode=True, conf_threshold=0.25, coord_scale=4.0,
                                      out={"joints": jo}, tuning=tuning), iters)
    out["head_fused_val_ms"] = t
    out["head_fused_val_GBps"] = algo_val / t / 1e6
    t = run(lambda: pb.sbp_head_fused(feats, weight, kpt, sigma=2, want_grad=False, decode=True, conf_threshold=0.25, coord_scale=4.0,
                                      out={"joints": jo}, residual=False, tuning=tuning), iters)
    out["head_fused_val_no_residual_ms"] = t
    # the unfused pair: library conv (fp32, no TF32) + our fused loss kernel on its output
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    w4 = weight.view(k, c, 1, 1)
    t_conv = run(lambda: torch.nn.functional.conv2d(feats, w4), iters)
    logits = torch.nn.functional.conv2d(feats, w4)
    t_loss = run(lambda: pb.sbp_fused(logits, keypoints=kpt, sigma=2, want_grad=True, decode=True, conf_threshold=0.25, coord_scale=4.0,
                                      out={"dlogits": dl, "joints": jo}), iters)
    t_pair = run(lambda: pb.sbp_fused(torch.nn.functional.conv2d(feats, w4), keypoints=kpt, sigma=2, want_grad=True, decode=True,
                                      conf_threshold=0.25, coord_scale=4.0, out={"dlogits": dl, "joints": jo}), iters)
    out["conv2d_fp32_ms"] = t_conv
    out["sbp_fused_on_logits_ms"] = t_loss
    out["conv2d_then_sbp_fused_ms"] = t_pair
    torch.backends.cudnn.allow_tf32 = True
    t_conv_tf32 = run(lambda: torch.nn.functional.conv2d(feats, w4), iters)
    out["conv2d_tf32_ms"] = t_conv_tf32
    torch.backends.cudnn.allow_tf32 = False
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--acc", type=int, nargs="*", default=[2, 8])
    ap.add_argument("--time", type=int, nargs="*", default=[])
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--tuning", type=int, nargs="*", default=[0])
    ap.add_argument("--c", type=int, default=512)
    ap.add_argument("--k", type=int, default=17)
    ap.add_argument("--out", default="gpurun_out/head_check.jsonl")
    a = ap.parse_args()
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    with open(a.out, "a") as f:
        for tn in a.tuning:
            for b in a.acc:
                r = accuracy(b, a.c, a.k, 64, 48, tn)
                print(json.dumps(r), flush=True)
                f.write(json.dumps(r) + "\n")
            for b in a.time:
                r = timing(b, a.c, a.k, 64, 48, a.iters, tn)
                print(json.dumps(r), flush=True)
                f.write(json.dumps(r) + "\n")
