#!/usr/bin/env python
"""Per-kernel device timings (CUDA events) at the BASELINE.json shapes, as a fraction of the HBM roofline.

    python tools/kbench.py [--batch 4096] [--reps 30] [--out gpurun_out/kbench.json] [--only NAME]

Every stage is timed alone, inputs resident in HBM and larger than L2, 5 warm-up + `reps` timed launches, median and best.
Algorithmic bytes per heat map are the SURVEY.md 8(d) figures.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import pose_b200 as pb  # noqa: E402


EAGER = False      # --eager: time back-to-back eager launches instead of CUDA-graph replays


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


def timeit(fn, reps, warm=5):
    """Device time per call: the call is captured in a CUDA graph -- as many back-to-back copies as fill ~1 ms, so the host's
    replay rate never shows in a short kernel's number -- and replayed `reps` times between one event pair (the eager
    queue-saturated loop is the fallback if capture fails); best and median of 3 such measurements."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    fn()
    b.record()
    torch.cuda.synchronize()
    inner = max(1, min(16, int(1.0 / max(a.elapsed_time(b), 1e-3))))
    run = None
    try:
        if EAGER:
            raise RuntimeError("eager timing requested")
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(inner):
                fn()
        g.replay()
        torch.cuda.synchronize()
        run = g.replay
    except Exception:       # noqa: BLE001
        torch.cuda.synchronize()
        run = fn
        inner = 1
    ts = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            run()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b) / (reps * inner))
    ts.sort()
    return ts[1], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--reps", type=int, default=30)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "kbench.json"))
    ap.add_argument("--only", default="")
    ap.add_argument("--hw", default="64x48")
    ap.add_argument("--k", type=int, default=17)
    ap.add_argument("--no-tma", dest="tma", action="store_false", help="register-staged kernels instead of the bulk-async (TMA) staged ones")
    ap.add_argument("--no-spm", action="store_true")
    ap.add_argument("--spm-n", type=int, default=1024, help="second SPM batch size (config 4: 1024 images)")
    ap.add_argument("--eager", action="store_true", help="no CUDA graph: queue-saturated eager launches")
    args = ap.parse_args()
    global EAGER
    EAGER = args.eager
    dev = torch.device("cuda", 0)
    B, K = args.batch, args.k
    H, W = map(int, args.hw.split("x"))
    sigma = 2 if H == 64 else -1
    gen = torch.Generator(device=dev).manual_seed(0)
    logits = torch.randn(B, K, H, W, device=dev, generator=gen) * 3
    kp = torch.stack([torch.rand(B, K, device=dev, generator=gen, dtype=torch.float64) * W,
                      torch.rand(B, K, device=dev, generator=gen, dtype=torch.float64) * H], -1)
    kp[torch.rand(B, K, device=dev, generator=gen) >= 0.85] = -1
    bbox = torch.rand(B, 4, device=dev, generator=gen, dtype=torch.float64) * 300 + 40
    gen_t = pb.SBPHeatmapGenerator([H, W], K, sigma)
    target = gen_t.render_batch(kp)
    realistic = torch.logit((target + 0.05 * torch.rand(target.shape, device=dev, generator=gen)).clamp(1e-4, 1 - 1e-4))
    joints = pb.decode_batch(logits, 0.25, 4.0, True)
    maps = B * K
    map_bytes = H * W * 4
    pk = peak()
    res = {}

    def run(name, fn, bytes_per_map):
        if args.only and args.only not in name:
            return
        med, best = timeit(fn, args.reps)
        gbs = bytes_per_map * maps / (med * 1e-3) / 1e9
        res[name] = {"ms_median": med, "ms_best": best, "heatmaps_per_s": maps / (med * 1e-3), "GBps": gbs, "frac_of_measured_peak": gbs / pk,
                     "bytes_per_map": bytes_per_map}
        print(f"{name:44s} {med*1e3:9.1f} us (best {best*1e3:8.1f})  {gbs:8.1f} GB/s  {gbs/pk*100:5.1f}%  {maps/(med*1e-3)/1e6:8.1f} M maps/s", flush=True)

    out = torch.empty_like(logits)
    run("render_only", lambda: gen_t.render_batch(kp, out=out), map_bytes + 8)
    tma = args.tma
    dl = torch.empty_like(logits)
    fo = dict(dlogits=dl, joints=torch.empty(B, K, 3, device=dev), loss=torch.empty((), device=dev), loss_num=torch.empty(2, dtype=torch.float64, device=dev))
    run("fused_render_loss_grad", lambda: pb.sbp_fused(logits, keypoints=kp, sigma=sigma, tma=tma, out=fo), 2 * map_bytes + 8)
    run("fused_render_loss_grad_decode", lambda: pb.sbp_fused(logits, keypoints=kp, sigma=sigma, decode=True, coord_scale=4.0, tma=tma, out=fo), 2 * map_bytes + 20)
    run("fused_render_loss_only(no grad)", lambda: pb.sbp_fused(logits, keypoints=kp, sigma=sigma, want_grad=False, tma=tma, out=fo), map_bytes + 8)
    run("fused_render_loss_decode(no grad)", lambda: pb.sbp_fused(logits, keypoints=kp, sigma=sigma, want_grad=False, decode=True, tma=tma, out=fo), map_bytes + 20)
    run("fused_full_step(+backproject epilogue)", lambda: pb.sbp_fused(logits, keypoints=kp, sigma=sigma, decode=True, coord_scale=4.0, tma=tma, out=fo,
                                                                         bbox=bbox, input_size=(4 * H, 4 * W)), 2 * map_bytes + 20 + 24)
    run("dense_loss_grad", lambda: pb.sbp_fused(logits, target=target), 3 * map_bytes)
    run("dense_loss_only(no grad)", lambda: pb.sbp_fused(logits, target=target, want_grad=False), 2 * map_bytes)
    for pred in (True, False):
        for src, tag in ((logits, "randn"), (realistic, "realistic")):
            run(f"decode_pred{int(pred)}_{tag}", lambda: pb.decode_batch(src, 0.25, 4.0, pred), map_bytes + 12)
    run("decode_pred1_randn_sigmoid_cuda", lambda: pb.decode_batch(logits, 0.25, 4.0, True, sigmoid_ref="cuda"), map_bytes + 12)
    xf = realistic.flip(-1).contiguous()
    run("decode_flip_test_pred1_realistic", lambda: pb.decode_batch(realistic, 0.25, 4.0, True, flipped=xf), 2 * map_bytes + 12)
    run("decode_pred0_target_thr.99", lambda: pb.decode_batch(target, 0.99, 4.0, False), map_bytes + 12)
    run("backproject_rows", lambda: pb.backproject_rows(joints, bbox, (256, 192)), 24)

    if (not args.only or "spm" in args.only) and not args.no_spm:
        from _inputs import spm_inputs
        n = 256
        c, j, cnt, t, x = spm_inputs(n, dev)
        img_bytes = 35 * 128 * 128 * 4

        def run_spm(name, fn, bytes_per_img):
            med, best = timeit(fn, args.reps)
            gbs = bytes_per_img * n / (med * 1e-3) / 1e9
            res[name] = {"ms_median": med, "ms_best": best, "images_per_s": n / (med * 1e-3), "GBps": gbs, "frac_of_measured_peak": gbs / pk}
            print(f"{name:44s} {med*1e3:9.1f} us (best {best*1e3:8.1f})  {gbs:8.1f} GB/s  {gbs/pk*100:5.1f}%  {n/(med*1e-3)/1e3:8.1f} K img/s", flush=True)

        run_spm("spm_render(N=256)", lambda: pb.spm_render_batch(c, j, cnt, 128, 1), img_bytes)
        run_spm("spm_loss_grad(N=256)", lambda: pb.spm_loss_fused(x, t), 3 * img_bytes)
        run_spm("spm_loss_only(N=256)", lambda: pb.spm_loss_fused(x, t, want_grad=False), 2 * img_bytes)
        run_spm("spm_fused_render_loss_grad(N=256)", lambda: pb.spm_fused(x, c, j, cnt, 1), 2 * img_bytes)
        run_spm("spm_fused_render_loss_only(N=256)", lambda: pb.spm_fused(x, c, j, cnt, 1, want_grad=False), img_bytes)
        run_spm("spm_fused_render_loss_grad_target(N=256)", lambda: pb.spm_fused(x, c, j, cnt, 1, want_target=True), 3 * img_bytes)
        run_spm("spm_decode(N=256,thr=.5)", lambda: pb.spm_decode_batch(x, 512, 1, 0.5, True, 32), 128 * 128 * 4)
        n = args.spm_n
        if n != 256:
            c, j, cnt, t, x = spm_inputs(n, dev)
            run_spm(f"spm_fused_render_loss_grad(N={n})", lambda: pb.spm_fused(x, c, j, cnt, 1), 2 * img_bytes)
            run_spm(f"spm_loss_grad(N={n})", lambda: pb.spm_loss_fused(x, t), 3 * img_bytes)
            run_spm(f"spm_render(N={n})", lambda: pb.spm_render_batch(c, j, cnt, 128, 1), img_bytes)
            run_spm(f"spm_decode(N={n},thr=.5)", lambda: pb.spm_decode_batch(x, 512, 1, 0.5, True, 32), 128 * 128 * 4)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump({"batch": B, "K": K, "H": H, "W": W, "tma": args.tma, "peak_GBps": pk, "results": res}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
