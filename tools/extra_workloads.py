"""The other BASELINE.json configurations as bench lines (bench.py appends them as `extra_workloads`, timed OUTSIDE the headline
region): config 4 (SPM: fused render+loss+grad, render, root-peak + displacement decode on multi-person maps), config 5 (decode-
only sweep: 96x72 maps, the 11-joint sbp_pis shape, B in {256, 16384}) and config 3 as written (global batch 32 768 split over
the ranks).  Every entry: device time per call (CUDA-graph replay, CUDA events), the algorithmic bytes of SURVEY.md 8(d) and the
fraction of the measured HBM peak.  Inputs are resident and (except B=256, marked) larger than L2.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def graph_time(fn, reps=20, warm=3, rounds=3):
    """ms per call: the call is captured in a CUDA graph -- as many back-to-back copies as fill ~1 ms, so that the host's replay
    rate never shows in the number -- and replayed `reps` times between one event pair; median of `rounds`."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    fn()
    b.record()
    torch.cuda.synchronize()
    inner = max(1, min(16, int(1.0 / max(a.elapsed_time(b), 1e-3))))
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
    ts = []
    for _ in range(rounds):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            g.replay()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / (reps * inner))
    return sorted(ts)[len(ts) // 2]


def _entry(name, ms, units, unit_name, bytes_per_unit, peak, note=None):
    gbs = bytes_per_unit * units / (ms * 1e-3) / 1e9
    e = {"workload": name, "ms": ms, unit_name + "_per_s": units / (ms * 1e-3), "algorithmic_bytes": bytes_per_unit * units,
         "achieved_GBps": gbs, "frac_of_measured_hbm_peak": gbs / peak}
    if note:
        e["note"] = note
    return e


def spm_config4(pb, dev, peak, n=1024, reps=10):
    from _inputs import spm_inputs
    c, j, cnt, target, x = spm_inputs(n, dev)
    img = 35 * 128 * 128 * 4
    out = []
    out.append(_entry(f"config4 SPM fused render+loss+grad, N={n} x 35 x 128x128 (persons in, target never in HBM)",
                      graph_time(lambda: pb.spm_fused(x, c, j, cnt, 1), reps), n, "images", 2 * img, peak))
    out.append(_entry(f"config4 SPM fused render+loss (no grad), N={n}",
                      graph_time(lambda: pb.spm_fused(x, c, j, cnt, 1, want_grad=False), reps), n, "images", img, peak))
    out.append(_entry(f"config4 SPM target render, N={n}", graph_time(lambda: pb.spm_render_batch(c, j, cnt, 128, 1), reps), n, "images", img, peak))
    out.append(_entry(f"config4 SPM dense-target loss+grad (reference signature), N={n}",
                      graph_time(lambda: pb.spm_loss_fused(x, target), reps), n, "images", 3 * img, peak))
    ms = graph_time(lambda: pb.spm_decode_batch(x, 512, 1, 0.5, True, 32), reps)
    _, _, counts, _ = pb.spm_decode_batch(x, 512, 1, 0.5, True, 32)
    roots = int(counts.sum().item())
    per_img = 128 * 128 * 4 + (roots / n) * 34 * 32        # root plane + one 32-byte sector per gathered displacement
    out.append(_entry(f"config4 SPM decode (root NMS + displacement gather), N={n}, thr 0.5, {roots / n:.1f} roots/image",
                      ms, n, "images", per_img, peak, note="one wave of per-image CTAs: latency-bound tail (greedy NMS + scattered gathers)"))
    return out


def decode_config5(pb, dev, peak, reps=20):
    out = []
    gen = torch.Generator(device=dev).manual_seed(0)
    for (h, w, k, b, in_w) in ((96, 72, 17, 4096, 288), (64, 48, 11, 4096, 192), (64, 48, 17, 16384, 192), (64, 48, 17, 256, 192)):
        x = torch.randn(b, k, h, w, device=dev, generator=gen) * 3.0
        maps = b * k
        per = h * w * 4 + 12
        note = "3.3 MB: L2-resident and launch-latency bound, not an HBM number" if b == 256 else None
        for pred, thr in ((True, 0.25), (False, 0.99)):
            src = x if pred else torch.sigmoid(x)
            ms = graph_time(lambda: pb.decode_batch(src, thr, in_w / w, pred), reps)
            out.append(_entry(f"config5 decode {h}x{w} K={k} B={b} pred={pred} thr={thr}", ms, maps, "heatmaps", per, peak, note))
        del x
    return out


def validation_config2(pb, dev, peak, b=4096, k=17, h=64, w=48, reps=20):
    """configs[1] shapes, the VALIDATION forms of the fused kernel (module/sbp_detector.py:33-41: loss, then update_state): nothing is
    stored but the joint rows, so the bytes per heat map are the logits alone (SURVEY 8 d: 12 296 / 12 308)."""
    gen = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(b, k, h, w, device=dev, generator=gen) * 3.0
    kp = torch.stack([torch.rand(b, k, device=dev, generator=gen, dtype=torch.float64) * w,
                      torch.rand(b, k, device=dev, generator=gen, dtype=torch.float64) * h], dim=-1)
    kp[torch.rand(b, k, device=dev, generator=gen) >= 0.85] = -1.0
    jo = torch.empty((b, k, 3), device=dev)
    out = []
    ms = graph_time(lambda: pb.sbp_fused(x, keypoints=kp, sigma=2, want_grad=False), reps)
    out.append(_entry(f"config2 validation loss (render + loss, no grad), B={b}", ms, b * k, "heatmaps", h * w * 4 + 8, peak))
    ms = graph_time(lambda: pb.sbp_fused(x, keypoints=kp, sigma=2, want_grad=False, decode=True, conf_threshold=0.25, coord_scale=4.0,
                                         out={"joints": jo}), reps)
    out.append(_entry(f"config2 validation step (render + loss + decode, no grad), B={b}", ms, b * k, "heatmaps", h * w * 4 + 20, peak))
    return out


def head_fusion(pb, dev, peak, b=1024, c=512, k=17, h=64, w=48, reps=5):
    """the detector's 1x1 head (512 -> 17, models/detector/sbp.py:35-37) fused with loss + dlogits + decode on tcgen05
    (pose_sbp_head_fused), next to the unfused pair on the same box: torch's fp32 conv2d (cuDNN / cuBLAS, TF32 off -- the
    reference's arithmetic) writing the logits, then the fused loss kernel reading them."""
    gen = torch.Generator(device=dev).manual_seed(0)
    feats = torch.randn((b, c, h, w), generator=gen, device=dev).relu_()
    weight = torch.randn((k, c), generator=gen, device=dev) * (2.0 / c) ** 0.5
    kp = torch.stack([torch.rand(b, k, device=dev, generator=gen, dtype=torch.float64) * w,
                      torch.rand(b, k, device=dev, generator=gen, dtype=torch.float64) * h], dim=-1)
    kp[torch.rand(b, k, device=dev, generator=gen) >= 0.85] = -1.0
    dl = torch.empty((b, k, h, w), device=dev)
    jo = torch.empty((b, k, 3), device=dev)
    fbytes, lbytes = feats.numel() * 4, dl.numel() * 4
    out = []
    ms = graph_time(lambda: pb.sbp_head_fused(feats, weight, kp, sigma=2, want_grad=True, decode=True, conf_threshold=0.25, coord_scale=4.0,
                                              out={"dlogits": dl, "joints": jo}), reps)
    e = _entry(f"f-3 head fusion, training form: 1x1 conv {c}->{k} (tcgen05, 3xTF32) + loss + dlogits + decode, B={b} x {c} x {h}x{w} features; "
               "logits never in HBM", ms, b, "images", (fbytes + lbytes) / b, peak)
    out.append(e)
    ms_v = graph_time(lambda: pb.sbp_head_fused(feats, weight, kp, sigma=2, want_grad=False, decode=True, conf_threshold=0.25, coord_scale=4.0,
                                                out={"joints": jo}), reps)
    out.append(_entry(f"f-3 head fusion, validation form (loss + decode, nothing written but {b * k} joint rows), B={b}", ms_v, b, "images",
                      fbytes / b, peak))
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    w4 = weight.view(k, c, 1, 1)
    ms_conv = graph_time(lambda: torch.nn.functional.conv2d(feats, w4), reps)
    ms_pair = graph_time(lambda: pb.sbp_fused(torch.nn.functional.conv2d(feats, w4), keypoints=kp, sigma=2, want_grad=True, decode=True,
                                              conf_threshold=0.25, coord_scale=4.0, out={"dlogits": dl, "joints": jo}), reps)
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    e["unfused_baseline"] = {"what": "torch conv2d fp32 (library kernel, TF32 off) -> logits in HBM -> pose_sbp_fused (render+loss+grad+decode), same inputs",
                             "conv2d_ms": ms_conv, "conv2d_then_fused_loss_ms": ms_pair, "speedup": ms_pair / ms,
                             "algorithmic_bytes": fbytes + 3 * lbytes}
    return out


def config3_as_written(pb, pd, dev, peak, world, rank, global_batch=32768, reps=10):
    """BASELINE.json configs[2]: global batch 32 768 split over the ranks (N=2: 16 384 images per GPU), fused step with the
    exchange, graph replay, max over ranks."""
    import torch.distributed as dist
    K, H, W = 17, 64, 48
    b = global_batch // world
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    logits = torch.randn(b, K, H, W, device=dev, generator=gen) * 3.0
    kp = torch.stack([torch.rand(b, K, device=dev, generator=gen, dtype=torch.float64) * W,
                      torch.rand(b, K, device=dev, generator=gen, dtype=torch.float64) * H], dim=-1)
    kp[torch.rand(b, K, device=dev, generator=gen) >= 0.85] = -1.0
    bbox = torch.rand(b, 4, device=dev, generator=gen, dtype=torch.float64) * 300 + 40
    iid = torch.arange(b, device=dev, dtype=torch.int64) + rank * b
    ex, kind = pd.make_exchange(b, K, dev, iid, torch.ones(b, device=dev, dtype=torch.int64), defer=1)
    outs = dict(dlogits=torch.empty_like(logits), joints=torch.empty((b, K, 3), dtype=torch.float32, device=dev),
                loss=torch.empty((), dtype=torch.float32, device=dev))
    if kind != "p2p":
        outs.update(ex.out_views())

    def step():
        if kind == "p2p":
            pb.sbp_fused(logits, keypoints=kp, sigma=2, want_grad=True, decode=True, conf_threshold=0.25, coord_scale=4.0,
                         global_batch=global_batch, bbox=bbox, input_size=(256, 192), out=outs, exchange=ex)
            return ex.finish(global_batch)
        r = pb.sbp_fused(logits, keypoints=kp, sigma=2, want_grad=True, decode=True, conf_threshold=0.25, coord_scale=4.0,
                         global_batch=global_batch, bbox=bbox, input_size=(256, 192), out=outs)
        ex.exchange()
        return ex.global_loss(global_batch, local_loss=r["loss"])

    for _ in range(4):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(4):
            step()
    torch.cuda.current_stream().wait_stream(side)
    host_steps = getattr(ex, "steps", 0)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(4):                      # a whole turn of the exchange ring per replay
            step()
    if kind == "p2p":
        ex.steps = host_steps
    g.replay()
    if kind == "p2p":
        ex.advance(4)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        g.replay()
    if kind == "p2p":
        ex.advance(4 * reps)
        ex.flush(global_batch)
    e.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(e) / (4 * reps)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ent = _entry(f"config3 as written: global batch {global_batch} over {world} GPU(s) = {b} images per GPU, fused render+loss+grad+decode + "
                 f"back-projection + exchange ({kind})", ms, global_batch * K, "heatmaps", 24596, peak * world)
    ent["scaling"] = "strong"
    ent["batch_per_gpu"] = b
    return ent
