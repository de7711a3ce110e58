#!/usr/bin/env python
"""Headline benchmark: heatmaps/sec of the SBP hot path (render + loss + grad + decode) on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU algorithm (oracle port), all host cores

One "step" = one pass of the hot path over one synthetic batch (BASELINE.json configs[1]: B=4096 per GPU,
K=17, 64x48 fp32 heat maps, sigma 2, 256x192 input): fused render+loss+grad+decode kernel + the epilogue launch
(fixed-order loss reduction, back-projection to COCO rows and, for N>1, the exchange of rows / loss numerators over NVLink peer
memory).  Prints ONE JSON line on rank 0:

    value / ms_per_step   K timed steps (CUDA-graph replays), inputs resident, CUDA events, max over ranks
    eager_ms_per_step     the same steps issued eagerly through the Python API (no graph): host launch path included
    roofline              the fused kernel alone against the measured HBM copy peak
    e2e                   the same step from pinned HOST buffers: chunked H2D on a copy stream, the kernels of chunk c run
                          under the copy of chunk c+1, rows + loss copied back; `pcie` = what a plain H2D of the same bytes
                          reaches on this box at this N (the roofline of e2e)
    aten_baseline         the reference's own op chain (ATen kernels) on the SAME GPU: SBPLoss fwd+bwd on the full batch,
                          the per-sample / per-joint nms_sbp loop on a bounded sample
    cpu_baseline          the reference's CPU algorithm (oracle port) on this box's host cores (N=1 only)
    exchange_check        N>1: parity of the three exchange implementations + a 20 000-step stress of the in-band mode
    extra_workloads       configs 3 (as written), 4 (SPM), 5 (decode sweep) and the f-3 head fusion, timed outside the headline region
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

K, H, W, SIGMA, IN_H, IN_W, THR = 17, 64, 48, 2, 256, 192, 0.25
BYTES_FUSED = 24596          # algorithmic bytes per heat map, full pipeline with logits read once (SURVEY.md 8 d, row C)
METRIC, UNIT = "heatmaps_per_sec", "heatmaps/s"


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any (profiles/traffic.json)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("sbp_fused_bytes_per_launch")
        except Exception:
            return None
    return None


_emit = print


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _once(self):
        nv = self.nv
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                     "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10, "applications_clocks_setting": 0x2}
            for k, bit in names.items():
                if r & bit:
                    self.reasons.add(k)
        except Exception:
            pass

    def _run(self):
        while not self._stop.is_set():
            self._once()
            self._stop.wait(0.02)

    def __enter__(self):
        if self.nv:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        if self.nv:
            self._once()
            self._stop.set()
            self._thread.join()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def bind_to_gpu_numa_node(index):
    """Pin this process to the CPUs NVML reports as closest to GPU `index` BEFORE any pinned host buffer is allocated, so that
    first-touch places those buffers on the GPU's NUMA node (the H2D copies of 8 ranks then do not all cross one socket link).
    Returns a short description for the JSON line; failures are not fatal."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        before = len(os.sched_getaffinity(0))
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * i + b for i, wd in enumerate(mask) for b in range(64) if (int(wd) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus and len(cpus) < before:
            os.sched_setaffinity(0, cpus)
            return f"bound to {len(cpus)} of {before} CPUs (NVML ideal CPU set of GPU {index})"
        return f"not narrowed: NVML ideal set covers all {before} usable CPUs"
    except Exception as e:      # noqa: BLE001
        return f"unchanged ({type(e).__name__})"


# --------------------------------------------------------------------------------------------- CPU baseline (oracle port)
def make_cpu_pipeline(n, procs):
    """The reference's CPU algorithm on `n` images per pass: per-sample render loop -> SBPLoss fwd+bwd -> per-sample
    decode loop -> back-projection + COCO rows (utils/sbp_utils.py:33-53, :103-118, :131-164; models/loss/sbp_loss.py),
    the per-sample loops spread over `procs` worker processes (oracle/cpu_pipeline.py)."""
    from oracle import sbp_oracle as so
    from oracle.cpu_pipeline import ReferencePipeline
    return ReferencePipeline(so.make_config1_inputs(n, K, H, W), H, W, SIGMA, IN_H, IN_W, THR, procs=procs)


def ref_shape(args):
    """(usable cores, worker processes, images per pass) of the CPU arm: every usable core, at least 8 images per worker."""
    from oracle.cpu_pipeline import usable_cores
    cores = usable_cores()
    procs = args.ref_procs if args.ref_procs > 0 else min(cores, 64)
    n = args.ref_sample if args.ref_sample > 0 else max(64, 8 * procs)
    return cores, procs, n


CPU_WARM, CPU_MAX_PASSES, CPU_MAX_SECONDS = 2, 20, 15.0


def time_cpu_pipeline(pipe, max_passes=CPU_MAX_PASSES, warm=CPU_WARM, exact=False):
    """ONE statistic for both places the CPU port is timed (the `cpu_baseline` key of the CUDA line and the `--impl reference`
    arm): `warm` warm-up passes, then the MEDIAN pass time of up to 20 passes or 15 s (`cpu_baseline`), or of EXACTLY
    `max_passes` passes (`exact`: the reference arm runs the --steps / --warmup it was given).  -> (seconds per pass, passes)."""
    for _ in range(warm):
        pipe.run_pass()
    ts = []
    t_start = time.perf_counter()
    while (len(ts) < max_passes) if exact else (len(ts) < 3 or (len(ts) < max_passes and time.perf_counter() - t_start < CPU_MAX_SECONDS)):
        t0 = time.perf_counter()
        pipe.run_pass()
        ts.append(time.perf_counter() - t0)
    ts.sort()
    return ts[len(ts) // 2], len(ts)


def cpu_baseline_record(value, n, passes, procs, threads, cores, warm=CPU_WARM):
    return {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "statistic": "median pass",
            "sample": (f"{n} images of the same workload per pass, {warm} warm-up + {passes} timed passes (median); oracle PORT of the "
                       f"reference's per-sample Python loops (the reference is pure Python and is not on this box; the port is pinned to it by "
                       f"tests/golden): render and decode over {procs} worker process(es), SBPLoss fwd+bwd on {threads} torch threads")}


def run_reference(args):
    """--impl reference: rank 0 times the oracle port (the reference is pure Python and cannot travel to the GPU box;
    its CPU algorithm is restated in oracle/, pinned to it by tests/golden) on all the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cores, procs, n = ref_shape(args)
    # exactly --warmup warm-up and --steps timed passes, each a bounded sample of the workload; the sample shrinks (never below
    # one image per worker) if the requested passes would not finish within ~3 minutes at the first pass's speed
    steps, warm = max(1, args.steps), max(0, args.warmup)
    pipe = make_cpu_pipeline(n, procs)          # forks its workers before the parent's torch thread pool exists
    torch.set_num_threads(cores)
    t0 = time.perf_counter()
    pipe.run_pass()
    first = time.perf_counter() - t0
    if first * (steps + warm) > 180.0 and not args.ref_sample:
        pipe.close()
        n = max(procs, int(n * 180.0 / (first * (steps + warm))))
        pipe = make_cpu_pipeline(n, procs)
        pipe.run_pass()
    dt, passes = time_cpu_pipeline(pipe, steps, warm=warm, exact=True)
    pipe.close()
    value = n * K / dt
    rec = cpu_baseline_record(value, n, passes, procs, torch.get_num_threads(), cores, warm=warm)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": passes,
        "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"SBP 256x192 (configs[1]): {K}x{H}x{W} heat maps, sigma {SIGMA}: render + JointsMSE fwd/bwd + decode + "
                               f"back-projection; bounded sample of {n} images per step (of the B=4096 workload)",
                   "batch_per_step": n, "threads": torch.get_num_threads(), "worker_processes": procs},
        "cpu_baseline": rec,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(json.dumps(line))


# --------------------------------------------------------------------------------------------- the reference's ATen op chain on this GPU
def aten_baseline(torch, dev, logits, kp, bbox, sample=48):
    """What the reference executes when its tensors live on the GPU (models/loss/sbp_loss.py:29-49 + autograd; utils/sbp_utils.py:
    56-82, :103-118 per sample and joint; :137-146) -- the op chains are restated in oracle/sbp_oracle.py and run here on CUDA
    tensors as a BASELINE LEG (never on the product path).  The dense target is what the reference's dataloader would have rendered
    on the CPU (not timed: it is not a GPU stage in the reference)."""
    from oracle import sbp_oracle as so
    import pose_b200 as pb
    B = logits.size(0)
    target = pb.SBPHeatmapGenerator([H, W], K, SIGMA).render_batch(kp)

    def loss_step():
        x = logits.detach().requires_grad_(True)
        loss = so.sbp_loss(x, target)
        loss.backward()
        return loss

    for _ in range(2):
        loss_step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    a.record()
    for _ in range(reps):
        loss_step()
    b.record()
    torch.cuda.synchronize()
    loss_ms = a.elapsed_time(b) / reps
    del target
    # decode + back-projection: the reference calls DecodeSBP once per sample (batch of one) from update_state
    n = min(sample, B)
    bb = bbox[:n].cpu()

    def decode_sample(i):
        j = so.sbp_decode_loop_torch(logits[i:i + 1], IN_W, THR, True)
        j[..., :1] *= bb[i, 2] / IN_W           # utils/sbp_utils.py:141-146 (bbox is a CPU float64 tensor from collate)
        j[..., 1:2] *= bb[i, 3] / IN_H
        j[..., :1] += bb[i, 0]
        j[..., 1:2] += bb[i, 1]
        return j

    decode_sample(0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        decode_sample(i)
    torch.cuda.synchronize()
    dec_ms_per_sample = (time.perf_counter() - t0) * 1e3 / n
    total_ms = loss_ms + dec_ms_per_sample * B
    return {"what": "the reference's ATen op chain on this GPU (same box, same inputs): SBPLoss forward + backward over the full batch "
                    "(CUDA events) and the per-sample / per-joint nms_sbp + back-projection loop (wall clock: it synchronises the host ~6 "
                    f"times per joint) extrapolated from {n} samples; target render excluded (a CPU dataloader stage in the reference)",
            "loss_fwd_bwd_ms": loss_ms, "decode_ms_per_sample": dec_ms_per_sample, "decode_sample": n,
            "ms_per_step_extrapolated": total_ms, "value": B * K / (total_ms * 1e-3), "unit": UNIT,
            "loss_only_value": B * K / (loss_ms * 1e-3)}


# --------------------------------------------------------------------------------------------- CUDA arm
def run_cuda(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cpu_pipe = None
    if world == 1 and not args.no_cpu_baseline:
        # the CPU-baseline workers are forked now, before this process owns a CUDA context; they sleep until the GPU part is done
        cpu_cores, cpu_procs, cpu_n = ref_shape(args)
        cpu_pipe = make_cpu_pipeline(cpu_n, cpu_procs)
    affinity = bind_to_gpu_numa_node(local) if world > 1 and not args.no_affinity else "not bound (single rank)"
    import pose_b200 as pb
    from pose_b200 import dist as pd
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: pose_b200 has no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pb.lib()

    B = args.batch
    global_batch = B * world
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    logits = torch.randn(B, K, H, W, device=dev, generator=gen) * 3.0
    kp = torch.stack([torch.rand(B, K, device=dev, generator=gen, dtype=torch.float64) * W,
                      torch.rand(B, K, device=dev, generator=gen, dtype=torch.float64) * H], dim=-1)
    kp[torch.rand(B, K, device=dev, generator=gen) >= 0.85] = -1.0
    bbox = torch.stack([torch.rand(B, device=dev, generator=gen, dtype=torch.float64) * 400,
                        torch.rand(B, device=dev, generator=gen, dtype=torch.float64) * 400,
                        torch.rand(B, device=dev, generator=gen, dtype=torch.float64) * 260 + 40,
                        torch.rand(B, device=dev, generator=gen, dtype=torch.float64) * 340 + 60], dim=-1)
    image_id = torch.arange(B, device=dev, dtype=torch.int64) + rank * B
    category_id = torch.ones(B, device=dev, dtype=torch.int64)
    ex, ex_kind = pd.make_exchange(B, K, dev, image_id, category_id, prefer_p2p=not args.nccl, defer=args.defer)
    dlogits = torch.empty_like(logits)
    joints = torch.empty((B, K, 3), dtype=torch.float32, device=dev)
    loss_local = torch.empty((), dtype=torch.float32, device=dev)
    outs = dict(dlogits=dlogits, joints=joints, loss=loss_local)
    if ex_kind != "p2p":
        outs.update(ex.out_views())

    def step(src_logits=logits, src_kp=kp, src_bbox=bbox):
        """One pass of the hot path over the batch.  N=1: 2 launches (fused kernel + epilogue).  N>1: the same 2 launches -- the
        epilogue stores rows / numerators into every rank over NVLink and completes the previous step (in-band) -- or, on
        the NCCL fallback, + 1 all-gather + 1 reduce launch."""
        if ex_kind == "p2p":
            pb.sbp_fused(src_logits, keypoints=src_kp, sigma=SIGMA, want_grad=True, decode=True, conf_threshold=THR,
                         coord_scale=IN_W / W, global_batch=global_batch, bbox=src_bbox, input_size=(IN_H, IN_W), out=outs,
                         exchange=ex)
            return ex.finish(global_batch)
        r = pb.sbp_fused(src_logits, keypoints=src_kp, sigma=SIGMA, want_grad=True, decode=True, conf_threshold=THR,
                         coord_scale=IN_W / W, global_batch=global_batch, bbox=src_bbox, input_size=(IN_H, IN_W), out=outs)
        ex.exchange()
        return ex.global_loss(global_batch, local_loss=r["loss"])

    def fence():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step()
    fence()

    def timed(run, n, after=None):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fence()
        e0.record()
        for _ in range(n):
            run()
        if after is not None:
            after(n)
        e1.record()
        fence()
        ms = e0.elapsed_time(e1) / n
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def flush_after(n):
        if ex_kind == "p2p":
            ex.flush(global_batch)         # defer=1: the last step's exchange completes inside the timed region too

    # ---- eager: the Python API call by call (argument checks, ctypes, launches; no driver queries -- they are cached)
    c0 = pb.launch_count()
    eager_ms = timed(step, args.steps, flush_after)
    launches_per_step = (pb.launch_count() - c0 - (1 if ex_kind == "p2p" and getattr(ex, "defer", 0) else 0)) / args.steps

    # ---- headline: the same step replayed from a CUDA graph
    graph = None
    if not args.no_graph:
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        host_steps = getattr(ex, "steps", 0)
        with torch.cuda.graph(graph):
            step()
        if ex_kind == "p2p":
            ex.steps = host_steps          # capturing ran no device work: keep the host mirror of the step counter in sync
        for _ in range(3):
            graph.replay()
        if ex_kind == "p2p":
            ex.advance(3)
        fence()

    def replay_after(n):
        if graph is not None and ex_kind == "p2p":
            ex.advance(n)
        flush_after(n)

    with ClockSampler(local) as clk:
        ms = timed(graph.replay if graph is not None else step, args.steps, replay_after)
    value = global_batch * K / (ms * 1e-3)
    launches = int(round(launches_per_step * args.steps))

    # ---- the dominant kernel alone (fused render+loss+grad+decode + its epilogue launch), queue-saturated, same inputs
    def fused_only():
        pb.sbp_fused(logits, keypoints=kp, sigma=SIGMA, want_grad=True, decode=True, conf_threshold=THR, coord_scale=IN_W / W,
                     global_batch=global_batch, out=outs)
    for _ in range(5):
        fused_only()
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kreps = max(20, min(args.steps, 200))
    k0.record()
    for _ in range(kreps):
        fused_only()
    k1.record()
    torch.cuda.synchronize()
    kern_ms = k0.elapsed_time(k1) / kreps

    # ---- e2e: the same step through the public API from pinned HOST buffers, result back on the host.  The H2D copy is issued in
    # chunks on a copy stream; the kernels of chunk c (fused + epilogue on that slice of the batch) run while chunk c+1 is still
    # crossing PCIe, so only the last chunk's kernels and the D2H of the rows are exposed.
    C = max(1, args.e2e_chunks)
    bounds = [(B * c // C, B * (c + 1) // C) for c in range(C)]
    h_logits = torch.empty(logits.shape, dtype=logits.dtype, pin_memory=True).copy_(logits)
    h_kp = torch.empty(kp.shape, dtype=kp.dtype, pin_memory=True).copy_(kp)
    h_bbox = torch.empty(bbox.shape, dtype=bbox.dtype, pin_memory=True).copy_(bbox)
    h_packed = torch.empty((global_batch, 3 * K + 1), dtype=torch.float32, pin_memory=True)
    h_loss = torch.empty((), dtype=torch.float32, pin_memory=True)
    d_logits, d_kp, d_bbox = torch.empty_like(logits), torch.empty_like(kp), torch.empty_like(bbox)
    sx = pd.ShardExchange(B, K, dev)                      # the e2e path exchanges with ONE NCCL all-gather per step (rows + numerators)
    sx.ids.copy_(torch.stack([image_id, category_id], dim=1))
    chunk_nums = torch.zeros((C, 2), dtype=torch.float64, device=dev)
    chunk_loss = torch.empty((), dtype=torch.float32, device=dev)
    copy_stream = torch.cuda.Stream(device=dev)
    evs = [torch.cuda.Event() for _ in range(C)]
    done = torch.cuda.Event()

    def e2e_step():
        cur = torch.cuda.current_stream()
        copy_stream.wait_event(done)                      # the previous step's kernels are done with the device input buffers
        with torch.cuda.stream(copy_stream):
            d_kp.copy_(h_kp, non_blocking=True)
            d_bbox.copy_(h_bbox, non_blocking=True)
            for c, (lo, hi) in enumerate(bounds):
                d_logits[lo:hi].copy_(h_logits[lo:hi], non_blocking=True)
                evs[c].record(copy_stream)
        for c, (lo, hi) in enumerate(bounds):
            cur.wait_event(evs[c])
            pb.sbp_fused(d_logits[lo:hi], keypoints=d_kp[lo:hi], sigma=SIGMA, want_grad=True, decode=True, conf_threshold=THR,
                         coord_scale=IN_W / W, global_batch=global_batch, bbox=d_bbox[lo:hi], input_size=(IN_H, IN_W),
                         out=dict(dlogits=dlogits[lo:hi], joints=joints[lo:hi], loss=chunk_loss, packed=sx.packed[lo:hi], loss_num=chunk_nums[c]))
        done.record(cur)
        # numerators of the chunks -> this rank's numerators (fixed order), then the exchange and the global loss
        pb._cabi.check(pb.lib().pose_loss_reduce(pb._cabi.ptr(chunk_nums), C, 2, 5.0, 1.0, 1.0 / (2.0 * K * global_batch), None,
                                                 pb._cabi.ptr(sx.loss_num), pb._cabi.stream_ptr(dev)), "pose_loss_reduce")
        sx.exchange()
        loss = sx.global_loss(global_batch) if world > 1 else _single_loss()
        h_packed.copy_(sx.gathered_packed(), non_blocking=True)
        h_loss.copy_(loss, non_blocking=True)

    single_loss = torch.empty((), dtype=torch.float32, device=dev)

    def _single_loss():
        pb._cabi.check(pb.lib().pose_loss_reduce(pb._cabi.ptr(sx.loss_num), 1, 2, 5.0, 1.0, 1.0 / (2.0 * K * global_batch),
                                                 pb._cabi.ptr(single_loss), None, pb._cabi.stream_ptr(dev)), "pose_loss_reduce")
        return single_loss

    e2e_steps = max(3, min(args.steps, 20))
    done.record(torch.cuda.current_stream())
    for _ in range(3):
        e2e_step()
    e2e_ms = timed(e2e_step, e2e_steps)
    h2d = h_logits.numel() * 4 + h_kp.numel() * 8 + h_bbox.numel() * 8
    d2h = h_packed.numel() * 4 + 4
    loss_host = float(h_loss)
    # the loss the e2e path reports must be the loss of the device-resident step on the same data
    loss_ref = float(step())
    if ex_kind == "p2p":
        loss_ref = float(ex.flush(global_batch))
    e2e_loss_ok = abs(loss_host - loss_ref) <= 1e-5 * abs(loss_ref)

    # ---- PCIe roofline of e2e: a plain H2D copy of the same bytes, all ranks at once
    def h2d_only():
        d_logits.copy_(h_logits, non_blocking=True)
        d_kp.copy_(h_kp, non_blocking=True)
        d_bbox.copy_(h_bbox, non_blocking=True)
    for _ in range(2):
        h2d_only()
    pcie_ms = timed(h2d_only, max(3, min(args.steps, 10)))

    extras, xcheck, aten = [], None, None
    peak, peak_src = peak_hbm()
    if world > 1 and not args.no_exchange_check:
        import exchange_check
        xcheck = exchange_check.run(pb, pd, dev, B, world, rank, args.stress_steps)
        allok = torch.tensor([int(xcheck["ok"])], device=dev)
        dist.all_reduce(allok, op=dist.ReduceOp.MIN)
        xcheck["all_ranks_ok"] = bool(allok.item())
    if not args.no_extras:
        import extra_workloads as xw
        try:
            extras.append(xw.config3_as_written(pb, pd, dev, peak, world, rank))
        except Exception as e:      # noqa: BLE001
            extras.append({"workload": "config3 as written", "error": f"{type(e).__name__}: {e}"})
        if rank == 0 and world == 1:
            for fn in (xw.validation_config2, xw.spm_config4, xw.decode_config5, xw.head_fusion):
                try:
                    extras += fn(pb, dev, peak)
                except Exception as e:      # noqa: BLE001
                    extras.append({"workload": fn.__name__, "error": f"{type(e).__name__}: {e}"})
    if rank == 0 and world == 1 and not args.no_aten_baseline:
        try:
            aten = aten_baseline(torch, dev, logits, kp, bbox)
        except Exception as e:      # noqa: BLE001
            aten = {"error": f"{type(e).__name__}: {e}"}

    if rank == 0:
        achieved = BYTES_FUSED * B * K / (kern_ms * 1e-3) / 1e9
        e2e_value = global_batch * K / (e2e_ms * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"SBP 256x192 (configs[1]): B={B} per GPU x {K} joints x {H}x{W} fp32 heat maps, sigma {SIGMA}; "
                                   "fused render+loss+grad+decode + back-projection"
                                   + ({"p2p": "; rows + loss numerators exchanged by the epilogue kernel over NVLink peer memory (no NCCL call)",
                                       "nccl": "; one NCCL all-gather (rows + loss numerators + ids) per step"}.get(ex_kind, "")),
                       "batch_per_gpu": B, "global_batch": global_batch, "partition": f"images x{world}", "exchange": ex_kind, "exchange_defer": getattr(ex, "defer", 0),
                       "l2": "inputs (856 MB logits per GPU) larger than the 126 MB L2; no explicit flush",
                       "kp_dtype": "f64", "sigmoid_ref": pb._cabi.DEFAULT_SIGMOID_REF, "loss": loss_ref, "cuda_graph": graph is not None,
                       "cpu_affinity": affinity},
            "eager_ms_per_step": eager_ms,
            "roofline": {"bound": "hbm", "kernel": "sbp_fused_tma_kernel<GRAD,DECODE> (bulk-async staged, one 64-thread CTA per heat map; + its epilogue launch: two-level loss reduce + back-projection)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": measured_traffic(), "peak_source": peak_src, "kernel_ms": kern_ms,
                         "algorithmic_bytes_per_launch": BYTES_FUSED * B * K},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms, "steps": e2e_steps, "chunks": C, "loss_matches_resident_step": e2e_loss_ok,
                    "exchange": "one NCCL all-gather per step" if world > 1 else "none",
                    "pcie": {"h2d_only_ms": pcie_ms, "h2d_GBps_per_gpu": h2d / (pcie_ms * 1e-3) / 1e9,
                             "h2d_GBps_all_gpus": world * h2d / (pcie_ms * 1e-3) / 1e9,
                             "note": "plain cudaMemcpyAsync of the same pinned buffers, all ranks at once, max over ranks: the box's host->device roofline at this N"},
                    "pcie_frac": pcie_ms / e2e_ms},
            "gpu_launches": launches,
            "clocks": clk.summary(),
        }
        if aten is not None:
            line["aten_baseline"] = aten
            if "value" in aten:
                line["aten_baseline"]["device_value_over_aten"] = value / aten["value"]
                line["aten_baseline"]["device_value_over_aten_loss_only"] = value / aten["loss_only_value"]
        if xcheck is not None:
            line["exchange_check"] = xcheck
        if extras:
            line["extra_workloads"] = extras
        if cpu_pipe is not None:
            torch.set_num_threads(cpu_cores)
            dt, passes = time_cpu_pipeline(cpu_pipe)
            cpu_pipe.close()
            line["cpu_baseline"] = cpu_baseline_record(cpu_n * K / dt, cpu_n, passes, cpu_procs, torch.get_num_threads(), cpu_cores)
        _emit(json.dumps(line))
    failed = xcheck is not None and not xcheck.get("all_ranks_ok", True)
    if not e2e_loss_ok:
        failed = True
    if world > 1:
        # leave without tearing NCCL down: destroying a communicator whose collectives were captured in a CUDA graph
        # can block; every rank has finished its work and rank 0 has printed, so a barrier and a plain exit are enough
        fence()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(1 if failed else 0)
    if failed:
        sys.exit(1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="images per GPU")
    ap.add_argument("--ref-sample", type=int, default=0, help="images per CPU-baseline pass (0: max(64, 8 per worker process))")
    ap.add_argument("--ref-procs", type=int, default=0, help="worker processes of the CPU arm (0: every usable core, 1: single process)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-aten-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra_workloads (configs 3 as written, 4, 5)")
    ap.add_argument("--no-exchange-check", action="store_true", help="N>1: skip the exchange parity check and stress")
    ap.add_argument("--stress-steps", type=int, default=20000, help="N>1: in-band exchange steps replayed with a changing tag")
    ap.add_argument("--e2e-chunks", type=int, default=8, help="pieces the host->device copy of a step is cut into (kernels overlap the copy)")
    ap.add_argument("--no-affinity", action="store_true", help="N>1: do not bind the rank to the GPU's NUMA-local CPUs")
    ap.add_argument("--defer", type=int, default=int(os.environ.get("POSE_B200_EXCHANGE_DEFER", "1")),
                    help="N>1, peer exchange: 1 = a step's epilogue completes the PREVIOUS step's exchange (ranks may drift), 0 = lock-step")
    ap.add_argument("--nccl", action="store_true", help="N>1: use the NCCL all-gather exchange instead of the peer-memory epilogue")
    ap.add_argument("--no-graph", action="store_true", help="run the timed steps eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: anything libraries print (e.g. NCCL's version banner) is sent to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    global _emit
    def _emit(line):
        sys.stdout.flush()
        os.write(real_stdout, (line + "\n").encode())
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
