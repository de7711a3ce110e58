#!/usr/bin/env python
"""Headline benchmark: heatmaps/sec of the SBP hot path (render + loss + grad + decode) on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU algorithm (oracle port), all host cores

One "step" = one pass of the hot path over one synthetic batch (BASELINE.json configs[1]: B=4096 per GPU,
K=17, 64x48 fp32 heat maps, sigma 2, 256x192 input): fused render+loss+grad+decode kernel, back-projection /
COCO-row kernel and, for N>1, the loss all-reduce and prediction all-gather (NCCL).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

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


# --------------------------------------------------------------------------------------------- CPU baseline (oracle port)
def make_cpu_pipeline(n, procs):
    """The reference's CPU algorithm on `n` images per pass: per-sample render loop -> SBPLoss fwd+bwd -> per-sample
    decode loop -> back-projection + COCO rows (utils/sbp_utils.py:33-53, :103-118, :131-164; models/loss/sbp_loss.py),
    the per-sample loops spread over `procs` worker processes (oracle/cpu_pipeline.py)."""
    from oracle import sbp_oracle as so
    from oracle.cpu_pipeline import ReferencePipeline
    return ReferencePipeline(so.make_config1_inputs(n, K, H, W), H, W, SIGMA, IN_H, IN_W, THR, procs=procs)


def ref_shape(args):
    """(worker processes, images per pass) of the CPU arm: every usable core, at least 8 images per worker."""
    from oracle.cpu_pipeline import usable_cores
    cores = usable_cores()
    procs = args.ref_procs if args.ref_procs > 0 else min(cores, 64)
    n = args.ref_sample if args.ref_sample > 0 else max(64, 8 * procs)
    return cores, procs, n


def ref_sample_text(n, passes, procs, threads):
    return (f"{n} images of the same workload x {passes} passes; oracle port of the reference's per-sample Python loops: render and "
            f"decode over {procs} worker process(es), SBPLoss fwd+bwd on {threads} torch threads")


def run_reference(args):
    """--impl reference: rank 0 times the oracle port (the reference is pure Python and cannot travel to the GPU box;
    its CPU algorithm is restated in oracle/, pinned to it by tests/golden) on all the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cores, procs, n = ref_shape(args)
    pipe = make_cpu_pipeline(n, procs)          # forks its workers before the parent's torch thread pool exists
    torch.set_num_threads(cores)
    warm = max(1, min(args.warmup, 2))
    for _ in range(warm):
        pipe.run_pass()
    t0 = time.perf_counter()
    steps = max(1, min(args.steps, 20))
    for _ in range(steps):
        pipe.run_pass()
    dt = (time.perf_counter() - t0) / steps
    pipe.close()
    value = n * K / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"SBP 256x192: {K}x{H}x{W} heat maps, sigma {SIGMA}: render + JointsMSE fwd/bwd + decode + "
                               f"back-projection; bounded sample of {n} images per step (of the B=4096 workload)",
                   "batch_per_step": n, "threads": torch.get_num_threads(), "worker_processes": procs},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": ref_sample_text(n, steps, procs, torch.get_num_threads())},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(json.dumps(line))


# --------------------------------------------------------------------------------------------- CUDA arm
def run_cuda(args):
    import torch
    import torch.distributed as dist
    import pose_b200 as pb
    from pose_b200 import dist as pd

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cpu_pipe = None
    if world == 1 and not args.no_cpu_baseline:
        # the CPU-baseline workers are forked now, before this process owns a CUDA context; they sleep until the GPU part is done
        cpu_cores, cpu_procs, cpu_n = ref_shape(args)
        cpu_pipe = make_cpu_pipeline(cpu_n, cpu_procs)
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
        """One pass of the hot path over the batch.  N=1: 2 launches (fused kernel + epilogue).  N>1: + 1 launch that
        waits for the peers' rows (stored over NVLink by every rank's epilogue) and reduces the global loss -- or, on
        the NCCL fallback, 1 all-gather + 1 reduce launch."""
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

    for _ in range(max(args.warmup, 3)):
        step()
    fence()

    # the step is launch-bound on the host (~0.3 ms of GPU work behind several Python calls): capture it in a CUDA graph
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
            loss_graph = step()
        if ex_kind == "p2p":
            ex.steps = host_steps          # capturing ran no device work: keep the host mirror of the step counter in sync
        for _ in range(3):
            graph.replay()
        if ex_kind == "p2p":
            ex.advance(3)
        fence()
    run_step = graph.replay if graph is not None else step

    launches0 = pb.launch_count()
    per_step_launches = None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        fence()
        e0.record()
        for i in range(args.steps):
            run_step()
        if graph is not None and ex_kind == "p2p":
            ex.advance(args.steps)
        if ex_kind == "p2p":
            ex.flush(global_batch)         # defer=1: the last step's exchange completes inside the timed region too
        e1.record()
        fence()
    if graph is None:
        launches = pb.launch_count() - launches0
    else:            # graph replays do not pass through the library: count the launches of one eager step and scale
        c0 = pb.launch_count()
        step()
        launches = (pb.launch_count() - c0) * args.steps
        fence()
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = global_batch * K / (ms * 1e-3)

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

    # ---- e2e: the same step through the public API from pinned HOST buffers, result back on the host
    h_logits = torch.empty(logits.shape, dtype=logits.dtype, pin_memory=True).copy_(logits)
    h_kp = torch.empty(kp.shape, dtype=kp.dtype, pin_memory=True).copy_(kp)
    h_bbox = torch.empty(bbox.shape, dtype=bbox.dtype, pin_memory=True).copy_(bbox)
    h_packed = torch.empty((global_batch, 3 * K + 1), dtype=torch.float32, pin_memory=True)
    h_loss = torch.empty((), dtype=torch.float32, pin_memory=True)
    d_logits, d_kp, d_bbox = torch.empty_like(logits), torch.empty_like(kp), torch.empty_like(bbox)

    def e2e_step():
        d_logits.copy_(h_logits, non_blocking=True)
        d_kp.copy_(h_kp, non_blocking=True)
        d_bbox.copy_(h_bbox, non_blocking=True)
        loss = step(d_logits, d_kp, d_bbox)
        if ex_kind == "p2p":
            loss = ex.flush(global_batch)     # the host wants THIS step's rows: complete it now (no-op for defer=0)
        h_packed.copy_(ex.gathered_packed(), non_blocking=True)
        h_loss.copy_(loss, non_blocking=True)

    e2e_steps = max(3, min(args.steps, 20))
    for _ in range(3):
        e2e_step()
    fence()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(e2e_steps):
        e2e_step()
    f1.record()
    fence()
    e2e_ms = f0.elapsed_time(f1) / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    h2d = h_logits.numel() * 4 + h_kp.numel() * 8 + h_bbox.numel() * 8
    d2h = h_packed.numel() * 4 + 4
    loss_host = float(h_loss)

    if rank == 0:
        peak, peak_src = peak_hbm()
        achieved = BYTES_FUSED * B * K / (kern_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"SBP 256x192 (configs[1]): B={B} per GPU x {K} joints x {H}x{W} fp32 heat maps, sigma {SIGMA}; "
                                   "fused render+loss+grad+decode + back-projection"
                                   + ({"p2p": "; rows + loss numerators exchanged by the epilogue kernel over NVLink peer memory (no NCCL call)",
                                       "nccl": "; one NCCL all-gather (rows + loss numerators + ids) per step"}.get(ex_kind, "")),
                       "batch_per_gpu": B, "global_batch": global_batch, "partition": f"images x{world}", "exchange": ex_kind, "exchange_defer": getattr(ex, "defer", 0),
                       "l2": "inputs (856 MB logits per GPU) larger than the 126 MB L2; no explicit flush",
                       "kp_dtype": "f64", "loss": loss_host, "cuda_graph": graph is not None},
            "roofline": {"bound": "hbm", "kernel": "sbp_fused_kernel<4,RENDER,GRAD,DECODE> (+ its 1-CTA loss-reduce epilogue launch)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": measured_traffic(), "peak_source": peak_src, "kernel_ms": kern_ms,
                         "algorithmic_bytes_per_launch": BYTES_FUSED * B * K},
            "e2e": {"value": global_batch * K / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms, "steps": e2e_steps},
            "gpu_launches": launches,
            "clocks": clk.summary(),
        }
        if cpu_pipe is not None:
            torch.set_num_threads(cpu_cores)
            cpu_pipe.run_pass()
            best = None
            t_start = time.perf_counter()
            passes = 0
            while passes < 3 or (time.perf_counter() - t_start < 10.0 and passes < 40):
                t0 = time.perf_counter()
                cpu_pipe.run_pass()
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
                passes += 1
            cpu_pipe.close()
            line["cpu_baseline"] = {"value": cpu_n * K / best, "unit": UNIT, "cores": cpu_cores, "kind": "port",
                                    "sample": "best pass of: " + ref_sample_text(cpu_n, passes, cpu_procs, torch.get_num_threads())}
        _emit(json.dumps(line))
    if world > 1:
        # leave without tearing NCCL down: destroying a communicator whose collectives were captured in a CUDA graph
        # can block; every rank has finished its work and rank 0 has printed, so a barrier and a plain exit are enough
        fence()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


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
    ap.add_argument("--defer", type=int, default=int(os.environ.get("POSE_B200_EXCHANGE_DEFER", "1")),
                    help="N>1, peer exchange: 1 = a step's wait kernel completes the PREVIOUS step's exchange (ranks may drift), 0 = lock-step")
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
