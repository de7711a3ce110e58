#!/usr/bin/env python
"""Is a read-only SBP kernel's time a property of the kernel or of how long it has been running?  Times the validation step
(render + loss + decode, no grad; issue-bound) and the loss-only form in eager bursts of different lengths and from CUDA graphs with
different numbers of copies, sampling the SM clock (NVML) while the GPU is busy.  Output: gpurun_out/burst_vs_sustained.log"""
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pose_b200 as pb  # noqa: E402
import pynvml  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    B, K, H, W = 4096, 17, 64, 48
    gen = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(B, K, H, W, device=dev, generator=gen) * 3.0
    kp = torch.stack([torch.rand(B, K, device=dev, generator=gen, dtype=torch.float64) * W,
                      torch.rand(B, K, device=dev, generator=gen, dtype=torch.float64) * H], dim=-1)
    kp[torch.rand(B, K, device=dev, generator=gen) >= 0.85] = -1.0
    fo = dict(joints=torch.empty(B, K, 3, device=dev), loss=torch.empty((), device=dev), loss_num=torch.empty(2, dtype=torch.float64, device=dev))
    forms = {"loss+decode": lambda: pb.sbp_fused(x, keypoints=kp, sigma=2, want_grad=False, decode=True, conf_threshold=0.25, coord_scale=4.0, out=fo),
             "loss": lambda: pb.sbp_fused(x, keypoints=kp, sigma=2, want_grad=False, out=fo)}
    out = open(os.path.join(ROOT, "gpurun_out", "burst_vs_sustained.log"), "w")

    def say(s):
        print(s, flush=True)
        out.write(s + "\n")

    def timed(run, n):
        clocks, stop = [], threading.Event()

        def sample():
            while not stop.is_set():
                clocks.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
                time.sleep(0.002)
        th = threading.Thread(target=sample)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        th.start()
        a.record()
        for _ in range(n):
            run()
        b.record()
        b.synchronize()
        stop.set()
        th.join()
        mhz = sorted(c[0] for c in clocks)
        return a.elapsed_time(b) / n, mhz[len(mhz) // 2], mhz[0], max(c[1] for c in clocks)

    for name, fn in forms.items():
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        for n in (20, 60, 200, 1000, 4000):
            time.sleep(0.5)                                   # let the GPU idle between measurements
            ms, med, lo, pw = timed(fn, n)
            say(f"{name:12s} eager burst of {n:5d} calls ({ms*n:7.1f} ms busy): {ms*1e3:7.1f} us/call  SM clock median {med} min {lo} MHz, power max {pw:.0f} W")
        for inner in (1, 6, 16):
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
            for reps in (max(1, 20 // inner), max(1, 1000 // inner)):
                time.sleep(0.5)
                ms, med, lo, pw = timed(g.replay, reps)
                say(f"{name:12s} graph of {inner:2d} copies x {reps:4d} replays ({ms*reps:7.1f} ms busy): {ms/inner*1e3:7.1f} us/call  SM clock median {med} min {lo} MHz, power max {pw:.0f} W")
    out.close()


if __name__ == "__main__":
    main()
