#!/usr/bin/env python
"""Host->device roofline of the box with default pinned buffers vs WRITE-COMBINED pinned buffers (cudaHostAllocWriteCombined), all
ranks copying at once.  torchrun --nproc-per-node N tools/h2d_wc_probe.py; output: gpurun_out/h2d_wc_probe_nN.log (rank 0)."""
import ctypes
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nbytes = 4096 * 17 * 64 * 48 * 4
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    rt = ctypes.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else ctypes.CDLL("libcudart.so")
    rt.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
    rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
    bufs = {}
    for name, flags in (("default pinned", 0), ("write-combined pinned", 4), ("portable pinned", 1)):
        p = ctypes.c_void_p()
        rc = rt.cudaHostAlloc(ctypes.byref(p), nbytes, flags)
        assert rc == 0, (name, rc)
        ctypes.memset(p, 1, nbytes)
        bufs[name] = p
    st = torch.cuda.current_stream().cuda_stream
    lines = []
    for rep in range(2):
        for name, p in bufs.items():
            for _ in range(2):
                rt.cudaMemcpyAsync(d.data_ptr(), p, nbytes, 1, st)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                rt.cudaMemcpyAsync(d.data_ptr(), p, nbytes, 1, st)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 5
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            lines.append(f"N={world} {name:24s}: {float(t):7.2f} ms per 857 MB (max over ranks) = {nbytes / float(t) / 1e6:6.1f} GB/s per GPU, {world * nbytes / float(t) / 1e6:6.1f} GB/s all")
    if int(os.environ.get("RANK", "0")) == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"h2d_wc_probe_n{world}.log"), "w") as f:
            f.write("\n".join(lines) + "\n")
        print("\n".join(lines))
    if world > 1:
        dist.barrier()
        os._exit(0)


if __name__ == "__main__":
    main()
