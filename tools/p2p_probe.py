#!/usr/bin/env python
"""Where does the multi-GPU step time go?  torchrun --nproc-per-node 2 tools/p2p_probe.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pose_b200 as pb  # noqa: E402
from pose_b200 import dist as pd  # noqa: E402

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, K, H, W = 4096, 17, 64, 48
gen = torch.Generator(device=dev).manual_seed(rank)
logits = torch.randn(B, K, H, W, device=dev, generator=gen) * 3
kp = torch.rand(B, K, 2, device=dev, generator=gen, dtype=torch.float64) * 48
bbox = torch.rand(B, 4, device=dev, generator=gen, dtype=torch.float64) * 300 + 40
ex = pd.PeerExchange(B, K, dev, torch.arange(B, device=dev), torch.ones(B, dtype=torch.int64, device=dev))
outs = dict(dlogits=torch.empty_like(logits), joints=torch.empty(B, K, 3, device=dev), loss=torch.empty((), device=dev))
G = B * dist.get_world_size()


def t(fn, reps=100):
    for _ in range(10):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


base = dict(keypoints=kp, sigma=2, want_grad=True, decode=True, conf_threshold=0.25, coord_scale=4.0, global_batch=G, out=outs)
res = {
    "fused (no bbox)": t(lambda: pb.sbp_fused(logits, **base)),
    "fused + local backproject": t(lambda: pb.sbp_fused(logits, bbox=bbox, input_size=(256, 192), **base)),
    "fused + p2p epilogue (no wait)": t(lambda: pb.sbp_fused(logits, bbox=bbox, input_size=(256, 192), exchange=ex, **base)),
}
import copy
import ctypes
class _Ex:      # descriptor variants: where do the epilogue's stores go?
    def __init__(self, desc): self.desc = desc
other = 1 - rank
d_local = type(ex.desc)(); ctypes.memmove(ctypes.byref(d_local), ctypes.byref(ex.desc), ctypes.sizeof(ex.desc)); d_local.peer_base[other] = ex.desc.peer_base[rank]
d_remote = type(ex.desc)(); ctypes.memmove(ctypes.byref(d_remote), ctypes.byref(ex.desc), ctypes.sizeof(ex.desc)); d_remote.peer_base[rank] = ex.desc.peer_base[other]
plain = torch.zeros(ex.buf.numel(), dtype=torch.uint8, device=dev)      # ordinary cudaMalloc memory, not symmetric
d_plain = type(ex.desc)(); ctypes.memmove(ctypes.byref(d_plain), ctypes.byref(ex.desc), ctypes.sizeof(ex.desc)); d_plain.peer_base[0] = plain.data_ptr(); d_plain.peer_base[1] = plain.data_ptr()
res["p2p epilogue, both destinations LOCAL symmetric"] = t(lambda: pb.sbp_fused(logits, bbox=bbox, input_size=(256, 192), exchange=_Ex(d_local), **base))
res["p2p epilogue, both destinations REMOTE"] = t(lambda: pb.sbp_fused(logits, bbox=bbox, input_size=(256, 192), exchange=_Ex(d_remote), **base))
res["p2p epilogue, both destinations plain local"] = t(lambda: pb.sbp_fused(logits, bbox=bbox, input_size=(256, 192), exchange=_Ex(d_plain), **base))


def full():
    pb.sbp_fused(logits, bbox=bbox, input_size=(256, 192), exchange=ex, **base)
    ex.finish(G)
res["fused + p2p epilogue + wait/reduce"] = t(full)
# small-kernel costs: decode only as a spacer so the epilogue is not hidden behind launch gaps
res["epilogue-only proxy: backproject kernel"] = t(lambda: pb.backproject_packed(outs["joints"], bbox, (256, 192)))
if rank == 0:
    for k, v in res.items():
        print(f"{k:50s} {v:8.1f} us")
    print("exchange error flag:", ex.error(), "multicast:", ex.multicast, "mc_ptr:", hex(int(getattr(ex.handle, "multicast_ptr", 0) or 0)))
dist.barrier()
os._exit(0)
