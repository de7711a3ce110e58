#!/usr/bin/env python
"""A/B the SBP fused kernel's compile-time knobs (loads in flight per lane, resident CTAs per SM, shared-reciprocal sigmoid).

    python tools/tune_fused.py --build          # here (no GPU): nvcc one .so per variant into build/tune/
    python tools/tune_fused.py --run            # on the B200: time every variant, JSON to gpurun_out/tune_fused.json

Every variant is a full libpose_b200.so called through the C ABI directly (ctypes), so the product code path is measured.
Four kernels per library: grad+decode (headline), grad, loss (validation, read-only), loss+decode (validation, read-only).
"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "build", "tune")

# name -> (-D knobs, kernels worth timing for it).  Shipped ("tree"): see the POSE_TMA_* defaults in csrc/sbp_kernels.cuh
# (read-only fused kernels: 1 warp per map, 2 maps per CTA, 8 CTAs per SM).
GD, G, L, LD = "grad+decode", "grad", "loss", "loss+decode"
VARIANTS = {
    "g_m2_b8": (["-DPOSE_TMA_MPC_GRAD=2", "-DPOSE_TMA_MINB_GRAD=8"], (GD, G)),
    "g_w1m2": (["-DPOSE_TMA_WPM_GRAD=1", "-DPOSE_TMA_MPC_GRAD=2", "-DPOSE_TMA_MINB_GRAD=16"], (GD, G)),
    "nopdl": (["-DPOSE_TMA_FUSED_PDL=0"], (GD, G, L, LD)),
    "ro_w2m2_b8": (["-DPOSE_TMA_WPM_RO=2"], (L, LD)),
    "ro_w2m4_b4": (["-DPOSE_TMA_WPM_RO=2", "-DPOSE_TMA_MPC_RO=4", "-DPOSE_TMA_MINB_RO=4"], (L, LD)),
    "ro_w1m1_b16": (["-DPOSE_TMA_MPC_RO=1", "-DPOSE_TMA_MINB_RO=16"], (L, LD)),
    "ro_w1m3_b5": (["-DPOSE_TMA_MPC_RO=3", "-DPOSE_TMA_MINB_RO=5"], (L, LD)),
    "ro_w1m4_b4": (["-DPOSE_TMA_MPC_RO=4", "-DPOSE_TMA_MINB_RO=4"], (L, LD)),
}
FLAGS = {GD: 1 | 4 | 8, G: 1 | 8, L: 0 | 8, LD: 4 | 8}        # | 8: POSE_F_TMA (bulk-async staged kernels)


def _build_module():
    import importlib.util
    spec = importlib.util.spec_from_file_location("pose_b200_build", os.path.join(ROOT, "pytorch-pose-estimation_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def build():
    """one library per variant: only api_sbp.cu is recompiled with the variant's knobs, the other units come from the object cache"""
    from concurrent.futures import ThreadPoolExecutor
    bm = _build_module()
    os.makedirs(OUT, exist_ok=True)
    bm.build()                                            # the tree library first: fills the object cache for the shared units

    def one(item):
        n, (k, _) = item
        try:
            bm.build_library(os.path.join(OUT, f"sbp_{n}.so"), {"api_sbp.cu": k})
            log = bm.ptxas_log("api_sbp.cu", k)
            spills = [ln for ln in log.splitlines() if "spill" in ln and "0 bytes spill stores, 0 bytes spill loads" not in ln]
            return f"{n} ok ({len(spills)} kernels with spills)"
        except RuntimeError as e:
            return f"{n} FAILED\n{str(e)[-2000:]}"
    with ThreadPoolExecutor(max_workers=max(2, (os.cpu_count() or 4) - 1)) as pool:
        for line in pool.map(one, VARIANTS.items()):
            print(line, flush=True)


def run(reps):
    import torch
    from pose_b200 import _cabi
    from pose_b200.sbp_utils import _gauss_template
    dev = torch.device("cuda", 0)
    B, K, H, W = 4096, 17, 64, 48
    gen = torch.Generator(device=dev).manual_seed(0)
    logits = torch.randn(B, K, H, W, device=dev, generator=gen) * 3
    kp = torch.stack([torch.rand(B, K, device=dev, generator=gen, dtype=torch.float64) * W,
                      torch.rand(B, K, device=dev, generator=gen, dtype=torch.float64) * H], -1)
    kp[torch.rand(B, K, device=dev, generator=gen) >= 0.85] = -1
    import numpy as np
    lut = torch.from_numpy(np.pad(_gauss_template(2).astype("float32"), ((0, 1), (4, 4)))).to(dev)
    dl = torch.empty_like(logits)
    joints = torch.empty(B, K, 3, device=dev)
    loss = torch.empty((), device=dev)
    num = torch.empty(2, dtype=torch.float64, device=dev)
    ws = torch.empty(int(_cabi.lib().pose_sbp_fused_workspace_bytes(B, K)), dtype=torch.uint8, device=dev)
    st = _cabi.stream_ptr(dev)
    res, refs = {}, {}
    tree = os.path.join(ROOT, "pytorch-pose-estimation_b200", "libpose_b200.so")
    jobs = [("tree", tree, (GD, G, L, LD))] + [(n, os.path.join(OUT, f"sbp_{n}.so"), ks) for n, (_, ks) in VARIANTS.items()]
    for name, path, kernels in jobs:
        if not os.path.exists(path):
            continue
        lib = ctypes.CDLL(path)
        fn = lib.pose_sbp_fused
        fn.restype, fn.argtypes = _cabi.SIGNATURES["pose_sbp_fused"]
        for kern in kernels:
            flags = FLAGS[kern]

            def call():
                rc = fn(_cabi.ptr(logits), None, _cabi.ptr(kp), 1, 2.0, _cabi.ptr(lut), 15, _cabi.ptr(dl), None, _cabi.ptr(loss), _cabi.ptr(num),
                        _cabi.ptr(joints), 0.25, 4.0, B, K, H, W, 5.0, 1.0, 1.0 / (2 * K * B), flags, None, None, 0, 0, None, _cabi.ptr(ws),
                        ws.numel(), st)
                assert rc == 0, rc
            for _ in range(5):
                call()
            torch.cuda.synchronize()
            ts = []
            for _ in range(3):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(reps):
                    call()
                b.record()
                b.synchronize()
                ts.append(a.elapsed_time(b) / reps)
            sig = (float(loss), float(dl.double().abs().sum()) if flags & 1 else 0.0, float(joints.double().sum()) if flags & 4 else 0.0)
            ref = refs.setdefault(kern, sig)
            close = abs(sig[0] - ref[0]) <= 1e-6 * abs(ref[0]) and abs(sig[1] - ref[1]) <= 1e-6 * abs(ref[1]) and sig[2] == ref[2]
            nbytes = (24576 if flags & 1 else 12288) + 8 + (12 if flags & 4 else 0)
            key = f"{name}[{kern}]"
            res[key] = {"us": min(ts) * 1e3, "GBps": nbytes * B * K / (min(ts) * 1e-3) / 1e9, "same_result_as_tree": close}
            print(f"{key:32s}: {min(ts)*1e3:7.1f} us  {res[key]['GBps']:7.1f} GB/s  same={close}", flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "tune_fused.json"), "w"), indent=1)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--build", action="store_true")
    ap.add_argument("--run", action="store_true")
    ap.add_argument("--reps", type=int, default=40)
    a = ap.parse_args()
    if a.build:
        build()
    if a.run:
        run(a.reps)
