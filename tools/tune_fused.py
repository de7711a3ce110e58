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
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "build", "tune")
SRC = os.path.join(ROOT, "pytorch-pose-estimation_b200", "csrc", "api.cu")
NVCC = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared"]

# name -> (-D knobs, kernels worth timing for it).  Shipped: U=6/MINB=3 (grad), U_NG=8/MINB_NG=4, U_NGD=8/MINB_NGD=3, SIGMOID_SHARE=4.
GD, G, L, LD = "grad+decode", "grad", "loss", "loss+decode"
VARIANTS = {
    "share1": (["-DPOSE_SIGMOID_SHARE=1"], (L, LD)),                                            # one reciprocal per element (r01)
    "ngd_u6_m3": (["-DPOSE_FUSED_U_NGD=6"], (LD,)),
    "ngd_u4_m4": (["-DPOSE_FUSED_U_NGD=4", "-DPOSE_FUSED_MINB_NGD=4"], (LD,)),
    "ngd_u4_m3": (["-DPOSE_FUSED_U_NGD=4"], (LD,)),
    "ng_u6_m4": (["-DPOSE_FUSED_U_NG=6"], (L,)),
    "ng_u4_m4": (["-DPOSE_FUSED_U_NG=4"], (L,)),
    "ng_u8_m3": (["-DPOSE_FUSED_MINB_NG=3"], (L,)),
    "g_u4_m3": (["-DPOSE_FUSED_U=4"], (GD, G)),
    "g_u8_m3": (["-DPOSE_FUSED_U=8"], (GD, G)),
}
FLAGS = {GD: 1 | 4, G: 1, L: 0, LD: 4}


def build():
    os.makedirs(OUT, exist_ok=True)
    procs = [(n, subprocess.Popen(NVCC + k + ["-Xptxas", "-v", "-o", os.path.join(OUT, f"sbp_{n}.so"), SRC], stdout=subprocess.PIPE,
                                  stderr=subprocess.STDOUT, text=True)) for n, (k, _) in VARIANTS.items()]
    for n, p in procs:
        out, _ = p.communicate()
        spills = [ln for ln in out.splitlines() if "spill" in ln and "0 bytes spill stores, 0 bytes spill loads" not in ln]
        print(n, "ok" if p.returncode == 0 else "FAILED\n" + out[-2000:], f"({len(spills)} kernels with spills)")


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
    lut = torch.from_numpy(_gauss_template(2).astype("float32")).to(dev)
    dl = torch.empty_like(logits)
    joints = torch.empty(B, K, 3, device=dev)
    loss = torch.empty((), device=dev)
    num = torch.empty(2, dtype=torch.float64, device=dev)
    ws = torch.empty(1 << 17, dtype=torch.uint8, device=dev)
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
