#!/usr/bin/env python
"""A/B the fused kernel's compile-time knobs (loads in flight per lane, resident CTAs per SM).

    python tools/tune_fused.py --build          # here (no GPU): nvcc one .so per variant into build/tune/
    python tools/tune_fused.py --run            # on the B200: time every variant, JSON to gpurun_out/tune_fused.json
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
VARIANTS = [(u, m) for u in (4, 6, 8) for m in (2, 3, 4)]
# read-only (validation) render variants: loads in flight per lane, resident CTAs per SM
NG_VARIANTS = [(6, 3), (6, 4), (8, 3), (8, 4), (12, 2), (12, 3), (4, 4), (5, 4), (4, 5), (5, 5), (3, 6), (4, 6)]
# TMA-staged kernel: (stages, warps per CTA, float4 per tile)
TMA_VARIANTS = [(2, 8, 256), (3, 8, 256), (4, 8, 256), (3, 4, 256), (3, 16, 256), (2, 16, 256), (3, 8, 128), (4, 8, 128), (6, 8, 128),
                (2, 8, 768), (2, 4, 768), (3, 4, 768)]


# software-pipelined loop: (stage size for the grad kernel or 0, stage size for the read-only kernels or 0, resident CTAs per SM)
PIPE_VARIANTS = [(2, 0, 3), (3, 0, 3), (4, 0, 3), (3, 0, 2), (6, 0, 2), (0, 2, 4), (0, 3, 4), (0, 3, 3), (0, 4, 3), (0, 6, 3), (0, 2, 5)]


# read-only variant with decode: (loads in flight per lane, resident CTAs per SM)
NGD_VARIANTS = [(8, 3), (6, 3), (4, 4), (6, 4), (8, 2)]


def build(which="all"):
    os.makedirs(OUT, exist_ok=True)
    procs = []
    for u, m in (NGD_VARIANTS if which == "ngd" else []):
        lib = os.path.join(OUT, f"libpose_ngd_u{u}_m{m}.so")
        cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
               f"-DPOSE_FUSED_U_NGD={u}", f"-DPOSE_FUSED_MINB_NGD={m}", "-o", lib, SRC]
        procs.append((lib, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for pg, pn, m in (PIPE_VARIANTS if which in ("all", "pipe") else []):
        lib = os.path.join(OUT, f"libpose_pipe_g{pg}_n{pn}_m{m}.so")
        knob = [f"-DPOSE_FUSED_PIPE={pg}", f"-DPOSE_FUSED_MINB={m}"] if pg else [f"-DPOSE_FUSED_PIPE_NG={pn}", f"-DPOSE_FUSED_MINB_NG={m}"]
        cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared"] + knob + ["-o", lib, SRC]
        procs.append((lib, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    if which == "pipe":
        which = "none"
    for u, m in (VARIANTS if which == "all" else []):
        lib = os.path.join(OUT, f"libpose_u{u}_m{m}.so")
        cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
               f"-DPOSE_FUSED_U={u}", f"-DPOSE_FUSED_MINB={m}", "-o", lib, SRC]
        procs.append((lib, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for u, m in (NG_VARIANTS if which in ("all", "ng") else []):
        lib = os.path.join(OUT, f"libpose_ng_u{u}_m{m}.so")
        cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
               f"-DPOSE_FUSED_U_NG={u}", f"-DPOSE_FUSED_MINB_NG={m}", "-o", lib, SRC]
        procs.append((lib, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for st, w, tv in (TMA_VARIANTS if which == "all" else []):
        lib = os.path.join(OUT, f"libpose_tma_s{st}_w{w}_t{tv}.so")
        cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
               f"-DPOSE_TMA_STAGES={st}", f"-DPOSE_TMA_WARPS={w}", f"-DPOSE_TMA_TILE_VEC={tv}", "-o", lib, SRC]
        procs.append((lib, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for lib, p in procs:
        out, _ = p.communicate()
        print(os.path.basename(lib), "ok" if p.returncode == 0 else "FAILED\n" + out)


def run(reps, which="all"):
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
    res = {}
    ref = None
    jobs = []
    if which == "pipe":
        base = os.path.join(ROOT, "pytorch-pose-estimation_b200", "libpose_b200.so")
        jobs += [("default[grad+decode]", base, 1 | 4), ("default[loss]", base, 0), ("default[loss+decode]", base, 4)]
        for pg, pn, m in PIPE_VARIANTS:
            lib = os.path.join(OUT, f"libpose_pipe_g{pg}_n{pn}_m{m}.so")
            if pg:
                jobs.append((f"PIPE{pg}_M{m}[grad+decode]", lib, 1 | 4))
            else:
                jobs += [(f"PIPE_NG{pn}_M{m}[loss]", lib, 0), (f"PIPE_NG{pn}_M{m}[loss+decode]", lib, 4)]
    if which == "ngd":
        jobs += [(f"NGD_U{u}_M{m}[loss+decode]", os.path.join(OUT, f"libpose_ngd_u{u}_m{m}.so"), 4) for u, m in NGD_VARIANTS]
    if which == "ab":      # every build/tune/ab_*.so against the in-tree library, all four render variants
        import glob
        base = os.path.join(ROOT, "pytorch-pose-estimation_b200", "libpose_b200.so")
        libs = [("tree", base)] + [(os.path.basename(f)[3:-3], f) for f in sorted(glob.glob(os.path.join(OUT, "ab_*.so")))]
        for fl, tag in ((1 | 4, "grad+decode"), (1, "grad"), (0, "loss"), (4, "loss+decode")):
            jobs += [(f"{n}[{tag}]", f, fl) for n, f in libs]
    if which == "all":
        jobs += [(f"U{u}_M{m}", os.path.join(OUT, f"libpose_u{u}_m{m}.so"), 1 | 4) for u, m in VARIANTS]
        jobs += [(f"TMA_S{st}_W{w}_T{tv}", os.path.join(OUT, f"libpose_tma_s{st}_w{w}_t{tv}.so"), 1 | 4 | 8) for st, w, tv in TMA_VARIANTS]
    for fl, tag in (((0, "loss"), (4, "loss+decode")) if which in ("all", "ng") else ()):
        jobs += [(f"NG[{tag}]_U{u}_M{m}", os.path.join(OUT, f"libpose_ng_u{u}_m{m}.so"), fl) for u, m in NG_VARIANTS]
    refs = {}
    for name, path, flags in jobs:
        if not os.path.exists(path):
            continue
        L = ctypes.CDLL(path)
        fn = L.pose_sbp_fused
        fn.restype, fn.argtypes = _cabi.SIGNATURES["pose_sbp_fused"]

        def call():
            rc = fn(_cabi.ptr(logits), None, _cabi.ptr(kp), 1, 2.0, _cabi.ptr(lut), 15, _cabi.ptr(dl), None, _cabi.ptr(loss), _cabi.ptr(num),
                    _cabi.ptr(joints), 0.25, 4.0, B, K, H, W, 5.0, 1.0, 1.0 / (2 * K * B), flags, None, None, 0, 0, None, _cabi.ptr(ws), ws.numel(), st)
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
        ref = refs.setdefault(flags & 5, sig)
        nbytes = (24576 if flags & 1 else 12288) + 8 + (12 if flags & 4 else 0)
        res[name] = {"ms": min(ts), "GBps": nbytes * B * K / (min(ts) * 1e-3) / 1e9, "same_result": sig == ref}
        print(f"{name:30s}: {min(ts)*1e3:7.1f} us  {res[name]['GBps']:7.1f} GB/s  same={sig == ref} {sig}", flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "tune_fused.json"), "w"), indent=1)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--build", action="store_true")
    ap.add_argument("--run", action="store_true")
    ap.add_argument("--reps", type=int, default=40)
    ap.add_argument("--which", default="all", choices=["all", "ng", "pipe", "ab", "ngd"], help="ng: only the read-only (no-grad) variants")
    a = ap.parse_args()
    if a.build:
        build(a.which)
    if a.run:
        run(a.reps, a.which)
