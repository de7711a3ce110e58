#!/usr/bin/env python
"""A/B the fused SPM kernel's compile-time knobs (unit size per variant, resident CTAs per SM).

    python tools/tune_spm.py --build            # here (no GPU): nvcc one .so per variant into build/tune/
    python tools/tune_spm.py --run [--check]    # on the B200: time every variant (and the in-tree library) with tools/spm_skeleton.py;
                                                #   --check also runs tests/test_spm_gpu.py against each variant first

Every variant is a full libpose_b200.so selected through POSE_B200_LIB, so the product code path is the one measured.
"""
import argparse
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "build", "tune")
SRC = os.path.join(ROOT, "pytorch-pose-estimation_b200", "csrc", "api.cu")
NVCC = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared"]

# name -> -D knobs (csrc/spm_kernels.cuh).  Shipped: 8 stream + 4 patch warps per CTA, MINB=2, U=4 (grad), U_RO=4, U_RENDER=8.
VARIANTS = {
    "np2_m3": ["-DPOSE_SPM_PATCH_WARPS=2", "-DPOSE_SPM_FUSED_MINB=3"],
    "np4_m3": ["-DPOSE_SPM_FUSED_MINB=3"],
    "np8_m2": ["-DPOSE_SPM_PATCH_WARPS=8"],
    "np2_m2": ["-DPOSE_SPM_PATCH_WARPS=2"],
    "np4_m2_u8": ["-DPOSE_SPM_FUSED_U=8", "-DPOSE_SPM_FUSED_U_RO=8", "-DPOSE_SPM_FUSED_U_RENDER=16"],
    "np4_m2_u2": ["-DPOSE_SPM_FUSED_U=2", "-DPOSE_SPM_FUSED_U_RO=2", "-DPOSE_SPM_FUSED_U_RENDER=4"],
    "np1_m4": ["-DPOSE_SPM_PATCH_WARPS=1", "-DPOSE_SPM_FUSED_MINB=4"],
}


def build():
    os.makedirs(OUT, exist_ok=True)
    procs = [(n, subprocess.Popen(NVCC + k + ["-o", os.path.join(OUT, f"spm_{n}.so"), SRC], stdout=subprocess.PIPE,
                                  stderr=subprocess.STDOUT, text=True)) for n, k in VARIANTS.items()]
    for n, p in procs:
        out, _ = p.communicate()
        print(n, "ok" if p.returncode == 0 else "FAILED\n" + out)


def run(check):
    libs = [("tree", None)] + [(n, os.path.join(OUT, f"spm_{n}.so")) for n in VARIANTS if os.path.exists(os.path.join(OUT, f"spm_{n}.so"))]
    for name, path in libs:
        env = dict(os.environ)
        if path:
            env["POSE_B200_LIB"] = path
        print(f"== {name} {' '.join(VARIANTS.get(name, []))}", flush=True)
        if check:
            r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_spm_gpu.py"), "-m", "gpu", "-x", "-q"],
                               env=env, cwd=ROOT, capture_output=True, text=True)
            print("   tests:", r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:], flush=True)
            if r.returncode != 0:
                continue
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "spm_skeleton.py")], env=env, cwd=ROOT, capture_output=True, text=True)
        print("".join("   " + l + "\n" for l in r.stdout.splitlines() if l.startswith("N=")), end="", flush=True)
        if r.returncode != 0:
            print("   FAILED:", r.stderr[-300:])


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--build", action="store_true")
    ap.add_argument("--run", action="store_true")
    ap.add_argument("--check", action="store_true", help="run tests/test_spm_gpu.py against each variant before timing it")
    a = ap.parse_args()
    if a.build:
        build()
    if a.run:
        run(a.check)
