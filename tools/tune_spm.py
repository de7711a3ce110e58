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

# name -> -D knobs (csrc/spm_kernels.cuh).  Shipped: see the POSE_SPM_UNIT_* defaults.
VARIANTS = {
    "l512_t128_b12": ["-DPOSE_SPM_UNIT_QUADS_LOSS=512", "-DPOSE_SPM_UNIT_MINB_LOSS=12"],
    "l1024_b8": ["-DPOSE_SPM_UNIT_MINB_LOSS=8"],
    "l2048_b5": ["-DPOSE_SPM_UNIT_QUADS_LOSS=2048", "-DPOSE_SPM_UNIT_MINB_LOSS=5"],
    "r1024_b10": ["-DPOSE_SPM_UNIT_QUADS_RENDER=1024", "-DPOSE_SPM_UNIT_MINB_RENDER=10"],
    "r4096_b3": ["-DPOSE_SPM_UNIT_QUADS_RENDER=4096", "-DPOSE_SPM_UNIT_MINB_RENDER=3"],
    "ro_screen_covtest": ["-DPOSE_SPM_RO_SCREEN_ALL=0"],
}


def _build_module():
    import importlib.util
    spec = importlib.util.spec_from_file_location("pose_b200_build", os.path.join(ROOT, "pytorch-pose-estimation_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def build():
    """one library per variant: only api_spm.cu is recompiled with the variant's knobs, the other units come from the object cache"""
    from concurrent.futures import ThreadPoolExecutor
    bm = _build_module()
    os.makedirs(OUT, exist_ok=True)
    bm.build()

    def one(item):
        n, k = item
        try:
            bm.build_library(os.path.join(OUT, f"spm_{n}.so"), {"api_spm.cu": k})
            return f"{n} ok"
        except RuntimeError as e:
            return f"{n} FAILED\n{str(e)[-2000:]}"
    with ThreadPoolExecutor(max_workers=max(2, (os.cpu_count() or 4) - 1)) as pool:
        for line in pool.map(one, VARIANTS.items()):
            print(line, flush=True)


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
