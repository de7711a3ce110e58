"""Build recipe for libpose_b200.so (the C-ABI CUDA library), in-tree, sm_100a only.

    python pytorch-pose-estimation_b200/build.py          # or __graft_entry__.build()

Whether the library is up to date is decided by CONTENT, not by time stamps: a sha256 over the sources, the header and
the compiler flags is compiled into the binary (`pose_b200_source_hash()`), and both this recipe and the loader
(`_cabi.lib()`) compare it with the hash of the tree they see -- a snapshot that flattens mtimes cannot make a stale
binary pass for a fresh one.
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc", "api.cu")
DEPS = sorted(os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc")) if f.endswith((".cu", ".cuh", ".h"))) + [
    os.path.join(ROOT, "include", "pose_b200.h")]
LIB = os.path.join(HERE, "libpose_b200.so")
HASH_TAG = b"POSE_B200_SOURCE_HASH="

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=...)")


def source_hash():
    """sha256 over the file names + contents of every source the library is compiled from, and the compiler flags."""
    h = hashlib.sha256()
    for d in DEPS:
        h.update(os.path.basename(d).encode() + b"\0")
        with open(d, "rb") as f:
            h.update(f.read())
        h.update(b"\0")
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def embedded_hash(path=LIB):
    """The hash compiled into an existing library (read from the file, without loading it), or None."""
    try:
        with open(path, "rb") as f:
            blob = f.read()
    except OSError:
        return None
    i = blob.find(HASH_TAG)
    if i < 0:
        return None
    return blob[i + len(HASH_TAG):i + len(HASH_TAG) + 64].decode("ascii", "replace")


def needs_build():
    return embedded_hash() != source_hash()


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    tmp = LIB + ".tmp"
    cmd = [_nvcc()] + NVCC_FLAGS + ['-DPOSE_B200_SOURCE_HASH_STR="' + source_hash() + '"'] + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp, SRC]
    proc = subprocess.run(cmd, cwd=HERE, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    os.replace(tmp, LIB)
    if verbose:
        sys.stderr.write(proc.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
