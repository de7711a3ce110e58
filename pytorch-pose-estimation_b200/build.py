"""Build recipe for libpose_b200.so (the C-ABI CUDA library), in-tree, sm_100a only.

    python pytorch-pose-estimation_b200/build.py [--force] [-v]         # or __graft_entry__.build()

The library is a handful of translation units (csrc/api_*.cu, each including the kernel headers it launches) compiled
separately -- in parallel, objects cached under build/obj/ by a hash of the unit, the headers it includes and the flags -- and
linked into one shared object.  Whether the LIBRARY is up to date is decided by content, not by time stamps: a sha256 over all
sources, the header and the compiler flags is compiled into the binary (`pose_b200_source_hash()`), and both this recipe and
the loader (`_cabi.lib()`) compare it with the hash of the tree they see -- a snapshot that flattens mtimes cannot make a stale
binary pass for a fresh one.
"""
import hashlib
import os
import re
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
UNITS = ["api_core.cu", "api_sbp.cu", "api_spm.cu", "api_oks.cu", "api_head.cu"]
DEPS = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))) + [
    os.path.join(ROOT, "include", "pose_b200.h")]
LIB = os.path.join(HERE, "libpose_b200.so")
OBJ = os.path.join(ROOT, "build", "obj")
HASH_TAG = b"POSE_B200_SOURCE_HASH="

CC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-Xfatbin", "-compress-all"]
LINK_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC"]
NVCC_FLAGS = CC_FLAGS + ["-shared"]          # (what the source hash covers)


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


_INCLUDE = re.compile(r'^\s*#\s*include\s+"([^"]+)"', re.M)


def _closure(path, seen=None):
    """The file and every file it includes with "..." (recursively), as absolute paths."""
    seen = seen if seen is not None else set()
    path = os.path.normpath(path)
    if path in seen or not os.path.exists(path):
        return seen
    seen.add(path)
    with open(path, "r") as f:
        for inc in _INCLUDE.findall(f.read()):
            _closure(os.path.join(os.path.dirname(path), inc), seen)
    return seen


def _unit_key(unit, defines):
    h = hashlib.sha256()
    for d in sorted(_closure(os.path.join(CSRC, unit))):
        h.update(os.path.basename(d).encode() + b"\0")
        with open(d, "rb") as f:
            h.update(f.read())
    h.update(" ".join(CC_FLAGS + list(defines)).encode())
    return h.hexdigest()[:20]


def _compile(unit, defines, verbose=False):
    """-> path of the (cached) object of one translation unit compiled with extra -D flags; the ptxas -v text when verbose"""
    os.makedirs(OBJ, exist_ok=True)
    obj = os.path.join(OBJ, unit.replace(".cu", "") + "-" + _unit_key(unit, defines) + ".o")
    log = obj + ".log"
    if not os.path.exists(obj):
        tmp = obj + ".tmp%d" % os.getpid()
        cmd = [_nvcc()] + CC_FLAGS + list(defines) + ["-Xptxas", "-v", "-c", "-o", tmp, os.path.join(CSRC, unit)]
        proc = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
        with open(log, "w") as f:
            f.write(proc.stderr)
        os.replace(tmp, obj)
    if verbose and os.path.exists(log):
        sys.stderr.write(open(log).read())
    return obj


def build_library(out, defines=None, verbose=False):
    """Compile (or take from the cache) every unit and link `out`.  `defines`: {unit: [-D...]} extra flags per unit -- the tuning
    tools build knob variants of one unit this way and share the objects of the others."""
    defines = dict(defines or {})
    core_defs = ['-DPOSE_B200_SOURCE_HASH_STR="' + source_hash() + '"'] + list(defines.pop("api_core.cu", []))
    jobs = [(u, core_defs if u == "api_core.cu" else list(defines.get(u, []))) for u in UNITS]
    with ThreadPoolExecutor(max_workers=len(jobs)) as pool:
        objs = list(pool.map(lambda j: _compile(j[0], j[1], verbose), jobs))
    tmp = out + ".tmp%d" % os.getpid()
    cmd = [_nvcc()] + LINK_FLAGS + ["-o", tmp] + objs
    proc = subprocess.run(cmd, cwd=HERE, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    os.replace(tmp, out)
    return out


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    return build_library(LIB, verbose=verbose)


def ptxas_log(unit, defines=()):
    """registers / spills per kernel of a unit as ptxas printed them (compiles the unit if it is not cached)"""
    obj = _compile(unit, list(defines))
    return open(obj + ".log").read()


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
