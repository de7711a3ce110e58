"""CPU-side checks of the C-ABI library: it builds for sm_100a, loads, exports every symbol include/pose_b200.h
declares with the arity the ctypes table expects, and its host-only entry points work.  No compute calls (no GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "pose_b200.h")


@pytest.fixture(scope="module")
def pb():
    import __graft_entry__ as ge
    ge.build()
    import pose_b200
    return pose_b200


def _declared():
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    src = re.sub(r"^\s*#.*$", "", src, flags=re.M)
    decls = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(pose_\w+)\s*\(([^)]*)\)\s*;", src):
        args = m.group(3).strip()
        decls[m.group(2)] = 0 if args in ("", "void") else len(args.split(","))
    return decls


def test_header_symbols_are_exported_and_bound(pb):
    decls = _declared()
    assert len(decls) >= 16
    handle = ctypes.CDLL(pb.LIB_PATH)
    for name, nargs in decls.items():
        assert hasattr(handle, name), f"{name} declared in pose_b200.h but not exported"
        assert name in pb._cabi.SIGNATURES, f"{name} has no ctypes signature"
        assert len(pb._cabi.SIGNATURES[name][1]) == nargs, name
    assert set(pb._cabi.SIGNATURES) == set(decls)


def test_library_is_sm100a_only(pb):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", pb.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_version_and_host_template(pb):
    L = pb.lib()
    assert L.pose_b200_version() >= 100
    from oracle.sbp_oracle import gauss_template
    for sigma in (1.0, 1.5, 2.0, 3.0, 1.25):
        want = gauss_template(sigma).astype(np.float32)
        buf = (ctypes.c_float * want.size)()
        n = L.pose_gauss_template_host(sigma, ctypes.cast(buf, ctypes.c_void_p), want.size)
        assert n == want.shape[0]
        assert np.array_equal(np.frombuffer(buf, dtype=np.float32).reshape(n, n), want)
        assert np.array_equal(pb.sbp_utils._gauss_template(sigma), gauss_template(sigma))
    small = (ctypes.c_float * 4)()
    assert L.pose_gauss_template_host(2.0, ctypes.cast(small, ctypes.c_void_p), 4) == -1
    assert b"capacity" in L.pose_b200_last_error()


def test_workspace_sizes_and_cpu_tensor_rejected(pb):
    import torch
    L = pb.lib()
    # one fp64 pair per heat map + the slice sums of the two-level reduction + its counter
    assert L.pose_sbp_fused_workspace_bytes(4096, 17) >= 4096 * 17 * 16 + 34 * 16 + 4
    assert L.pose_sbp_fused_workspace_bytes(0, 17) >= 16 + 4
    assert L.pose_spm_loss_workspace_bytes(4, 17, 128) >= 4 * 35 * 4 * 16 + 4 * 2048
    with pytest.raises(pb.PoseB200Error):
        pb.decode_batch(torch.zeros(1, 1, 8, 8), 0.5)
    with pytest.raises(pb.PoseB200Error):
        pb.SPMLoss()(torch.zeros(1, 3, 8, 8), torch.zeros(1, 3, 8, 8))


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: only tests/, bench.py (cpu_baseline / --impl reference) and
    __graft_entry__.smoke() may touch it -- nothing under the package, the import shim or tools/."""
    for sub in ("pytorch-pose-estimation_b200", "pose_b200", "tools", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, sub)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    text = open(os.path.join(dirpath, f)).read()
                    assert "import oracle" not in text and "from oracle" not in text, os.path.join(dirpath, f)
    pkg = os.path.join(ROOT, "pytorch-pose-estimation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                assert "oracle" not in open(os.path.join(dirpath, f)).read(), f
