"""Build the oracle's C restatements (TEST INFRASTRUCTURE ONLY): gcc -> oracle/_build/*.so.

    python -m oracle.build_native

-ffp-contract=off / no fast-math: every operation in the C files is meant literally (fmaf where the original fuses).
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build")
TARGETS = {"libaten_sigmoid.so": ["csrc/aten_sigmoid.c"]}


def build(force=False):
    os.makedirs(OUT, exist_ok=True)
    built = {}
    for lib, srcs in TARGETS.items():
        dst = os.path.join(OUT, lib)
        paths = [os.path.join(HERE, s) for s in srcs]
        if force or not os.path.exists(dst) or any(os.path.getmtime(p) > os.path.getmtime(dst) for p in paths):
            cmd = ["gcc", "-O2", "-mfma", "-fno-fast-math", "-ffp-contract=off", "-shared", "-fPIC", "-o", dst] + paths + ["-lm"]
            subprocess.run(cmd, check=True)
        built[lib] = dst
    return built


def aten_sigmoid_lib():
    import ctypes
    lib = ctypes.CDLL(build()["libaten_sigmoid.so"])
    lib.pose_aten_sigmoid_array.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long]
    lib.pose_aten_sigmoid_array.restype = None
    return lib


def aten_sigmoid(x):
    """np.float32 array -> the C restatement of torch.sigmoid (CPU, vectorised body) applied elementwise."""
    import ctypes

    import numpy as np
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.empty_like(x)
    aten_sigmoid_lib().pose_aten_sigmoid_array(x.ctypes.data_as(ctypes.c_void_p), y.ctypes.data_as(ctypes.c_void_p), x.size)
    return y


if __name__ == "__main__":
    print(build(force=True))
