"""ctypes binding of libpose_b200.so (include/pose_b200.h).  No CPU fallback: a missing library or a
non-CUDA tensor is an error."""
import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# POSE_B200_LIB points at another build of the SAME library (tools/tune_*.py compile variants with different -D knobs)
LIB_PATH = os.environ.get("POSE_B200_LIB") or os.path.join(_HERE, "libpose_b200.so")

KP_F32, KP_F64 = 0, 1
F_GRAD, F_TARGET_OUT, F_DECODE, F_TMA, F_SIGMOID_CUDA = 1, 2, 4, 8, 16
F_HEAD_LOGITS_OUT, F_HEAD_NO_RESIDUAL = 32, 64
SIGMOID_ATEN_CPU, SIGMOID_ATEN_CUDA = 0, 1

# Which torch.sigmoid the decoders reproduce bit for bit when ranking near-equal logits (include/pose_b200.h).
# "cpu": the reference on CPU tensors (what the golden vectors hold) -- the default, it is what the parity tests pin;
# "cuda": the reference on CUDA tensors (what the Lightning modules ran).  POSE_B200_SIGMOID_REF overrides the default.
DEFAULT_SIGMOID_REF = os.environ.get("POSE_B200_SIGMOID_REF", "cpu")
# stage heat maps through shared memory with the TMA engine (cp.async.bulk + mbarrier) instead of per-thread LDG: the fast path
# (maps in flight live in shared memory, not in registers); POSE_B200_TMA=0 selects the register-staged kernels
DEFAULT_TMA = os.environ.get("POSE_B200_TMA", "1") == "1"


def sigmoid_ref_code(name=None):
    name = DEFAULT_SIGMOID_REF if name is None else name
    if name in ("cpu", "aten_cpu", SIGMOID_ATEN_CPU):
        return SIGMOID_ATEN_CPU
    if name in ("cuda", "aten_cuda", SIGMOID_ATEN_CUDA):
        return SIGMOID_ATEN_CUDA
    raise ValueError(f"sigmoid_ref must be 'cpu' or 'cuda', got {name!r}")

_c = ctypes
_vp, _i, _u, _f, _d, _ull = _c.c_void_p, _c.c_int, _c.c_uint, _c.c_float, _c.c_double, _c.c_ulonglong

# name -> (restype, argtypes); mirrors include/pose_b200.h one to one
SIGNATURES = {
    "pose_b200_version": (_i, []),
    "pose_b200_source_hash": (_c.c_char_p, []),
    "pose_b200_last_error": (_c.c_char_p, []),
    "pose_b200_launch_count": (_ull, []),
    "pose_gauss_template_host": (_i, [_d, _vp, _i]),
    "pose_gauss_template_padded_host": (_i, [_d, _vp, _i]),
    "pose_sbp_render": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _d, _vp, _i, _vp]),
    "pose_sbp_fused_workspace_bytes": (_ull, [_i, _i]),
    "pose_sbp_fused": (_i, [_vp, _vp, _vp, _i, _d, _vp, _i, _vp, _vp, _vp, _vp, _vp, _f, _f,
                            _i, _i, _i, _i, _f, _f, _d, _u, _vp, _vp, _i, _i, _vp, _vp, _ull, _vp]),
    "pose_exchange_layout": (_ull, [_vp]),
    "pose_exchange_finish": (_i, [_vp, _d, _d, _d, _vp, _vp]),
    "pose_exchange_flush": (_i, [_vp, _d, _d, _d, _vp, _vp]),
    "pose_loss_reduce": (_i, [_vp, _i, _c.c_longlong, _d, _d, _d, _vp, _vp, _vp]),
    "pose_scale_grad": (_i, [_vp, _vp, _ull, _vp]),
    "pose_sbp_decode": (_i, [_vp, _vp, _i, _i, _i, _i, _f, _i, _f, _i, _i, _vp]),
    "pose_sbp_decode_flip": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _f, _i, _vp]),
    "pose_sbp_backproject": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "pose_spm_render": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _d, _vp, _i, _vp, _ull, _vp]),
    "pose_spm_loss_workspace_bytes": (_ull, [_i, _i, _i]),
    "pose_spm_loss": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _f, _d, _i, _vp, _ull, _vp]),
    "pose_spm_fused_workspace_bytes": (_ull, [_i, _i, _i]),
    "pose_spm_fused": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _d, _vp, _i, _f, _f, _d, _u, _vp, _ull, _vp]),
    "pose_spm_decode": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _d, _i, _i, _f, _vp]),
    "pose_spm_rescale": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _f, _vp]),
    "pose_spm_gather": (_i, [_vp, _vp, _vp, _i, _i, _i, _d, _vp]),
    "pose_spm_gather_chain": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _d, _vp]),
    "pose_oks_matrix": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _c.c_longlong, _i, _vp]),
    "pose_oks_match_workspace_bytes": (_ull, [_i, _i, _i]),
    "pose_oks_match": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _ull, _vp]),
    "pose_ap_accumulate_workspace_bytes": (_ull, [_i, _i, _i]),
    "pose_ap_accumulate": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _ull, _vp]),
    "pose_sbp_head_workspace_bytes": (_ull, [_i, _i, _i]),
    "pose_sbp_head_fused": (_i, [_vp, _vp, _vp, _i, _d, _vp, _i, _vp, _vp, _vp, _vp, _vp, _f, _f, _i, _i, _i, _i, _i, _f, _f, _d, _u,
                                 _vp, _vp, _i, _i, _i, _vp, _ull, _vp]),
    "pose_sigmoid_ref_eval": (_i, [_vp, _vp, _ull, _i, _vp]),
    "pose_sigmoid_window_check": (_i, [_vp, _i, _vp]),
}

MAX_PEERS = 16
EXCHANGE_SLOTS = 4


class ExchangeDesc(ctypes.Structure):
    """pose_exchange_t (include/pose_b200.h)."""
    _fields_ = [("world", _i), ("rank", _i), ("batch_local", _i), ("num_keypoints", _i), ("row_stride", _i), ("defer", _i),
                ("peer_base", _vp * MAX_PEERS),
                ("off_ctrl", _ull), ("off_flags", _ull), ("off_rows", _ull * EXCHANGE_SLOTS), ("off_nums", _ull * EXCHANGE_SLOTS),
                ("off_ids", _ull * EXCHANGE_SLOTS),
                ("ids_local", _vp), ("multicast_base", _vp), ("loss_prev", _vp)]


_lib = None
_lock = threading.Lock()


class PoseB200Error(RuntimeError):
    pass


def lib():
    """The loaded C-ABI library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise PoseB200Error(
                        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(nvcc, sm_100a).  pose_b200 has no CPU or PyTorch fallback.")
                if not os.environ.get("POSE_B200_LIB"):        # an explicitly named variant build is the caller's business
                    from . import build as _build
                    have, want = _build.embedded_hash(LIB_PATH), _build.source_hash()
                    if have != want:
                        raise PoseB200Error(
                            f"{LIB_PATH} is stale: it was built from sources with hash {have}, the tree has {want}.  Rebuild it "
                            "with `python -c 'import __graft_entry__ as g; g.build()'`.")
                handle = ctypes.CDLL(LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(handle, name)
                    fn.restype, fn.argtypes = res, args
                _lib = handle
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().pose_b200_last_error().decode(errors="replace")
        raise PoseB200Error(f"{what or 'pose_b200'} failed (rc={rc}): {msg}")


def launch_count():
    return int(lib().pose_b200_launch_count())


def require_cuda(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise PoseB200Error(f"{name} must be a CUDA tensor: pose_b200 runs only its sm_100a kernels (no CPU fallback)")
    return t


def dense(t, name, dtype=torch.float32):
    """CUDA, contiguous, expected dtype -- converts layout/dtype on device if needed, never moves to the CPU."""
    require_cuda(t, name)
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def stream_ptr(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


_ws = {}


def workspace(device, nbytes):
    """Per-(device, stream) scratch for the loss partials; allocated once, reused (stream-ordered use only)."""
    key = (device.index if device.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(device).cuda_stream)
    buf = _ws.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 16), dtype=torch.uint8, device=device)
        _ws[key] = buf
    return buf
