"""ctypes binding of libdlsg_b200.so (the C ABI declared in include/dlsg_b200.h).

There is no fallback: if the CUDA library has not been built the import of any operator
fails with an ImportError that says how to build it (`python -c "import __graft_entry__ as g; g.build()"`).
"""
from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DLSG_LIB_PATH") or os.path.join(_HERE, "csrc", "libdlsg_b200.so")   # override: tuning variants only

DLSG_OK = 0
DLSG_EINVAL = -1
DLSG_EWORKSPACE = -2
DLSG_EUNSUPPORTED = -3
DLSG_ELAUNCH = -4

_vp = ctypes.c_void_p
_i = ctypes.c_int
_sz = ctypes.c_size_t
_f = ctypes.c_float
_shape7 = [_i] * 7  # B, E, C, T, Y, X, ndims

# name -> (restype, argtypes); mirrors include/dlsg_b200.h one to one
SIGNATURES = {
    "dlsg_version": (_i, []),
    "dlsg_launch_count": (ctypes.c_ulonglong, []),
    "dlsg_status_string": (ctypes.c_char_p, [_i]),
    "dlsg_sense_workspace_bytes": (_sz, _shape7 + [_sz]),
    "dlsg_sense_forward_c64": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _sz] + _shape7 + [_vp]),
    "dlsg_sense_adjoint_c64": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _sz] + _shape7 + [_vp]),
    "dlsg_sense_normal_c64": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _sz] + _shape7 + [_vp]),
    "dlsg_sense_normal_reg_c64": (_i, [_vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _sz] + _shape7 + [_vp]),
    "dlsg_sense_dc_step_c64": (_i, [_vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _sz] + _shape7 + [_vp]),
    "dlsg_sense_line_plan_bytes": (_sz, [_i, _i, _i]),
    "dlsg_sense_line_plan_build": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "dlsg_sense_line_plan_build_capped": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "dlsg_sense_lines_supported": (_i, [_i, _i, _i, _i]),
    "dlsg_sense_lines_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _sz]),
    "dlsg_sense_forward_lines_c64": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _sz] + [_i] * 6 + [_vp]),
    "dlsg_sense_adjoint_lines_c64": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _sz] + [_i] * 6 + [_vp]),
    "dlsg_sense_adjoint_lines_gather_c64": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _sz] + [_i] * 6 + [_vp]),
    "dlsg_sense_normal_lines_c64": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _f, _vp, _i, _vp, _vp, _sz] + [_i] * 6 + [_vp]),
    "dlsg_fftmod_c64": (_i, [_vp, ctypes.c_longlong, _i, _i, _vp]),
    "dlsg_get_mask_c64": (_i, [_vp, _vp, ctypes.c_longlong, _f, _vp]),
    "dlsg_sliding_window_c64": (_i, [_vp, _vp, ctypes.c_longlong, _i, ctypes.c_longlong, _i, _i, _f, _f, _vp]),
    "dlsg_swin_ln_partition_f32": (_i, [_vp, _vp, _vp, _f, _vp] + [_i] * 5 + [_vp, _vp, _vp]),
    "dlsg_swin_window_reverse_f32": (_i, [_vp, _vp] + [_i] * 5 + [_vp, _vp, _vp]),
    "dlsg_swin_window_attention_f32": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _f, _i, _vp, _i, _vp]),
    "dlsg_swin_linear_f32": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "dlsg_swin_window_attention_f16": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _i, _vp, _i, _vp]),
    "dlsg_swin_qkv_fused_f32": (_i, [_vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp] + [_i] * 5 + [_vp, _vp, _i, _f, _vp]),
    "dlsg_swin_mlp_f32": (_i, [_vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_longlong, _i, _i, _i, _vp]),
    "dlsg_swin_qkv_stream_f32": (_i, [_vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp] + [_i] * 5 + [_vp, _vp, _i, _f, _vp]),
    "dlsg_swin_proj_scatter_residual_f32": (_i, [_vp, _vp, _vp, _vp, _vp] + [_i] * 5 + [_vp, _vp, _vp]),
    "dlsg_swin_proj_scatter_f32": (_i, [_vp, _vp, _vp, _vp] + [_i] * 5 + [_vp, _vp, _vp]),
    "dlsg_conv3d_k3_supported": (_i, [_i, _i]),
    "dlsg_conv3d_k3_tf32x3_f32": (_i, [_vp, _vp, _vp, _vp, _f, _vp] + [_i] * 9 + [_vp]),
    "dlsg_conv3d_pointwise_tf32x3_f32": (_i, [_vp, _vp, _vp, _vp] + [_i] * 9 + [_vp]),
    "dlsg_conv3d_tapsum_f32": (_i, [_vp, _vp, _vp] + [_i] * 8 + [_vp]),
    "dlsg_conv3d_k3_fewout_f32": (_i, [_vp, _vp, _vp, _vp] + [_i] * 7 + [_vp]),
    "dlsg_conv3d_patch_tf32x3_f32": (_i, [_vp, _vp, _vp, _vp] + [_i] * 7 + [_vp]),
    "dlsg_conv3d_unpatch_tf32x3_f32": (_i, [_vp, _vp, _vp, _vp] + [_i] * 7 + [_vp]),
}

_lock = threading.Lock()
_lib = None


def load() -> ctypes.CDLL:
    """Load (once) and return the library with argtypes/restypes set."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: the sm_100a CUDA library has not been built. "
                "Build it with `python -c \"import __graft_entry__ as g; g.build()\"` from the repo root "
                "(there is no CPU or PyTorch fallback for these operators).")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def status_string(status: int) -> str:
    return load().dlsg_status_string(int(status)).decode()


def check(status: int, what: str) -> None:
    """Turn a DLSG_E* status into the Python exception the shim documents."""
    if status == DLSG_OK:
        return
    msg = f"{what}: {status_string(status)} (status {status})"
    if status in (DLSG_EINVAL, DLSG_EUNSUPPORTED):
        raise ValueError(msg)
    raise RuntimeError(msg)
