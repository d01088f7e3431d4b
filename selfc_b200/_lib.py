"""ctypes binding of the C-ABI in include/selfc_b200.h.  There is no fallback: a missing library is an error."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libselfc_b200.so")

MODE_FP32 = 0
MODE_BF16 = 1
MODE_BF16X3 = 2      # (hi, lo) bf16 pairs, three tcgen05 MMAs per product: the <=1e-3 numerics gate on tensor cores
NUM_PARAMS = 354

_lib = None

_vp, _i, _u64, _sz = C.c_void_p, C.c_int, C.c_uint64, C.c_size_t
SIGNATURES = {
    "selfc_version": (_i, []),
    "selfc_last_error": (C.c_char_p, []),
    "selfc_ctx_create": (_i, [C.POINTER(_vp), _i, _i]),
    "selfc_ctx_destroy": (_i, [_vp]),
    "selfc_ctx_mode": (_i, [_vp]),
    "selfc_ctx_load_weights": (_i, [_vp, C.POINTER(_vp), _i, _vp]),
    "selfc_workspace_bytes": (_sz, [_vp, _i, _i, _i, _i]),
    "selfc_down": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _sz, _vp]),
    "selfc_up": (_i, [_vp, _vp, _vp, _u64, _u64, _vp, _vp, _i, _i, _i, _i, _vp, _sz, _vp]),
    "selfc_down_u8": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _sz, _vp]),
    "selfc_up_u8": (_i, [_vp, _vp, _vp, _u64, _u64, _vp, _i, _i, _i, _i, _vp, _sz, _vp]),
    "selfc_rescale_u8": (_i, [_vp, _vp, _vp, _u64, _u64, _vp, _vp, _i, _i, _i, _i, _vp, _sz, _vp]),
    "selfc_frames_from_u8": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "selfc_frames_to_u8": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "selfc_fa_fwd": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "selfc_fa_rev": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "selfc_rgb_to_y": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "selfc_frame_metrics": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "selfc_fa2_fwd": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "selfc_fa2_rev": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "selfc_haar_fwd": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "selfc_haar_rev": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "selfc_quantize": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "selfc_gaussian_down": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "selfc_d2dt": (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _i, _vp, _sz, _vp]),
    "selfc_d2dt_backward": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _sz, _vp]),
    "selfc_invblock_backward": (_i, [_vp, _i, _i, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _sz, _vp]),
    "selfc_train_tape_bytes": (_sz, [_i, _i, _i, _i]),
    "selfc_adam_step": (_i, [_vp, _vp, _i, C.c_longlong, _vp, _vp, _vp, _vp, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                            C.c_float, C.c_float, _i, _vp]),
    "selfc_train_grads": (_i, [_vp, _vp, _vp, _vp, _u64, _u64, _vp, _i, _vp, _i, _i, _i, _i, _vp, _sz, _vp, _sz, _vp]),
    "selfc_global_agg_backward": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _sz, _vp, _sz, _vp]),
    "selfc_head_sampler_backward": (_i, [_vp, _vp, _vp, _vp, _u64, _u64, _vp, _vp, _i, _i, _i, _i, _vp, _sz, _vp, _sz, _vp]),
    "selfc_conv3x3": (_i, [_vp, _i, _i, _vp, _vp, _i, _i, _i, _i, _vp, _sz, _vp]),
    "selfc_global_agg": (_i, [_vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _sz, _vp]),
    "selfc_gmm_sample": (_i, [_vp, _vp, _u64, _u64, _vp, _i, _i, _i, _i, _vp]),
    "selfc_gmm_sample_planar": (_i, [_vp, _vp, _u64, _u64, _vp, _i, _i, _i, _i, _i, _vp]),
    "selfc_export_eps": (_i, [_vp, _u64, _u64, _i, _i, _i, _i, _vp]),
    "selfc_dense_fused_schedule": (_i, [_i, C.POINTER(_i)]),
    "selfc_wgrad_geometry": (_i, [_i, _i, _i, _i, C.POINTER(C.c_longlong)]),
    "selfc_launch_count": (_u64, []),
    "selfc_prof_enable": (_i, [_vp, _i]),
    "selfc_prof_read": (_i, [_vp, _i, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_u64)]),
}


def lib() -> C.CDLL:
    """Load libselfc_b200.so (built in-tree by selfc_b200.build / __graft_entry__.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the sm_100a CUDA library has not been built "
                "(run `python -m selfc_b200.build`). selfc_b200 has no CPU or PyTorch fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError here = header / library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().selfc_last_error().decode(errors="replace")
        raise RuntimeError(f"selfc_b200 {what} failed (code {rc}): {msg}")
