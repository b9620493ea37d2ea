"""ctypes binding of liby3d_b200.so (C ABI declared in include/y3d.h).

There is NO fallback: if the shared library is missing or a call returns non-zero this raises.  The CUDA
library is the product; PyTorch only supplies device memory and the current stream.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("Y3D_LIB_PATH") or os.path.join(_HERE, "liby3d_b200.so")  # (override: developer A/B builds)

STAGE_POSTPROCESS, STAGE_TAL_ASSIGN, STAGE_V8_LOSS, STAGE_TAL_ASSIGN3D, STAGE_DECODE_TOPK, STAGE_DD_LOSS = 1, 2, 3, 4, 5, 6
MAX_LEVELS, MAX_TOPK, MAX_DET = 4, 32, 1024

_vp, _i, _i64, _f, _d, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double, C.c_size_t
_LEVELS = [_vp, _vp, _vp, _vp, _vp, _i]  # lvl_ptr, lvl_sB, lvl_sC, lvl_hw, lvl_stride, nl

SIGNATURES = {
    "y3d_strerror": (C.c_char_p, [_i]),
    "y3d_abi_version": (_i, []),
    "y3d_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i, _i]),
    "y3d_pack_targets": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _f, _vp, _vp, _vp]),
    "y3d_decode2d": (_i, _LEVELS + [_i, _i, _i, _i, _vp, _vp]),
    "y3d_postprocess": (_i, [_vp, _i64, _i64, _i64, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "y3d_decode_topk2d": (_i, _LEVELS + [_i, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "y3d_gather_buffer_bytes": (_sz, [_i, _i, _i]),
    "y3d_decode_topk2d_sharded": (_i, _LEVELS + [_i, _i, _i, _i, _i, _i, _i, _vp, C.c_uint64, _vp, _vp, _sz, _vp]),
    "y3d_tal_assign": (_i, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _f, _f, _vp, _vp,
                            _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "y3d_v8_loss_fwd": (_i, _LEVELS + [_i, _i, _i, _vp, _i, _i, _f, _f, _f, _i, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp,
                                      _sz, _vp]),
    "y3d_train_decode": (_i, _LEVELS + [_i, _i, _i, _vp, _vp, _vp]),
    "y3d_v10_loss_fwd": (_i, [_vp] * 8 + [_i, _i, _i, _i, _vp, _i, _i, _i, _f, _f, _f, _i, _vp, _vp, _f, _vp, _vp, _vp,
                              _vp, _vp, _sz, _vp]),
    "y3d_v10_loss_fwd_sharded": (_i, [_vp] * 8 + [_i, _i, _i, _i, _vp, _i, _i, _i, _f, _f, _f, _vp, _vp, _f, _vp, _i, _i,
                                      _vp, C.c_uint64, _i, _vp, _vp, _vp, _sz, _vp]),
    "y3d_loss_exchange_resolve": (_i, [_i, _i, _i, _vp, C.c_uint64, _f, _f, _f, _f, _vp, _vp, _vp, _vp, _vp]),
    "y3d_v8_loss_finalize": (_i, [_vp, _i, _f, _f, _f, _vp, _vp]),
    "y3d_xrank_buffer_bytes": (_sz, [_i]),
    "y3d_loss_allreduce_finalize": (_i, [_vp, _i, _i, _i, _vp, C.c_uint64, _f, _f, _f, _vp, _vp, _vp, _vp]),
    "y3d_v8_loss_bwd": (_i, [_vp] * 8 + [_i, _i, _i, _i, _vp, _i, _i, _f, _f, _f, _vp, _vp, _vp, _sz, _vp]),
    "y3d_v10_loss_bwd": (_i, [_vp] * 14 + [_i, _i, _i, _i, _vp, _i, _i, _i, _f, _f, _f, _vp, _vp, _vp, _sz, _vp]),
    "y3d_dd_loss_fwd": (_i, _LEVELS + [_i, _i, _vp, _i, _vp, _vp, _i, _f, _f, _f, _i, _vp, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "y3d_dd_loss_dual_fwd": (_i, [_vp] * 8 + [_i, _i, _i, _vp, _i, _vp, _vp, _i, _i, _f, _f, _f, _i, _vp, _i, _vp, _vp, _vp, _vp,
                                  _sz, _vp]),
    "y3d_dd_loss_finalize": (_i, [_vp, _i, _vp, _vp, _vp]),
    "y3d_dd_loss_bwd": (_i, [_vp] * 8 + [_i, _i, _i, _vp, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "y3d_select_candidates": (_i, [_vp, _i64, _i64, _i, _i, _i, _i, _i, _vp, _vp]),
    "y3d_extract_patches": (_i, [_vp, _i64, _i64, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "y3d_scatter_candidates": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "y3d_rotate_iou_eval": (_i, [_vp, _i, _vp, _i, _i, _vp, _vp]),
    "y3d_decode3d": (_i, _LEVELS + [_i, _i, _vp, _vp]),
    "y3d_decode_preds3d": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _d, _vp, _vp, _vp]),
    "y3d_tal_assign3d": (_i, [_vp] * 9 + [_i, _i, _i, _i, _i, _f, _f, _f, _f, _i, _vp, _vp, _i] + [_vp] * 7 +
                         [_vp, _sz, _vp]),
}

_lib = None


class Y3DError(RuntimeError):
    pass


def lib():
    """Loads liby3d_b200.so once.  Raises (never falls back) when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Y3DError(
                f"{LIB_PATH} is missing: build it with `python __graft_entry__.py build` "
                "(nvcc, sm_100a).  yolov10-3d_b200 has no CPU or PyTorch fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the library does not export a declared symbol
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(rc: int):
    if rc != 0:
        raise Y3DError(f"liby3d_b200: {lib().y3d_strerror(rc).decode()} (rc={rc})")


def workspace_bytes(stage, B=0, A=0, nc=0, M=0, k=0, D=0) -> int:
    return int(lib().y3d_workspace_bytes(stage, B, A, nc, M, k, D))
