"""ctypes binding of the C ABI in include/ultra_rspmm.h (libultra_rspmm.so, built by ultra_b200.build).

There is deliberately no fallback: if the library is missing, loading raises with the build command.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ULTRA_RSPMM_LIB") or os.path.join(_HERE, "lib", "libultra_rspmm.so")  # env: dev A/B builds

SUM_IDS = {"add": 0, "min": 1, "max": 2}
MUL_IDS = {"mul": 0, "add": 1}
F32, F64 = 0, 1

# every symbol include/ultra_rspmm.h declares: name -> (restype, argtypes)
_vp, _i64, _int, _sz = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_size_t
SYMBOLS = {
    "ultra_rspmm_version": (ctypes.c_char_p, []),
    "ultra_rspmm_last_error": (ctypes.c_char_p, []),
    "ultra_rspmm_set_tunable": (_int, [ctypes.c_char_p, _int]),
    "ultra_rspmm_plan_set_sm_limit": (_int, [_vp, _int]),
    "ultra_rspmm_debug_cut_items": (_int, [_vp, _int, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _vp]),
    "ultra_rspmm_plan_debug_read": (_int, [_vp, _int, _int, _vp, _i64]),
    "ultra_rspmm_plan_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "ultra_rspmm_plan_build": (_int, [ctypes.POINTER(_vp), _vp, _vp, _i64, _i64, _i64, _vp, _sz, _vp]),
    "ultra_rspmm_plan_apply_mask": (_int, [_vp, _vp, _vp]),
    "ultra_rspmm_plan_destroy": (None, [_vp]),
    "ultra_rspmm_plan_input_sorted": (_int, [_vp]),
    "ultra_rspmm_plan_info": (_i64, [_vp, _int]),
    "ultra_rspmm_scratch_bytes": (_sz, [_vp, _i64, _int]),
    "ultra_rspmm_set_edge_weight": (_int, [_vp, _vp, _int, _vp]),
    "ultra_rspmm_forward": (_int, [_vp, _int, _int, _int, _vp, _vp, _vp, _i64, _vp, _sz, _vp]),
    "ultra_rspmm_backward": (_int, [_vp, _int, _int, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _sz, _vp]),
    "ultra_rspmm_forward_pna": (_int, [_vp, _int, _int, _vp, _vp, _vp, _i64, _vp, _sz, _vp]),
    "ultra_rspmm_forward_heads": (_int, [_vp, _int, _int, _int, _vp, _vp, _vp, _vp, _i64, _i64, _int, _vp, _sz, _vp]),
    "ultra_rspmm_backward_heads": (_int, [_vp, _int, _int, _int, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _int, _vp,
                                          _sz, _vp]),
    "ultra_relation_project_f32": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp]),
    "ultra_score_head_f32": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _vp]),
    "ultra_nbfnet_update_f32": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, ctypes.c_float, _int, _int,
                                       _int, _int, _vp]),
    "ultra_nbfnet_update_train_f32": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, ctypes.c_float,
                                             _int, _int, _int, _vp]),
    "ultra_nbfnet_update_bwd_grid": (_int, [_i64]),
    "ultra_nbfnet_update_bwd_f32": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _int, _vp]),
    "ultra_rspmm_host_staging_bytes": (_sz, [_vp, _i64, _i64, _int]),
    "ultra_rspmm_forward_host": (_int, [_vp, _int, _int, _int, _vp, _vp, _vp, _i64, _i64, _vp, _sz, _vp]),
}

_lib = None


class RspmmLibraryError(RuntimeError):
    pass


def lib():
    """Load libultra_rspmm.so once; raise loudly if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RspmmLibraryError(
                "ultra_b200: %s not found - build it with `python -m ultra_b200.build` "
                "(there is no CPU/PyTorch fallback for rspmm)" % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)  # AttributeError if the .so is stale
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def set_tunable(name, value):
    check(lib().ultra_rspmm_set_tunable(name.encode(), int(value)), "ultra_rspmm_set_tunable")


def check(rc, what):
    if rc != 0:
        msg = lib().ultra_rspmm_last_error().decode()
        raise RuntimeError("%s failed (code %d): %s" % (what, rc, msg))
