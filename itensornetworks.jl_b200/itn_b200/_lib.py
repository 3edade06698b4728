"""ctypes binding of libitn_b200.so (the same C ABI a Julia `ccall` shim binds).

The library is the product: if it cannot be loaded, or no B200 is present, every call fails loudly.
There is no CPU fallback on this path.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libitn_b200.so")


class ITNError(RuntimeError):
    """Mirror of the reference's `error(...)` -> ErrorException convention (src/apply.jl:120-128)."""

    def __init__(self, code, msg):
        super().__init__(f"[itn_b200 status {code}] {msg}")
        self.code = code


_lib = None

_i32p = C.POINTER(C.c_int32)
_dp = C.POINTER(C.c_double)
_vp = C.c_void_p

_SIGS = {
    "itn_last_error": (C.c_char_p, []),
    "itn_version": (C.c_int, []),
    "itn_ctx_create": (C.c_int, [C.c_int, _vp, C.POINTER(_vp)]),
    "itn_ctx_destroy": (C.c_int, [_vp]),
    "itn_ctx_sync": (C.c_int, [_vp]),
    "itn_nccl_unique_id": (C.c_int, [_vp]),
    "itn_ctx_init_dist": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "itn_ctx_create_group": (C.c_int, [C.c_int, _i32p, C.POINTER(_vp)]),
    "itn_ctx_rank": (C.c_int, [_vp, _i32p, _i32p]),
    "itn_net_create": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _i32p, _i32p, _i32p, _i32p, _i32p, C.POINTER(_vp)]),
    "itn_net_clone": (C.c_int, [_vp, C.POINTER(_vp)]),
    "itn_net_destroy": (C.c_int, [_vp]),
    "itn_sync": (C.c_int, [_vp]),
    "itn_net_edge_dim": (C.c_int, [_vp, C.c_int, _i32p]),
    "itn_net_tensor_size": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_int64)]),
    "itn_net_set_tensor": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _i32p]),
    "itn_net_set_bra_tensor": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _i32p]),
    "itn_net_clear_bra": (C.c_int, [_vp]),
    "itn_net_set_tensors": (C.c_int, [_vp, C.c_int, _i32p, C.POINTER(_vp), _i32p, _i32p, C.c_int]),
    "itn_net_get_tensor": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _i32p]),
    "itn_msg_set_identity": (C.c_int, [_vp]),
    "itn_msg_set": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "itn_msg_get": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "itn_msg_get_all": (C.c_int, [_vp, _vp, C.c_int64]),
    "itn_bp_update": (C.c_int, [_vp, _i32p, _i32p, C.c_int, _i32p, C.c_int, C.c_int, C.c_double, C.c_int, _i32p, _dp]),
    "itn_updated_message": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _vp]),
    "itn_message_residuals": (C.c_int, [_vp, _i32p, _i32p, C.c_int, _dp]),
    "itn_region_scalars": (C.c_int, [_vp, _vp, _vp]),
    "itn_logscalar": (C.c_int, [_vp, _dp]),
    "itn_rescale": (C.c_int, [_vp]),
    "itn_rescale_verts": (C.c_int, [_vp, _i32p, C.c_int]),
    "itn_expect1": (C.c_int, [_vp, _i32p, C.c_int, _vp, _vp]),
    "itn_rdm2": (C.c_int, [_vp, _i32p, C.c_int, _vp]),
    "itn_apply1": (C.c_int, [_vp, _i32p, C.c_int, _vp, C.c_int]),
    "itn_apply2": (C.c_int, [_vp, _i32p, C.c_int, _vp, C.c_int, C.c_double, C.c_int, C.c_int, _i32p, _dp, _dp, C.c_int]),
    "itn_apply_layers": (C.c_int, [_vp, C.c_int, _i32p, _i32p, _vp, C.c_int, C.c_double, C.c_int, C.c_int, _i32p, _i32p, C.c_int,
                                   _i32p, C.c_int, C.c_int, C.c_double, C.c_int, _i32p, _dp, _dp, C.c_int, _i32p]),
    "itn_gauge_walk": (C.c_int, [_vp, _i32p, _i32p, C.c_int]),
    "itn_map_eigvals": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, C.c_double]),
    "itn_tensordot": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _i32p, _vp, C.c_int, _i32p, C.c_int, _i32p, _i32p, _vp]),
    "itn_svd_batch": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _dp, _vp, C.c_int, _dp]),
    "itn_block_plan_export": (C.c_int, [C.c_int, C.c_int, C.c_int, _i32p, C.c_int, _i32p, C.c_int, _i32p]),
    "itn_ctx_launch_count": (C.c_int, [_vp, C.POINTER(C.c_int64)]),
    "itn_ctx_path_counts": (C.c_int, [_vp, C.POINTER(C.c_int64)]),
    "itn_ctx_cholqr2_count": (C.c_int, [_vp, C.POINTER(C.c_int64)]),
    "itn_ctx_set_path": (C.c_int, [_vp, C.c_int]),
    "itn_bp_last_timing": (C.c_int, [_vp, _dp, _dp]),
}

EXPORTED_SYMBOLS = tuple(_SIGS)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ITNError(-1, f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            f = getattr(l, name)
            f.restype = res
            f.argtypes = args
        _lib = l
    return _lib


def check(status):
    if status != 0:
        raise ITNError(status, lib().itn_last_error().decode("utf-8", "replace"))


def i32(a):
    import numpy as np
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(_i32p)
