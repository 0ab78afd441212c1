"""ctypes binding of libtgnn.so (the C ABI declared in include/tgnn.h).

The shared object is built in-tree by ``__graft_entry__.build()`` (``tilingnn_b200/_C/libtgnn.so``).
There is deliberately no fallback: if the library is missing or a call fails, a
``RuntimeError`` is raised (the reference's contract is that a failed forward surfaces as a Python
exception, /root/reference/graph_networks/network_utils.py:10-19).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_C", "libtgnn.so")

TGNN_BN_TRAIN, TGNN_BN_EVAL = 0, 1
ABI_VERSION = 5


class tgnn_cfg(C.Structure):
    _fields_ = [("d_x", C.c_int32), ("d_e", C.c_int32), ("width", C.c_int32),
                ("depth", C.c_int32), ("bn_mode", C.c_int32), ("device", C.c_int32)]


class tgnn_info(C.Structure):
    _fields_ = [(n, C.c_int64) for n in
                ("n_own", "n_rows", "n_global", "e_adj", "e_col", "n_edge_types", "adj_slots",
                 "launches_per_forward", "workspace_bytes", "collectives_per_forward", "conv_kernel",
                 "tile_rows", "peer_exchange", "t_rows", "t_blocks", "gin_kernel", "gin_window_tiles", "gin_direct_tiles",
                 "range_fallback_layers")]


_vp, _i64, _i32 = C.c_void_p, C.c_int64, C.c_int32

# name -> (restype, argtypes); the list is also what tests check against include/tgnn.h
SIGNATURES = {
    "tgnn_abi_version": (C.c_int, []),
    "tgnn_create": (C.c_int, [C.POINTER(tgnn_cfg), C.POINTER(_vp)]),
    "tgnn_destroy": (C.c_int, [_vp]),
    "tgnn_set_param": (C.c_int, [_vp, C.c_char_p, _vp, C.POINTER(_i64), _i32]),
    "tgnn_missing_params": (C.c_int, [_vp, C.c_char_p, _i32]),
    "tgnn_set_bn_mode": (C.c_int, [_vp, _i32]),
    "tgnn_set_graph": (C.c_int, [_vp, _i64, _i64, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "tgnn_forward": (C.c_int, [_vp, _vp, _vp, _vp]),
    "tgnn_set_node_mask": (C.c_int, [_vp, _vp, C.POINTER(_i64), _vp]),
    "tgnn_check_error": (C.c_int, [_vp, _vp, _i32]),
    "tgnn_nccl_unique_id": (C.c_int, [_vp]),
    "tgnn_shard_init": (C.c_int, [_vp, _vp, _i32, _i32]),
    "tgnn_set_graph_shard": (C.c_int, [_vp, _i64, _i64, _i64, _i64, _vp, _i64, _vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    "tgnn_set_halo_peers": (C.c_int, [_vp, _vp, _i64, _vp]),
    "tgnn_get_info": (C.c_int, [_vp, C.POINTER(tgnn_info)]),
    "tgnn_debug_set_stop_layer": (C.c_int, [_vp, _i32]),
    "tgnn_debug_read": (C.c_int, [_vp, C.c_char_p, _vp, _vp]),
    "tgnn_debug_graph": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tgnn_debug_graph_t": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "tgnn_debug_role_cycles": (C.c_int, [_vp, C.POINTER(_i64)]),
    "tgnn_set_profiling": (C.c_int, [_vp, _i32]),
    "tgnn_get_profile": (C.c_int, [_vp, C.c_char_p, C.POINTER(C.c_float), C.POINTER(_i32)]),
    "tgnn_last_error": (C.c_char_p, [_vp]),
}

_lib = None


def load():
    """Load libtgnn.so (once) and declare every signature.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"tilingnn_b200: {LIB_PATH} is missing -- build it with "
            f"`python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    if lib.tgnn_abi_version() != ABI_VERSION:
        raise RuntimeError("tilingnn_b200: libtgnn.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def check(handle, rc, what):
    if rc != 0:
        msg = load().tgnn_last_error(handle)
        raise RuntimeError(f"{what}: {msg.decode() if msg else 'unknown error'}")
