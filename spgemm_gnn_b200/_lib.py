"""ctypes binding of libmaxk_b200.so (the C ABI declared in include/maxk_b200.h).

There is no CPU fallback: if the shared object is missing or fails to load, every product
entry point raises.  Build it with `python -m spgemm_gnn_b200.build` (or
`__graft_entry__.build()`).
"""
from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("MAXK_LIB") or os.path.join(HERE, "libmaxk_b200.so")

MK_OK, MK_EINVAL, MK_EUNSUPPORTED, MK_ECUDA, MK_ENODEVICE = 0, -1, -2, -3, -4

# name -> (restype, argtypes); must list every symbol of include/maxk_b200.h
_i32, _i64, _vp = ctypes.c_int, ctypes.c_int64, ctypes.c_void_p
SIGNATURES = {
    "mk_version": (_i32, []),
    "mk_error_string": (ctypes.c_char_p, [_i32]),
    "mk_last_cuda_error": (ctypes.c_char_p, []),
    "mk_device_ok": (_i32, []),
    "mk_topk_cbsr": (_i32, [_vp, _i64, _i32, _i32, _vp, _vp, _i32, _vp]),
    "mk_topk_cbsr_bank": (_i32, [_vp, _i64, _i32, _i32, _vp, _vp, _i32, _vp, _vp, _vp, _vp]),
    "mk_cbsr_scatter": (_i32, [_vp, _vp, _i32, _vp, _i64, _i32, _i32, _vp]),
    "mk_cbsr_gather": (_i32, [_vp, _vp, _i32, _vp, _i64, _i32, _i32, _vp]),
    "mk_partition": (_i32, [_vp, _i64, _i32, _vp, ctypes.POINTER(_i64), ctypes.POINTER(_i64), _vp]),
    "mk_block_ptr": (_i32, [_vp, _vp, _i64, _i32, _i32, _vp, _vp]),
    "mk_partition_ranges": (_i32, [_vp, _vp, _i64, _i64, _i32, _i32, _vp, ctypes.POINTER(_i64),
                                   ctypes.POINTER(_i64), _vp]),
    "mk_spgemm_fwd": (_i32, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _i64, _i32, _i32, _vp]),
    "mk_sspmm_bwd": (_i32, [_vp, _i64, _vp, _vp, _vp, _vp, _i32, _vp, _i64, _i64, _i32, _i32, _vp]),
    "mk_sspmm_bwd_tiled": (_i32, [_vp, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _i64, _i64, _i32, _i32, _vp]),
    "mk_sspmm_bwd_tma": (_i32, [_vp, _i64, _vp, _vp, _vp, _vp, _i32, _vp, _i64, _i64, _i32, _i32, _i32, _vp]),
    "mk_layernorm_parts": (_i32, []),
    "mk_add_layernorm_fwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, ctypes.c_float, _vp]),
    "mk_layernorm_bwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp]),
    "mk_banked_supported": (_i32, [_i32, _i32]),
    "mk_banked_rows": (_i32, [_i32]),
    "mk_cbsr_bank": (_i32, [_vp, _vp, _i32, _vp, _vp, _vp, _i64, _i32, _i32, _vp]),
    "mk_spgemm_fwd_banked": (_i32, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp]),
    "mk_spgemm_fwd_banked_ex": (_i32, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32,
                                       _vp, _vp, _vp]),
    "mk_spgemm_fwd_banked_phase": (_i32, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32,
                                          _vp, _vp, _vp]),
    "mk_spgemm_fwd_banked_ln": (_i32, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32,
                                       _vp, _vp]),
    "mk_packed_supported": (_i32, [_i32, _i32]),
    "mk_cbsr_bank_packed": (_i32, [_vp, _vp, _i32, _vp, _i64, _i32, _i32, _vp]),
    "mk_spgemm_fwd_packed_ex": (_i32, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32,
                                       _vp, _vp, _vp]),
    "mk_sspmm_bwd_banked": (_i32, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i32, _i32, _vp]),
    "mk_peer_alloc": (_i32, [_i64, ctypes.POINTER(_vp)]),
    "mk_peer_free": (_i32, [_vp]),
    "mk_peer_export": (_i32, [_vp, ctypes.c_char_p]),
    "mk_peer_open": (_i32, [ctypes.c_char_p, ctypes.POINTER(_vp)]),
    "mk_peer_close": (_i32, [_vp]),
    "mk_peer_epoch": (_i32, [_vp, ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32), _vp]),
    "mk_peer_begin_push": (_i32, [_vp, _i32, _i32, _i32, _i32, _vp]),
    "mk_peer_publish": (_i32, [_vp, _i32, _i32, _vp]),
    "mk_peer_push": (_i32, [ctypes.POINTER(_vp), _i32, _i32, _i32, ctypes.POINTER(_i64), ctypes.POINTER(_i64), _vp]),
    "mk_peer_push_steps": (_i32, [ctypes.POINTER(_vp), _i32, _i32, _i32, ctypes.POINTER(_i64), ctypes.POINTER(_i64),
                                  _i32, _i32, _vp]),
    "mk_peer_push_sm": (_i32, [ctypes.POINTER(_vp), _i32, _i32, _i32, ctypes.POINTER(_i64), ctypes.POINTER(_i64),
                               _i32, _vp]),
    "mk_peer_wait_all": (_i32, [_vp, _i32, _i32, _vp]),
    "mk_peer_release": (_i32, [ctypes.POINTER(_vp), _i32, _i32, _vp]),
    "mk_peer_reduce_scatter": (_i32, [ctypes.POINTER(_vp), _i32, _i32, _i64, _i64, _vp, _i32, _i32, _vp]),
    "mk_peer_push_mc": (_i32, [ctypes.POINTER(_vp), _vp, _i32, _i32, _i32, ctypes.POINTER(_i64), ctypes.POINTER(_i64),
                               _i32, _vp]),
    "mk_peer_reduce_scatter_mc": (_i32, [ctypes.POINTER(_vp), _vp, _i32, _i32, _i64, _i64, _vp, _i32, _i32, _vp]),
    "mk_peer_reduce_scatter_virtual": (_i32, [ctypes.POINTER(_vp), _i32, _i64, _i64, ctypes.POINTER(_vp), _i32, _i32, _vp]),
}

class FwdExchange(ctypes.Structure):
    """`mk_fwd_exchange` of include/maxk_b200.h."""
    _fields_ = [("window", _vp), ("world", ctypes.c_int32), ("rank", ctypes.c_int32),
                ("rows_per_rank", _i64), ("timeout_ms", ctypes.c_int32), ("pushers", ctypes.c_int32),
                ("h_windows", ctypes.POINTER(_vp)), ("n_seg", ctypes.c_int32),
                ("h_offsets", ctypes.POINTER(_i64)), ("h_bytes", ctypes.POINTER(_i64))]


class FwdPhase(ctypes.Structure):
    """`mk_fwd_phase` of include/maxk_b200.h."""
    _fields_ = [("blk_ptr", _vp), ("row_stride", _i64), ("n_blocks", ctypes.c_int32),
                ("a0", ctypes.c_int32), ("a1", ctypes.c_int32), ("b0", ctypes.c_int32), ("b1", ctypes.c_int32),
                ("accumulate", ctypes.c_int32), ("last", ctypes.c_int32)]


class FwdEpilogue(ctypes.Structure):
    """`mk_fwd_epilogue` of include/maxk_b200.h."""
    _fields_ = [("h_self", _vp), ("bias", _vp), ("gamma", _vp), ("beta", _vp), ("z", _vp), ("mean", _vp),
                ("rstd", _vp), ("eps", ctypes.c_float)]


_lib = None


class MaxKLibraryError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise MaxKLibraryError(
                f"{SO_PATH} is missing: the CUDA library of the hot path is not built "
                "(run `python -m spgemm_gnn_b200.build`); there is no CPU fallback")
        L = ctypes.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError here == header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int, what: str) -> None:
    if rc == MK_OK:
        return
    L = lib()
    msg = L.mk_error_string(rc).decode()
    if rc == MK_ECUDA:
        msg += ": " + L.mk_last_cuda_error().decode()
    raise RuntimeError(f"{what} failed: {msg}")
