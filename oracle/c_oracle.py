"""ctypes front-end of oracle/maxk_oracle.c (test infrastructure, NOT product code)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmaxk_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "maxk_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        i64, i32, vp = ctypes.c_int64, ctypes.c_int, ctypes.c_void_p
        L.mko_num_threads.restype = i32
        L.mko_topk_cbsr.argtypes = [vp, i64, i32, i32, vp, vp, i32]
        L.mko_cbsr_scatter.argtypes = [vp, vp, i32, vp, i64, i32, i32]
        L.mko_cbsr_gather.argtypes = [vp, vp, i32, vp, i64, i32, i32]
        L.mko_partition.argtypes = [vp, i64, i32, vp, vp]
        L.mko_partition.restype = i64
        L.mko_spgemm_fwd.argtypes = [vp, vp, vp, vp, vp, i32, vp, i64, i32, i32]
        L.mko_sspmm_bwd.argtypes = [vp, vp, vp, vp, vp, i32, vp, i64, i64, i32, i32]
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def num_threads() -> int:
    return lib().mko_num_threads()


def maxk_cbsr(x, k):
    x = _c(x, np.float32)
    n, d = x.shape
    idt = np.uint8 if d <= 256 else np.uint16
    sp_data = np.empty((n, k), np.float32)
    sp_index = np.empty((n, k), idt)
    rc = lib().mko_topk_cbsr(_p(x), n, d, k, _p(sp_data), _p(sp_index), sp_index.itemsize)
    if rc != 0:
        raise ValueError(f"mko_topk_cbsr failed ({rc})")
    return sp_data, sp_index


def cbsr_scatter(g, sp_index, d):
    g = _c(g, np.float32)
    sp_index = np.ascontiguousarray(sp_index)
    n, k = g.shape
    out = np.empty((n, d), np.float32)
    lib().mko_cbsr_scatter(_p(g), _p(sp_index), sp_index.itemsize, _p(out), n, k, d)
    return out


def cbsr_gather(dense, sp_index):
    dense = _c(dense, np.float32)
    sp_index = np.ascontiguousarray(sp_index)
    n, d = dense.shape
    k = sp_index.shape[1]
    out = np.empty((n, k), np.float32)
    lib().mko_cbsr_gather(_p(dense), _p(sp_index), sp_index.itemsize, _p(out), n, k, d)
    return out


def partition_rows(indptr, max_nz):
    indptr = _c(indptr, np.int32)
    n = indptr.size - 1
    slots = ctypes.c_int64(0)
    p = lib().mko_partition(_p(indptr), n, max_nz, None, ctypes.byref(slots))
    recs = np.empty((p, 4), np.int32)
    lib().mko_partition(_p(indptr), n, max_nz, _p(recs), ctypes.byref(slots))
    return recs, int(slots.value)


def spgemm_fwd(indptr, indices, val, sp_data, sp_index, d):
    indptr, indices = _c(indptr, np.int32), _c(indices, np.int32)
    val, sp_data = _c(val, np.float32), _c(sp_data, np.float32)
    sp_index = np.ascontiguousarray(sp_index)
    n = indptr.size - 1
    k = sp_data.shape[1]
    out = np.empty((n, d), np.float64)
    lib().mko_spgemm_fwd(_p(indptr), _p(indices), _p(val), _p(sp_data), _p(sp_index),
                         sp_index.itemsize, _p(out), n, k, d)
    return out


def sspmm_bwd(indptr, indices, val, dy, sp_index):
    indptr, indices = _c(indptr, np.int32), _c(indices, np.int32)
    val, dy = _c(val, np.float32), _c(dy, np.float32)
    sp_index = np.ascontiguousarray(sp_index)
    n = indptr.size - 1
    ns, k = sp_index.shape
    out = np.empty((ns, k), np.float64)
    rc = lib().mko_sspmm_bwd(_p(indptr), _p(indices), _p(val), _p(dy), _p(sp_index),
                             sp_index.itemsize, _p(out), n, ns, k, dy.shape[1])
    if rc != 0:
        raise MemoryError("mko_sspmm_bwd")
    return out
