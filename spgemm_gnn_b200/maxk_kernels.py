"""`maxk_kernels` -- the reference's PyTorch-extension entry points, kept name for name and
argument for argument (positional), as a thin shim over the C ABI of libmaxk_b200.so.

Reference surface (pybind module `maxk_kernels`, kernels/maxk_bindings.cpp -- source absent,
signatures and TORCH_CHECK strings recovered from the shipped binary):

    maxk_forward(input, k) -> Tensor
    maxk_backward(grad_output, indices) -> Tensor
    spgemm_forward(ptr, idx, val, sp_data, sp_index, num_nodes, num_edges, dim_sparse, dim_origin)
        -> (Tensor, Tensor)
    spgemm_backward(ptr, idx, val, grad_output, sp_index, num_nodes, num_edges, dim_sparse, dim_origin)
        -> Tensor

Additions (new names, existing ones untouched): `maxk_forward_cbsr`, `cbsr_scatter`,
`cbsr_gather`, `partition`, `clear_partition_cache`, `set_max_nz`.

Everything launches on torch's current CUDA stream; nothing synchronises the device except
the one-off partition size query per graph.  No CPU path exists.
"""
from __future__ import annotations

import ctypes
import os
import weakref
from typing import Optional, Tuple

import torch

from . import _lib

__all__ = [
    "maxk_forward", "maxk_backward", "spgemm_forward", "spgemm_backward",
    "maxk_forward_cbsr", "maxk_forward_cbsr_banked", "cbsr_scatter", "cbsr_gather", "partition", "forward_phases",
    "clear_partition_cache", "install_partition", "set_max_nz", "get_max_nz", "launch_count",
    "banked_supported", "cbsr_bank", "block_split", "packed_supported", "cbsr_bank_packed",
    "spgemm_forward_packed", "use_packed", "ForwardExchange", "set_backward_tiled", "backward_tiles",
    "block_pointers", "maxk_forward_banked", "spgemm_forward_banked", "spgemm_forward_ln",
    "spgemm_backward_banked", "set_banked", "set_backward_tma", "use_banked", "forward_variant", "partition_blocked", "backward_blocks",
    "set_backward_block_mb", "add_layernorm_supported", "add_layernorm_forward", "layernorm_backward",
]

_MAX_NZ = int(os.environ.get("MAXK_MAX_NZ", "1024"))
# The backward has no partial rows to fold, so shorter records cost nothing and shorten the tail of
# the grid (measured on an 8-way shard of the Reddit shape: 0.370 ms at 1024, 0.346 ms at 256).
_BWD_MAX_NZ = int(os.environ.get("MAXK_BWD_MAX_NZ", "0"))      # 0: same as the forward
# Records handed to the CTAs longest first (LPT): the grid then drains on the shortest records instead
# of on a straggling 1024-entry one.  The row-ordered list stays what the fold of multi-record rows uses.
_EXEC_SORTED = os.environ.get("MAXK_EXEC_ORDER", "1") != "0"
_BANKED = os.environ.get("MAXK_BANKED", "1") != "0"
# mean stored entries per work record below which banking does not pay: every record zeroes and folds
# 8 KB of cells, which 50 entries do not amortise (products shape, k = 32: 5.72 ms plain, 8.70 ms banked;
# Flickr shape 0.079 / 0.138 ms -- profiles/r2/banked_short_records_call26.log)
_BANKED_MIN_RECORD = int(os.environ.get("MAXK_BANKED_MIN_RECORD", "96"))
_PACKED = os.environ.get("MAXK_PACKED", "1") != "0"   # k = 8, 16: banked + packed 8-byte entries
# experimental backward: this many of the 4 neighbours of a warp step (k = 32) reduce through the TMA
# unit (csrc/sspmm_bwd.cu, mk_sspmm_bwd_tma; measured slower); 0 = the shipped kernel
_BWD_TMA = int(os.environ.get("MAXK_BWD_TMA", "0"))
_launches = 0  # kernels launched through this module (bench.py reports it)


def launch_count() -> int:
    return _launches


def set_max_nz(max_nz: int) -> None:
    """Stored entries per work record (the reference hard-wires WARP_MAX_NZ=64,
    README_INTEGRATED.md:256)."""
    global _MAX_NZ
    if max_nz < 1:
        raise ValueError("max_nz must be positive")
    _MAX_NZ = int(max_nz)


def get_max_nz() -> int:
    return _MAX_NZ


def set_backward_tma(neighbours: int) -> None:
    """Experimental: route `neighbours` (1, 2 or 4) of the 128/k neighbours a warp handles per step
    of the backward through bulk shared->global reductions (TMA) instead of REDG.  0 switches back."""
    global _BWD_TMA
    _BWD_TMA = int(neighbours)


def set_banked(on: bool) -> None:
    """Let `spgemm_forward` re-order the CBSR table into the conflict-free banked form first
    (csrc/bank.cu).  Same result; on by default where it is faster."""
    global _BANKED
    _BANKED = bool(on)


def use_banked(num_parts: int, num_edges: int, dim_sparse: int, dim_origin: int) -> bool:
    # measured on the Reddit shape: banked wins at k = 32 (3.94 -> 3.0 ms incl. banking) and 64,
    # ties at 16, loses at 8 (too few entries per row to balance 8 banks)
    return (_BANKED and dim_sparse >= 32 and banked_supported(dim_sparse, dim_origin)
            and num_edges >= _BANKED_MIN_RECORD * max(num_parts, 1))


def use_packed(num_parts: int, num_edges: int, dim_sparse: int, dim_origin: int) -> bool:
    """k = 8, 16 on long records: the banked forward on the packed table (8-byte entries)."""
    return (_BANKED and _PACKED and dim_sparse in (8, 16) and packed_supported(dim_sparse, dim_origin)
            and num_edges >= _BANKED_MIN_RECORD * max(num_parts, 1))


def forward_variant(num_parts: int, num_edges: int, dim_sparse: int, dim_origin: int) -> str:
    """Which forward `spgemm_forward` runs for this shape (reported by bench.py)."""
    if use_banked(num_parts, num_edges, dim_sparse, dim_origin):
        return "banked (mk_cbsr_bank + mk_spgemm_fwd_banked, both inside the forward time)"
    if use_packed(num_parts, num_edges, dim_sparse, dim_origin):
        return "packed banked (mk_cbsr_bank_packed + mk_spgemm_fwd_packed, both inside the forward time)"
    return "plain (mk_spgemm_fwd)"


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _chk(cond: bool, msg: str) -> None:
    # TORCH_CHECK -> RuntimeError, messages as in the reference binary
    if not cond:
        raise RuntimeError(msg)


def _cuda_contig(t: torch.Tensor, name: str) -> None:
    _chk(isinstance(t, torch.Tensor) and t.is_cuda, f"{name} must be a CUDA tensor")
    _chk(t.is_contiguous(), f"{name} must be contiguous")


def _index_bytes(sp_index: torch.Tensor, dim_origin: int) -> int:
    if sp_index.dtype == torch.uint8:
        _chk(dim_origin <= 256, "sp_index must be uint16 when dim_origin > 256")
        return 1
    if sp_index.dtype in (torch.uint16, torch.int16):
        return 2
    raise RuntimeError("sp_index must be uint8 or uint16")


def _index_dtype(dim_origin: int) -> torch.dtype:
    return torch.uint8 if dim_origin <= 256 else torch.uint16


# ---------------------------------------------------------------------------------------
# MaxK
# ---------------------------------------------------------------------------------------
def maxk_forward_cbsr(input: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Exact top-k per row as CBSR: (sp_data fp32 [N,k], sp_index uint8|uint16 [N,k]),
    ascending columns, ties to the lower column.  The native `maxk_forward` of the
    reference computes both and throws the index away (maxk_cuda_kernels.o@0x31d-0x3a5)."""
    global _launches
    _cuda_contig(input, "input")
    _chk(input.dim() == 2, "Input must be 2D tensor")
    _chk(input.dtype == torch.float32, "input must be float32")
    n, d = input.shape
    _chk(1 <= k <= d, "k must be between 1 and input dimension")
    _chk(d <= 65536, "input dimension above 65536 is not supported")
    sp_data = torch.empty((n, k), dtype=torch.float32, device=input.device)
    sp_index = torch.empty((n, k), dtype=_index_dtype(d), device=input.device)
    with torch.cuda.device(input.device):
        rc = _lib.lib().mk_topk_cbsr(input.data_ptr(), n, d, k, sp_data.data_ptr(),
                                     sp_index.data_ptr(), sp_index.element_size(), _stream())
    _lib.check(rc, "mk_topk_cbsr")
    _launches += 1
    return sp_data, sp_index


def maxk_forward_cbsr_banked(input: torch.Tensor, k: int, *, want_data: bool = False, packed: bool = False,
                             out=None):
    """Top-k and banking in ONE kernel (mk_topk_cbsr_bank, f-3): the dense row is read once and leaves
    as the sorted column ids (for the backward) plus the banked table the forward SpGEMM reads.
    Returns `(sp_data | None, sp_index, bk_data, bk_slot)`, or with `packed` (k = 8, 16)
    `(sp_data | None, sp_index, bk_pack, None)`.  Bit-identical to `maxk_forward_cbsr` + `cbsr_bank` /
    `cbsr_bank_packed`.  `out = (bk_data, bk_slot)` or `(bk_pack,)`: write the banked rows there (a
    rank's rows of a peer window)."""
    global _launches
    _cuda_contig(input, "input")
    _chk(input.dim() == 2, "Input must be 2D tensor")
    _chk(input.dtype == torch.float32, "input must be float32")
    n, d = input.shape
    _chk(1 <= k <= d, "k must be between 1 and input dimension")
    _chk(banked_supported(k, d), "banked CBSR needs k in {8,16,32,64}, dim % 8 == 0, dim <= 512")
    _chk(not packed or packed_supported(k, d), "packed CBSR needs k in {8,16}")
    dev = input.device
    sp_data = torch.empty((n, k), dtype=torch.float32, device=dev) if want_data else None
    sp_index = torch.empty((n, k), dtype=_index_dtype(d), device=dev)
    bk_data = bk_slot = bk_pack = None
    if packed:
        bk_pack = out[0] if out is not None else torch.empty((n, k, 2), dtype=torch.int32, device=dev)
        _cuda_contig(bk_pack, "out")
        _chk(bk_pack.dtype == torch.int32 and tuple(bk_pack.shape) == (n, k, 2), "out must be int32 [n, k, 2]")
    else:
        if out is not None:
            bk_data, bk_slot = out
        else:
            bk_data = torch.empty((n, k), dtype=torch.float32, device=dev)
            bk_slot = torch.empty((n, k), dtype=torch.int16, device=dev)
        _cuda_contig(bk_data, "out")
        _cuda_contig(bk_slot, "out")
        _chk(bk_data.dtype == torch.float32 and bk_slot.dtype in (torch.int16, torch.uint16)
             and tuple(bk_data.shape) == (n, k) and tuple(bk_slot.shape) == (n, k),
             "out must be (float32 [n,k], int16 [n,k])")
    with torch.cuda.device(dev):
        rc = _lib.lib().mk_topk_cbsr_bank(
            input.data_ptr(), n, d, k, sp_data.data_ptr() if want_data else None, sp_index.data_ptr(),
            sp_index.element_size(), None if packed else bk_data.data_ptr(),
            None if packed else bk_slot.data_ptr(), bk_pack.data_ptr() if packed else None, _stream())
    _lib.check(rc, "mk_topk_cbsr_bank")
    _launches += 1
    return sp_data, sp_index, (bk_pack if packed else bk_data), bk_slot


def maxk_forward(input: torch.Tensor, k: int) -> torch.Tensor:
    """Reference signature `(Tensor input, int k) -> Tensor`: the [N,k] CBSR values
    (SURVEY.md section 2.2).  Use `maxk_forward_cbsr` to get the column ids as well."""
    return maxk_forward_cbsr(input, k)[0]


def cbsr_scatter(grad: torch.Tensor, sp_index: torch.Tensor, dim_origin: int) -> torch.Tensor:
    """dense [N, dim_origin]: zeros with dense[i, sp_index[i,t]] = grad[i,t]."""
    global _launches
    _cuda_contig(grad, "grad_output")
    _cuda_contig(sp_index, "indices")
    _chk(grad.dim() == 2, "grad_output must be 2D tensor")
    _chk(grad.dtype == torch.float32, "grad_output must be float32")
    _chk(sp_index.shape == grad.shape, "indices must have the shape of grad_output")
    n, k = grad.shape
    _chk(1 <= k <= dim_origin, "k must be between 1 and input dimension")
    ib = _index_bytes(sp_index, dim_origin)
    out = torch.empty((n, dim_origin), dtype=torch.float32, device=grad.device)
    with torch.cuda.device(grad.device):
        rc = _lib.lib().mk_cbsr_scatter(grad.data_ptr(), sp_index.data_ptr(), ib, out.data_ptr(),
                                        n, k, dim_origin, _stream())
    _lib.check(rc, "mk_cbsr_scatter")
    _launches += 1
    return out


def cbsr_gather(dense: torch.Tensor, sp_index: torch.Tensor) -> torch.Tensor:
    """[N,k]: dense[i, sp_index[i,t]]."""
    global _launches
    _cuda_contig(dense, "input")
    _cuda_contig(sp_index, "indices")
    _chk(dense.dim() == 2 and sp_index.dim() == 2, "Input must be 2D tensor")
    _chk(dense.dtype == torch.float32, "input must be float32")
    n, d = dense.shape
    k = sp_index.shape[1]
    _chk(sp_index.shape[0] == n and 1 <= k <= d, "k must be between 1 and input dimension")
    ib = _index_bytes(sp_index, d)
    out = torch.empty((n, k), dtype=torch.float32, device=dense.device)
    with torch.cuda.device(dense.device):
        rc = _lib.lib().mk_cbsr_gather(dense.data_ptr(), sp_index.data_ptr(), ib, out.data_ptr(),
                                       n, k, d, _stream())
    _lib.check(rc, "mk_cbsr_gather")
    _launches += 1
    return out


def maxk_backward(grad_output: torch.Tensor, indices: torch.Tensor,
                  dim_origin: Optional[int] = None) -> torch.Tensor:
    """Reference signature `(Tensor grad_output, Tensor indices) -> Tensor`: scatter a
    CBSR-shaped gradient [N,k] to dense [N,D] (maxk_backward_cuda, maxk_cuda_kernels.o@0x4d0).
    Like the reference, D defaults to `indices.max()+1` (a device sync); pass `dim_origin`
    to avoid it.  int64 `indices` (what utils/maxk_layers.py:23 saves from torch.topk) are
    accepted and narrowed."""
    _cuda_contig(grad_output, "grad_output")
    _cuda_contig(indices, "indices")
    _chk(grad_output.dim() == 2, "grad_output must be 2D tensor")
    if dim_origin is None:
        dim_origin = int(indices.max().item()) + 1 if indices.numel() else 1
        dim_origin = max(dim_origin, grad_output.shape[1])
    if indices.dtype not in (torch.uint8, torch.uint16, torch.int16):
        if dim_origin <= 256:
            indices = indices.to(torch.uint8)
        else:  # narrow through int16: same bits as uint16, and every torch build can cast to it
            indices = indices.to(torch.int16).view(torch.uint16)
    return cbsr_scatter(grad_output, indices, dim_origin)


# ---------------------------------------------------------------------------------------
# work partition cache (per graph)
# ---------------------------------------------------------------------------------------
class _Partition:
    __slots__ = ("parts", "num_parts", "num_slots", "max_nz", "partial", "ptr_ref", "version", "_exec")

    def exec_parts(self) -> torch.Tensor:
        """The records in the order the CTAs take them.  Long records (mean >= 96 stored entries):
        longest first (stable), so that the grid drains on the shortest records instead of on a
        straggling max_nz one -- Reddit shape: forward 2.89 -> 2.83 ms, on an 8-way shard 0.427 ->
        0.385 ms (profiles/r2/exec_order_call2.log).  Short records (products shape, mean 51): only the
        heavy tail (> 8x the mean) moves to the front, the rest stays in row order, which is what keeps
        the edge arrays and the dense rows streaming.  MAXK_EXEC_ORDER=0: row order."""
        ex = getattr(self, "_exec", None)
        if ex is None:
            ex = self.parts
            if _EXEC_SORTED and self.num_parts > 1:
                lens = self.parts[: self.num_parts, 2]
                mean = float(lens.sum().item()) / self.num_parts
                if mean >= 96:
                    order = torch.argsort(lens, descending=True, stable=True)
                    ex = self.parts[order].contiguous()
                else:
                    heavy = lens > max(8.0 * mean, 128.0)
                    hi = heavy.nonzero().squeeze(1)
                    if hi.numel() > 0:
                        hi = hi[torch.argsort(lens[hi], descending=True, stable=True)]
                        order = torch.cat([hi, (~heavy).nonzero().squeeze(1)])
                        ex = self.parts[order].contiguous()
            self._exec = ex
        return ex

    def partial_for(self, d: int, device) -> Optional[torch.Tensor]:
        """Scratch rows for the records of multi-record rows.  A fresh tensor per call (the caching
        allocator makes that free) so that calls on different streams never share scratch."""
        if self.num_slots == 0:
            return None
        return torch.empty((self.num_slots, d), dtype=torch.float32, device=device)


_part_cache = {}


def _evict(cache: dict, limit: int) -> None:
    """Keep a record cache bounded: entries whose row pointer is gone go first, then the oldest
    (dicts keep insertion order) -- never the whole cache at once."""
    if len(cache) < limit:
        return
    for key in [k for k, v in cache.items() if v.ptr_ref() is None]:
        del cache[key]
    while len(cache) >= limit:
        del cache[next(iter(cache))]


def clear_partition_cache() -> None:
    _part_cache.clear()
    _block_cache.clear()
    _split_cache.clear()
    _blkptr_cache.clear()


def partition(ptr: torch.Tensor, num_nodes: int, max_nz: Optional[int] = None) -> _Partition:
    """Work records of a CSR row pointer, built on the GPU once per graph and cached on
    `(ptr.data_ptr(), num_nodes, max_nz)` -- replaces generate_meta.py + the `.warp4` file."""
    global _launches
    max_nz = _MAX_NZ if max_nz is None else int(max_nz)
    key = (ptr.device.index, ptr.data_ptr(), int(num_nodes), max_nz)
    hit = _part_cache.get(key)
    if hit is not None and hit.ptr_ref() is ptr and hit.version == ptr._version:
        return hit
    _chk(ptr.numel() >= num_nodes + 1, "ptr must have num_nodes + 1 entries")
    L = _lib.lib()
    np_, ns_ = ctypes.c_int64(0), ctypes.c_int64(0)
    with torch.cuda.device(ptr.device):
        rc = L.mk_partition(ptr.data_ptr(), num_nodes, max_nz, None, ctypes.byref(np_),
                            ctypes.byref(ns_), _stream())
        _lib.check(rc, "mk_partition")
        p = _Partition()
        p.num_parts, p.num_slots, p.max_nz = int(np_.value), int(ns_.value), max_nz
        p.parts = torch.empty((max(p.num_parts, 1), 4), dtype=torch.int32, device=ptr.device)
        p.partial = None
        p._exec = None
        rc = L.mk_partition(ptr.data_ptr(), num_nodes, max_nz, p.parts.data_ptr(), None, None,
                            _stream())
        _lib.check(rc, "mk_partition")
    _launches += 5
    p.ptr_ref = weakref.ref(ptr)
    p.version = ptr._version
    _evict(_part_cache, 64)
    _part_cache[key] = p
    return p


_BWD_BLOCK_BYTES = int(os.environ.get("MAXK_BWD_BLOCK_MB", "80")) << 20   # CBSR-gradient bytes per column block
# Off by default: measured on the products shape (2.45 M nodes, mean degree 51) it does not pay --
# 6.83 ms plain, 6.58 / 6.62 / 7.48 ms with 2 / 3 / 4 blocks -- because every (row, block) record
# re-stages dY[r] and the records get too short.  Kept for high-degree graphs with huge node counts.
_BWD_BLOCKED = os.environ.get("MAXK_BWD_BLOCKED", "0") != "0"
_block_cache = {}


def set_backward_block_mb(mb: int) -> None:
    """Target size of one destination block of the column-blocked backward (0 disables it)."""
    global _BWD_BLOCK_BYTES, _BWD_BLOCKED
    _BWD_BLOCKED = mb > 0
    _BWD_BLOCK_BYTES = max(int(mb), 1) << 20
    _block_cache.clear()


def backward_blocks(n_src: int, dim_sparse: int, num_nodes: int, num_edges: int) -> int:
    """How many destination blocks the backward should walk: 1 while the CBSR gradient
    (n_src*k*4 B) fits in L2 next to the streams, else enough blocks of ~80 MB -- but never so
    many that a (row, block) record drops below ~64 stored entries (each record re-stages dY[r])."""
    if not _BWD_BLOCKED:
        return 1
    need = -(-(n_src * dim_sparse * 4) // _BWD_BLOCK_BYTES)
    if need <= 1:
        return 1
    cap = max(int(num_edges // max(num_nodes, 1)) // 64, 1)
    return max(min(need, cap, 16), 1)


def partition_blocked(ptr: torch.Tensor, idx: torch.Tensor, num_nodes: int, n_src: int,
                      n_blocks: int, max_nz: Optional[int] = None) -> _Partition:
    """Work records ordered by destination block (column range) first, row second: one backward
    launch then walks the blocks in order and its vector reductions stay L2-resident.  Needs
    ascending column ids inside every row (checked once; falls back to the row-major records)."""
    global _launches
    max_nz = _MAX_NZ if max_nz is None else int(max_nz)
    key = (ptr.device.index, ptr.data_ptr(), idx.data_ptr(), int(num_nodes), int(n_src), int(n_blocks), max_nz)
    hit = _block_cache.get(key)
    if hit is not None and hit.ptr_ref() is ptr and hit.version == ptr._version:
        return hit
    e = int(idx.numel())
    if e > 1:  # ascending inside rows?
        bad = idx[1:] < idx[:-1]
        starts = ptr[1:num_nodes].to(torch.int64)
        starts = starts[(starts > 0) & (starts < e)]
        bad[starts - 1] = False
        if bool(bad.any()):
            return partition(ptr, num_nodes, max_nz)
    width = -(-n_src // n_blocks)
    L = _lib.lib()
    blk = torch.empty(((n_blocks + 1) * num_nodes,), dtype=torch.int32, device=ptr.device)
    np_, ns_ = ctypes.c_int64(0), ctypes.c_int64(0)
    with torch.cuda.device(ptr.device):
        _lib.check(L.mk_block_ptr(ptr.data_ptr(), idx.data_ptr(), num_nodes, n_blocks, width,
                                  blk.data_ptr(), _stream()), "mk_block_ptr")
        rs, re = blk.data_ptr(), blk.data_ptr() + 4 * num_nodes
        vrows = n_blocks * num_nodes
        _lib.check(L.mk_partition_ranges(rs, re, vrows, num_nodes, max_nz, 1, None, ctypes.byref(np_),
                                         ctypes.byref(ns_), _stream()), "mk_partition_ranges")
        p = _Partition()
        p.num_parts, p.num_slots, p.max_nz = int(np_.value), int(ns_.value), max_nz
        p.parts = torch.empty((max(p.num_parts, 1), 4), dtype=torch.int32, device=ptr.device)
        p.partial = None
        p._exec = p.parts      # block-major order IS the point of this list: never re-sorted
        _lib.check(L.mk_partition_ranges(rs, re, vrows, num_nodes, max_nz, 1, p.parts.data_ptr(), None,
                                         None, _stream()), "mk_partition_ranges")
    _launches += 6
    p.ptr_ref = weakref.ref(ptr)
    p.version = ptr._version
    _evict(_block_cache, 32)
    _block_cache[key] = p
    return p


def install_partition(ptr: torch.Tensor, num_nodes: int, max_nz: int, parts: torch.Tensor,
                      num_slots: int) -> _Partition:
    """Put work records that were built earlier (graph.load_graph) into the cache."""
    _chk(parts.is_cuda and parts.dtype == torch.int32 and parts.dim() == 2 and parts.shape[1] == 4,
         "parts must be a CUDA int32 [P,4] tensor")
    p = _Partition()
    p.num_parts, p.num_slots, p.max_nz = int(parts.shape[0]), int(num_slots), int(max_nz)
    p.parts = parts.contiguous()
    p.partial = None
    p._exec = None
    p.ptr_ref = weakref.ref(ptr)
    p.version = ptr._version
    _part_cache[(ptr.device.index, ptr.data_ptr(), int(num_nodes), int(max_nz))] = p
    return p


# ---------------------------------------------------------------------------------------
# SpGEMM forward / SSpMM backward
# ---------------------------------------------------------------------------------------
class ForwardExchange:
    """What a row-partitioned forward needs to run while its table is still arriving (built by
    peer.py, handed to `spgemm_forward_banked` / `spgemm_forward_packed` as `wait=`): the rank's own
    window, the rank layout and -- when the kernel itself is the all-gather -- every rank's window
    and the table segments its pusher CTAs copy to the peers."""

    def __init__(self, window_ptr, world, rank, rows_per_rank, timeout_ms, windows=None, offsets=None,
                 bytes_per_rank=None, pushers=0):
        x = _lib.FwdExchange()
        x.window = window_ptr
        x.world, x.rank, x.rows_per_rank, x.timeout_ms = int(world), int(rank), int(rows_per_rank), int(timeout_ms)
        self._keep = None
        if windows is not None and pushers > 0:
            n = len(offsets)
            offs = (ctypes.c_int64 * n)(*[int(o) for o in offsets])
            nb = (ctypes.c_int64 * n)(*[int(b) for b in bytes_per_rank])
            x.h_windows = ctypes.cast(windows, ctypes.POINTER(ctypes.c_void_p))
            x.n_seg, x.h_offsets, x.h_bytes, x.pushers = n, offs, nb, int(pushers)
            self._keep = (windows, offs, nb)
        self.struct = x

    def ref(self):
        return ctypes.byref(self.struct)


def _check_graph(ptr, idx, val):
    _cuda_contig(ptr, "ptr")
    _cuda_contig(idx, "idx")
    _cuda_contig(val, "val")
    _chk(ptr.dtype == torch.int32, "ptr must be int32")
    _chk(idx.dtype == torch.int32, "idx must be int32")
    _chk(val.dtype == torch.float32, "val must be float32")


def spgemm_forward(ptr, idx, val, sp_data, sp_index, num_nodes, num_edges, dim_sparse, dim_origin,
                   *, allow_banked: bool = True):
    """out[r, sp_index[j,t]] += val[e] * sp_data[j,t] over the stored entries e=(r<-j) of the
    CSR (ptr, idx, val).  Returns `(out fp32 [num_nodes, dim_origin], sp_index)` like the
    reference (spgemm_forward_cuda, maxk_cuda_kernels.o@0x1260)."""
    global _launches
    _check_graph(ptr, idx, val)
    _cuda_contig(sp_data, "sp_data")
    _cuda_contig(sp_index, "sp_index")
    _chk(sp_data.dtype == torch.float32, "sp_data must be float32")
    _chk(sp_data.dim() == 2 and sp_index.shape == sp_data.shape, "sp_index must have the shape of sp_data")
    _chk(sp_data.shape[1] == dim_sparse, "dim_sparse must equal sp_data.size(1)")
    _chk(1 <= dim_sparse <= dim_origin, "k must be between 1 and input dimension")
    _chk(idx.numel() >= num_edges and val.numel() >= num_edges, "idx/val must hold num_edges entries")
    ib = _index_bytes(sp_index, dim_origin)
    part = partition(ptr, num_nodes)
    if allow_banked and use_banked(part.num_parts, num_edges, dim_sparse, dim_origin):
        bk_data, _, bk_slot = cbsr_bank(sp_data, sp_index, dim_origin, with_index=False)
        out = spgemm_forward_banked(ptr, idx, val, bk_data, bk_slot, num_nodes, num_edges,
                                    dim_sparse, dim_origin)
        return out, sp_index
    if allow_banked and use_packed(part.num_parts, num_edges, dim_sparse, dim_origin):
        bk_pack = cbsr_bank_packed(sp_data, sp_index, dim_origin)
        out = spgemm_forward_packed(ptr, idx, val, bk_pack, num_nodes, num_edges, dim_sparse, dim_origin)
        return out, sp_index
    out = torch.empty((num_nodes, dim_origin), dtype=torch.float32, device=sp_data.device)
    partial = part.partial_for(dim_origin, sp_data.device)
    with torch.cuda.device(sp_data.device):
        rc = _lib.lib().mk_spgemm_fwd(
            part.parts.data_ptr(), part.num_parts, part.num_slots, idx.data_ptr(), val.data_ptr(),
            sp_data.data_ptr(), sp_index.data_ptr(), ib, out.data_ptr(),
            partial.data_ptr() if partial is not None else None, num_nodes, dim_sparse, dim_origin,
            _stream())
    _lib.check(rc, "mk_spgemm_fwd")
    _launches += 1 + (1 if part.num_slots else 0)
    return out, sp_index


# Column-blocked, row-tiled backward (csrc/sspmm_bwd.cu, mk_sspmm_bwd_tiled) for CBSR gradients that do
# not fit L2: "auto" = when n_src * k * 4 bytes exceed _BWD_TILED_MIN_MB, "0" / "1" force it off / on.
_BWD_TILED = os.environ.get("MAXK_BWD_TILED", "auto")
_BWD_TILED_MIN_MB = int(os.environ.get("MAXK_BWD_TILED_MIN_MB", "112"))
_BWD_TILE_MB = int(os.environ.get("MAXK_BWD_TILE_MB", "96"))     # slice of dXs one column block covers (products shape, k = 32: 5.70 ms at 64, 5.34 at 80-96, 5.49 at 112-128, 6.22 at 160; profiles/r2/bwd_tile_sizes_call27.log)
_blkptr_cache = {}


def set_backward_tiled(mode, tile_mb: Optional[int] = None) -> None:
    """mode: "auto", True / "1", False / "0"; tile_mb: MB of the CBSR gradient per column block."""
    global _BWD_TILED, _BWD_TILE_MB
    _BWD_TILED = mode if isinstance(mode, str) else ("1" if mode else "0")
    if tile_mb is not None:
        _BWD_TILE_MB = max(int(tile_mb), 1)


def backward_tiles(n_src: int, dim_sparse: int, dim_origin: int) -> int:
    """Column blocks the tiled backward would use for this shape (0: the plain backward runs)."""
    if _BWD_TILED == "0" or dim_sparse not in (8, 16, 32, 64) or dim_origin % 4:
        return 0
    nbytes = n_src * dim_sparse * 4
    if _BWD_TILED == "auto" and nbytes <= (_BWD_TILED_MIN_MB << 20):
        return 0
    return max(2, min(64, -(-nbytes // (_BWD_TILE_MB << 20))))


def _rows_ascending(ptr: torch.Tensor, idx: torch.Tensor, num_nodes: int) -> bool:
    """Column ids ascending inside every CSR row?  (One pass, one sync; callers cache the answer.)"""
    e = int(idx.numel())
    if e < 2:
        return True
    bad = idx[1:] < idx[:-1]
    starts = ptr[1:num_nodes].to(torch.int64)
    starts = starts[(starts > 0) & (starts < e)]
    bad[starts - 1] = False
    return not bool(bad.any())


def block_pointers(ptr: torch.Tensor, idx: torch.Tensor, num_nodes: int, n_blocks: int, width: int) -> torch.Tensor:
    """int32 [(n_blocks+1) * num_nodes]: blk[b*n + r] = first position of CSR row r whose column id is
    >= b * width (mk_block_ptr).  None when the column ids of some row are not ascending (the
    caller then stays on the un-blocked kernel).  Cached per graph."""
    global _launches
    key = (ptr.device.index, ptr.data_ptr(), idx.data_ptr(), int(num_nodes), int(n_blocks), int(width))
    hit = _blkptr_cache.get(key)
    if hit is not None and hit[0]() is ptr and hit[1] == ptr._version:
        return hit[2]
    blk = None
    if _rows_ascending(ptr, idx, num_nodes):
        blk = torch.empty(((n_blocks + 1) * num_nodes,), dtype=torch.int32, device=ptr.device)
        with torch.cuda.device(ptr.device):
            _lib.check(_lib.lib().mk_block_ptr(ptr.data_ptr(), idx.data_ptr(), num_nodes, n_blocks, width,
                                               blk.data_ptr(), _stream()), "mk_block_ptr")
        _launches += 1
    if len(_blkptr_cache) >= 16:
        del _blkptr_cache[next(iter(_blkptr_cache))]
    _blkptr_cache[key] = (weakref.ref(ptr), ptr._version, blk)
    return blk


def spgemm_backward(ptr, idx, val, grad_output, sp_index, num_nodes, num_edges, dim_sparse, dim_origin,
                    *, out: Optional[torch.Tensor] = None):
    """dXs[j,t] = sum over stored e=(r<-j) of val[e] * grad_output[r, sp_index[j,t]]; fp32
    [sp_index.size(0), dim_sparse] (spgemm_backward_cuda, maxk_cuda_kernels.o@0x1550).
    `out` (keyword, not in the reference): write into this fp32 [n_src, dim_sparse] buffer, e.g. a
    peer window (peer.py), instead of a new tensor."""
    global _launches
    _check_graph(ptr, idx, val)
    _cuda_contig(grad_output, "grad_output")
    _cuda_contig(sp_index, "sp_index")
    _chk(grad_output.dtype == torch.float32, "grad_output must be float32")
    _chk(grad_output.dim() == 2, "grad_output must be 2D tensor")
    _chk(grad_output.shape[0] == num_nodes and grad_output.shape[1] == dim_origin,
         "grad_output must be [num_nodes, dim_origin]")
    _chk(sp_index.dim() == 2 and sp_index.shape[1] == dim_sparse, "dim_sparse must equal sp_index.size(1)")
    _chk(1 <= dim_sparse <= dim_origin, "k must be between 1 and input dimension")
    ib = _index_bytes(sp_index, dim_origin)
    n_src = sp_index.shape[0]
    nb = backward_blocks(n_src, dim_sparse, num_nodes, num_edges)
    part = (partition(ptr, num_nodes, _BWD_MAX_NZ or None) if nb <= 1
            else partition_blocked(ptr, idx, num_nodes, n_src, nb))
    parts = part.exec_parts()
    if out is None:
        dxs = torch.empty((n_src, dim_sparse), dtype=torch.float32, device=grad_output.device)
    else:
        _cuda_contig(out, "out")
        _chk(out.dtype == torch.float32 and tuple(out.shape) == (n_src, dim_sparse),
             "out must be float32 [sp_index.size(0), dim_sparse]")
        dxs = out
    tma = _BWD_TMA if (ib == 1 and dim_sparse == 32 and _BWD_TMA in (1, 2, 4)) else 0
    tiles = backward_tiles(n_src, dim_sparse, dim_origin) if (nb <= 1 and not tma and num_nodes > 0) else 0
    with torch.cuda.device(grad_output.device):
        blk = None
        if tiles:
            width = -(-n_src // tiles)
            blk = block_pointers(ptr, idx, num_nodes, tiles, width)
        if blk is not None:
            rc = _lib.lib().mk_sspmm_bwd_tiled(
                blk.data_ptr(), tiles, idx.data_ptr(), val.data_ptr(), grad_output.data_ptr(),
                sp_index.data_ptr(), ib, dxs.data_ptr(), num_nodes, n_src, dim_sparse, dim_origin, _stream())
        elif tma:
            rc = _lib.lib().mk_sspmm_bwd_tma(
                parts.data_ptr(), part.num_parts, idx.data_ptr(), val.data_ptr(),
                grad_output.data_ptr(), sp_index.data_ptr(), ib, dxs.data_ptr(), num_nodes, n_src,
                dim_sparse, dim_origin, tma, _stream())
        else:
            rc = _lib.lib().mk_sspmm_bwd(
                parts.data_ptr(), part.num_parts, idx.data_ptr(), val.data_ptr(),
                grad_output.data_ptr(), sp_index.data_ptr(), ib, dxs.data_ptr(), num_nodes, n_src,
                dim_sparse, dim_origin, _stream())
    _lib.check(rc, "mk_sspmm_bwd")
    _launches += 2  # memset + kernel
    return dxs


# ---------------------------------------------------------------------------------------
# banked CBSR: the conflict-free internal form (not in the reference; see csrc/bank.cu)
# ---------------------------------------------------------------------------------------
def banked_supported(k: int, dim_origin: int) -> bool:
    return bool(_lib.lib().mk_banked_supported(int(k), int(dim_origin)))


def cbsr_bank(sp_data: torch.Tensor, sp_index: torch.Tensor, dim_origin: int, with_index: bool = True,
              *, out=None):
    """(sp_data, sp_index) -> (bk_data, bk_index, bk_slot): every row re-ordered and every entry
    given one of two shared-memory cells so that the aggregation kernels run without bank
    conflicts.  `(bk_data, bk_index)` is the same CBSR row in another entry order; with
    `with_index=False` bk_index is not produced (the forward kernel does not read it).
    `out = (bk_data, bk_slot)`: write into these [n, k] fp32 / int16 buffers (rows of a peer window)."""
    global _launches
    _cuda_contig(sp_data, "sp_data")
    _cuda_contig(sp_index, "sp_index")
    _chk(sp_data.dtype == torch.float32, "sp_data must be float32")
    _chk(sp_data.dim() == 2 and sp_index.shape == sp_data.shape, "sp_index must have the shape of sp_data")
    n, k = sp_data.shape
    _chk(banked_supported(k, dim_origin), "banked CBSR needs k in {8,16,32,64}, dim % 8 == 0, dim <= 512")
    ib = _index_bytes(sp_index, dim_origin)
    if out is None:
        bk_data = torch.empty_like(sp_data)
        bk_slot = torch.empty((n, k), dtype=torch.int16, device=sp_data.device)
    else:
        bk_data, bk_slot = out
        _cuda_contig(bk_data, "out")
        _cuda_contig(bk_slot, "out")
        _chk(bk_data.dtype == torch.float32 and bk_slot.dtype in (torch.int16, torch.uint16)
             and tuple(bk_data.shape) == (n, k) and tuple(bk_slot.shape) == (n, k),
             "out must be (float32 [n,k], int16 [n,k])")
    bk_index = torch.empty_like(sp_index) if with_index else None
    with torch.cuda.device(sp_data.device):
        rc = _lib.lib().mk_cbsr_bank(sp_data.data_ptr(), sp_index.data_ptr(), ib, bk_data.data_ptr(),
                                     bk_index.data_ptr() if with_index else None, bk_slot.data_ptr(),
                                     n, k, dim_origin, _stream())
    _lib.check(rc, "mk_cbsr_bank")
    _launches += 1
    return bk_data, bk_index, bk_slot


_split_cache = {}


def block_split(ptr: torch.Tensor, idx: torch.Tensor, num_nodes: int, world: int, rank: int,
                rows_per_rank: int) -> torch.Tensor:
    """int32 [num_nodes]: for every CSR row (ascending global column ids) the position in `idx` of
    its first stored entry whose column belongs to rank `rank` or a later rank.  The row-partitioned
    forward walks [split, end) first and [begin, split) second: own block, rank+1, ..., rank-1 --
    the order in which the peers' rows arrive (peer.py).  Built once per graph on the GPU."""
    global _launches
    key = (ptr.device.index, ptr.data_ptr(), idx.data_ptr(), int(num_nodes), int(world), int(rank), int(rows_per_rank))
    hit = _split_cache.get(key)
    if hit is not None and hit[0]() is ptr and hit[1] == ptr._version:
        return hit[2]
    blk = torch.empty(((world + 1) * num_nodes,), dtype=torch.int32, device=ptr.device)
    with torch.cuda.device(ptr.device):
        _lib.check(_lib.lib().mk_block_ptr(ptr.data_ptr(), idx.data_ptr(), num_nodes, world, rows_per_rank,
                                           blk.data_ptr(), _stream()), "mk_block_ptr")
    _launches += 1
    split = blk[rank * num_nodes:(rank + 1) * num_nodes].clone()
    if len(_split_cache) >= 32:
        del _split_cache[next(iter(_split_cache))]
    _split_cache[key] = (weakref.ref(ptr), ptr._version, split)
    return split


def maxk_forward_banked(input: torch.Tensor, k: int):
    """top-k -> CBSR -> banked form in one call: (bk_data, bk_index, bk_slot)."""
    sp_data, sp_index = maxk_forward_cbsr(input, k)
    return cbsr_bank(sp_data, sp_index, input.shape[1])


def forward_phases(world: int, rank: int):
    """Source-block phases of a row-partitioned forward, in arrival order: own block, the next
    ~3/7 of the senders, the rest -- as block ranges `(a0, a1, b0, b1)` for `mk_fwd_phase` ("rank,
    rank+1, ..." wraps around, hence two ranges).  8 ranks: {r}, {r+1..r+3}, {r+4..r+7}."""
    world, rank = int(world), int(rank)
    if world <= 1:
        return [(0, 1, 0, 0)]
    cuts = sorted({0, 1, min(world, 1 + max(1, round((world - 1) * 3 / 7))), world})
    out = []
    for s0, s1 in zip(cuts[:-1], cuts[1:]):
        lo, hi = rank + s0, rank + s1
        if hi <= world:
            out.append((lo, hi, 0, 0))
        elif lo >= world:
            out.append((lo - world, hi - world, 0, 0))
        else:
            out.append((lo, world, 0, hi - world))
    return out


def spgemm_forward_banked(ptr, idx, val, bk_data, bk_slot, num_nodes, num_edges, dim_sparse, dim_origin,
                          *, split: Optional[torch.Tensor] = None, wait=None, phases=None, blk=None,
                          n_blocks: int = 0):
    """`spgemm_forward` on a banked table: same result, no shared-memory bank conflicts.
    Row-partitioned form (dist.py): `split` int32 [num_nodes] makes every record walk the source
    blocks in arrival order, `wait` (a `ForwardExchange`) lets the kernel run while the peers' rows are
    still arriving and, with pushers, makes it the all-gather itself (peer.py).
    `phases` (from `forward_phases`) + `blk` (`block_pointers` over `n_blocks` source blocks): one launch
    per phase, each restricted to its blocks and adding to the rows of the earlier ones, so that whole
    launches overlap the transfer; `wait` may then be a list with one exchange per phase."""
    global _launches
    _check_graph(ptr, idx, val)
    _cuda_contig(bk_data, "sp_data")
    _cuda_contig(bk_slot, "sp_index")
    _chk(bk_data.dtype == torch.float32, "sp_data must be float32")
    _chk(bk_slot.dtype in (torch.int16, torch.uint16) and bk_slot.shape == bk_data.shape,
         "bk_slot must be 16-bit with the shape of bk_data")
    _chk(bk_data.shape[1] == dim_sparse, "dim_sparse must equal sp_data.size(1)")
    if split is not None:
        _cuda_contig(split, "split")
        _chk(split.dtype == torch.int32 and split.numel() >= num_nodes, "split must be int32 [num_nodes]")
    part = partition(ptr, num_nodes)
    out = torch.empty((num_nodes, dim_origin), dtype=torch.float32, device=bk_data.device)
    partial = part.partial_for(dim_origin, bk_data.device)
    ex = part.exec_parts()
    L = _lib.lib()
    if phases is not None:
        _chk(blk is not None and blk.dtype == torch.int32 and blk.numel() >= (n_blocks + 1) * num_nodes,
             "phases need the block pointers of the graph")
        waits = wait if isinstance(wait, (list, tuple)) else [wait] * len(phases)
        _chk(len(waits) == len(phases), "one exchange per phase")
        with torch.cuda.device(bk_data.device):
            for i, (a0, a1, b0, b1) in enumerate(phases):
                ph = _lib.FwdPhase(blk.data_ptr(), num_nodes, int(n_blocks), a0, a1, b0, b1,
                                   1 if i > 0 else 0, 1 if i == len(phases) - 1 else 0)
                w = waits[i]
                rc = L.mk_spgemm_fwd_banked_phase(
                    part.parts.data_ptr(), part.num_parts, part.num_slots,
                    ex.data_ptr() if ex is not part.parts else None, idx.data_ptr(), val.data_ptr(),
                    bk_data.data_ptr(), bk_slot.data_ptr(), out.data_ptr(),
                    partial.data_ptr() if partial is not None else None, num_nodes, dim_sparse, dim_origin,
                    w.ref() if w is not None else None, ctypes.byref(ph), _stream())
                _lib.check(rc, "mk_spgemm_fwd_banked_phase")
                _launches += 1 + (1 if part.num_slots else 0)
        return out
    with torch.cuda.device(bk_data.device):
        rc = L.mk_spgemm_fwd_banked_ex(
            part.parts.data_ptr(), part.num_parts, part.num_slots,
            ex.data_ptr() if ex is not part.parts else None, idx.data_ptr(), val.data_ptr(),
            bk_data.data_ptr(), bk_slot.data_ptr(), out.data_ptr(),
            partial.data_ptr() if partial is not None else None, num_nodes, dim_sparse, dim_origin,
            split.data_ptr() if split is not None else None, wait.ref() if wait is not None else None,
            _stream())
    _lib.check(rc, "mk_spgemm_fwd_banked_ex")
    _launches += 1 + (1 if part.num_slots else 0)
    return out


def spgemm_forward_ln(ptr, idx, val, table, bk_slot, num_nodes, num_edges, dim_sparse, dim_origin,
                      h_self, bias, gamma, beta, eps: float, *, keep_stats: bool = True):
    """Forward SpGEMM on a banked (`bk_slot` given) or packed (`bk_slot=None`) table with the layer's
    epilogue applied to every finished row (mk_spgemm_fwd_banked_ln, f-3):
        y = LayerNorm(h_self + A x Xs + bias) * gamma + beta
    Returns `(y, z, mean, rstd)`; z / mean / rstd (what `layernorm_backward` needs) are None with
    `keep_stats=False`.  Bit-identical to `spgemm_forward_banked` + `add_layernorm_forward`."""
    global _launches
    _check_graph(ptr, idx, val)
    _cuda_contig(table, "sp_data")
    _chk(dim_origin % 4 == 0 and dim_origin <= 512, "fused epilogue needs dim % 4 == 0, dim <= 512")
    for t, name in ((h_self, "h_self"), (bias, "bias"), (gamma, "gamma"), (beta, "beta")):
        if t is not None:
            _cuda_contig(t, name)
            _chk(t.dtype == torch.float32, f"{name} must be float32")
    _chk(h_self is None or tuple(h_self.shape) == (num_nodes, dim_origin), "h_self must be [num_nodes, dim]")
    _chk(gamma.numel() == dim_origin and beta.numel() == dim_origin, "gamma / beta must have dim entries")
    dev = table.device
    part = partition(ptr, num_nodes)
    y = torch.empty((num_nodes, dim_origin), dtype=torch.float32, device=dev)
    z = torch.empty_like(y) if keep_stats else None
    mean = torch.empty((num_nodes,), dtype=torch.float32, device=dev) if keep_stats else None
    rstd = torch.empty((num_nodes,), dtype=torch.float32, device=dev) if keep_stats else None
    partial = part.partial_for(dim_origin, dev)
    ex = part.exec_parts()
    ep = _lib.FwdEpilogue(h_self.data_ptr() if h_self is not None else None,
                          bias.data_ptr() if bias is not None else None, gamma.data_ptr(), beta.data_ptr(),
                          z.data_ptr() if keep_stats else None, mean.data_ptr() if keep_stats else None,
                          rstd.data_ptr() if keep_stats else None, float(eps))
    with torch.cuda.device(dev):
        rc = _lib.lib().mk_spgemm_fwd_banked_ln(
            part.parts.data_ptr(), part.num_parts, part.num_slots,
            ex.data_ptr() if ex is not part.parts else None, idx.data_ptr(), val.data_ptr(),
            table.data_ptr(), bk_slot.data_ptr() if bk_slot is not None else None, y.data_ptr(),
            partial.data_ptr() if partial is not None else None, num_nodes, dim_sparse, dim_origin,
            ctypes.byref(ep), _stream())
    _lib.check(rc, "mk_spgemm_fwd_banked_ln")
    _launches += 1 + (1 if part.num_slots else 0)
    return y, z, mean, rstd


def packed_supported(k: int, dim_origin: int) -> bool:
    return bool(_lib.lib().mk_packed_supported(int(k), int(dim_origin)))


def cbsr_bank_packed(sp_data: torch.Tensor, sp_index: torch.Tensor, dim_origin: int, *, out=None) -> torch.Tensor:
    """(sp_data, sp_index) -> bk_pack int32 [n, k, 2]: the banked row with value bits, cell offset and
    column of every entry in 8 bytes (k = 8, 16).  `out`: write into this buffer (rows of a peer window)."""
    global _launches
    _cuda_contig(sp_data, "sp_data")
    _cuda_contig(sp_index, "sp_index")
    _chk(sp_data.dtype == torch.float32, "sp_data must be float32")
    _chk(sp_data.dim() == 2 and sp_index.shape == sp_data.shape, "sp_index must have the shape of sp_data")
    n, k = sp_data.shape
    _chk(packed_supported(k, dim_origin), "packed CBSR needs k in {8,16}, dim % 8 == 0, dim <= 512")
    ib = _index_bytes(sp_index, dim_origin)
    if out is None:
        out = torch.empty((n, k, 2), dtype=torch.int32, device=sp_data.device)
    else:
        _cuda_contig(out, "out")
        _chk(out.dtype == torch.int32 and tuple(out.shape) == (n, k, 2), "out must be int32 [n, k, 2]")
    with torch.cuda.device(sp_data.device):
        rc = _lib.lib().mk_cbsr_bank_packed(sp_data.data_ptr(), sp_index.data_ptr(), ib, out.data_ptr(),
                                            n, k, dim_origin, _stream())
    _lib.check(rc, "mk_cbsr_bank_packed")
    _launches += 1
    return out


def spgemm_forward_packed(ptr, idx, val, bk_pack, num_nodes, num_edges, dim_sparse, dim_origin,
                          *, split: Optional[torch.Tensor] = None, wait=None):
    """`spgemm_forward` on a packed banked table (k = 8, 16); `split` / `wait` as in
    `spgemm_forward_banked`."""
    global _launches
    _check_graph(ptr, idx, val)
    _cuda_contig(bk_pack, "sp_data")
    _chk(bk_pack.dtype == torch.int32 and bk_pack.dim() == 3 and bk_pack.shape[1] == dim_sparse
         and bk_pack.shape[2] == 2, "bk_pack must be int32 [n_src, dim_sparse, 2]")
    if split is not None:
        _cuda_contig(split, "split")
        _chk(split.dtype == torch.int32 and split.numel() >= num_nodes, "split must be int32 [num_nodes]")
    part = partition(ptr, num_nodes)
    out = torch.empty((num_nodes, dim_origin), dtype=torch.float32, device=bk_pack.device)
    partial = part.partial_for(dim_origin, bk_pack.device)
    ex = part.exec_parts()
    with torch.cuda.device(bk_pack.device):
        rc = _lib.lib().mk_spgemm_fwd_packed_ex(
            part.parts.data_ptr(), part.num_parts, part.num_slots,
            ex.data_ptr() if ex is not part.parts else None, idx.data_ptr(), val.data_ptr(),
            bk_pack.data_ptr(), out.data_ptr(), partial.data_ptr() if partial is not None else None,
            num_nodes, dim_sparse, dim_origin, split.data_ptr() if split is not None else None,
            wait.ref() if wait is not None else None, _stream())
    _lib.check(rc, "mk_spgemm_fwd_packed_ex")
    _launches += 1 + (1 if part.num_slots else 0)
    return out


def spgemm_backward_banked(ptr, idx, val, grad_output, bk_slot, num_nodes, num_edges, dim_sparse, dim_origin):
    """`spgemm_backward` on a banked table; the [n_src, k] gradient comes out in the BANKED entry
    order (pair it with bk_index)."""
    global _launches
    _check_graph(ptr, idx, val)
    _cuda_contig(grad_output, "grad_output")
    _cuda_contig(bk_slot, "sp_index")
    _chk(grad_output.dtype == torch.float32, "grad_output must be float32")
    _chk(grad_output.dim() == 2 and grad_output.shape[0] == num_nodes and grad_output.shape[1] == dim_origin,
         "grad_output must be [num_nodes, dim_origin]")
    _chk(bk_slot.dtype in (torch.int16, torch.uint16) and bk_slot.shape[1] == dim_sparse,
         "bk_slot must be 16-bit [n_src, dim_sparse]")
    n_src = bk_slot.shape[0]
    part = partition(ptr, num_nodes)
    dxs = torch.empty((n_src, dim_sparse), dtype=torch.float32, device=grad_output.device)
    with torch.cuda.device(grad_output.device):
        rc = _lib.lib().mk_sspmm_bwd_banked(
            part.exec_parts().data_ptr(), part.num_parts, idx.data_ptr(), val.data_ptr(),
            grad_output.data_ptr(), bk_slot.data_ptr(), dxs.data_ptr(), num_nodes, n_src,
            dim_sparse, dim_origin, _stream())
    _lib.check(rc, "mk_sspmm_bwd_banked")
    _launches += 2
    return dxs


# ---------------------------------------------------------------------------------------
# f-3: fused  z = a + b + bias ; y = LayerNorm(z)  (csrc/layernorm.cu)
# ---------------------------------------------------------------------------------------
def add_layernorm_supported(a: torch.Tensor, d: int) -> bool:
    return bool(a.is_cuda and a.dtype == torch.float32 and d % 4 == 0 and 4 <= d <= 1024)


def add_layernorm_forward(a, b, bias, gamma, beta, eps: float):
    """(y, z, mean, rstd) with z = a + b + bias (b, bias optional) and y = LayerNorm(z)*gamma+beta."""
    global _launches
    _cuda_contig(a, "input")
    n, d = a.shape
    _chk(add_layernorm_supported(a, d), "fused LayerNorm needs CUDA float32, dim % 4 == 0, dim <= 1024")
    for t, name in ((b, "b"), (bias, "bias"), (gamma, "gamma"), (beta, "beta")):
        if t is not None:
            _cuda_contig(t, name)
            _chk(t.dtype == torch.float32, f"{name} must be float32")
    y = torch.empty_like(a)
    z = torch.empty_like(a)
    mean = torch.empty((n,), dtype=torch.float32, device=a.device)
    rstd = torch.empty((n,), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        rc = _lib.lib().mk_add_layernorm_fwd(
            a.data_ptr(), b.data_ptr() if b is not None else None,
            bias.data_ptr() if bias is not None else None, gamma.data_ptr(), beta.data_ptr(),
            z.data_ptr(), y.data_ptr(), mean.data_ptr(), rstd.data_ptr(), n, d, float(eps), _stream())
    _lib.check(rc, "mk_add_layernorm_fwd")
    _launches += 1
    return y, z, mean, rstd


def layernorm_backward(grad_y, z, gamma, mean, rstd, want_dbias: bool = False):
    """(gz, dgamma, dbeta, dbias) of the fused LayerNorm; dbias = column sums of gz or None."""
    global _launches
    _cuda_contig(grad_y, "grad_output")
    n, d = z.shape
    L = _lib.lib()
    gz = torch.empty_like(z)
    dgamma = torch.empty((d,), dtype=torch.float32, device=z.device)
    dbeta = torch.empty((d,), dtype=torch.float32, device=z.device)
    dbias = torch.empty((d,), dtype=torch.float32, device=z.device) if want_dbias else None
    ws = torch.empty((3 * L.mk_layernorm_parts() * d,), dtype=torch.float32, device=z.device)
    with torch.cuda.device(z.device):
        rc = L.mk_layernorm_bwd(grad_y.data_ptr(), z.data_ptr(), gamma.data_ptr(), mean.data_ptr(),
                                rstd.data_ptr(), gz.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(),
                                dbias.data_ptr() if want_dbias else None, ws.data_ptr(), n, d, _stream())
    _lib.check(rc, "mk_layernorm_bwd")
    _launches += 2
    return gz, dgamma, dbeta, dbias
