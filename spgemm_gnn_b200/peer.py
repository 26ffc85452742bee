"""Peer-memory exchange of the row-partitioned hot path (csrc/peer.cu).

The two exchanges of a sharded layer -- all-gather of the CBSR table in front of the forward
SpGEMM, reduce-scatter of the CBSR gradient behind the backward SSpMM -- as this library's own
kernels over NVLink instead of NCCL calls:

  * `bank_push`       csrc/bank.cu writes the banked rows it produces straight into every rank's
                      table (compute + all-gather in one kernel);
  * `allgather`       the plain form for the un-banked table (stores into every rank's table);
  * `reduce_scatter`  every rank loads its block from every rank's partial buffer and folds it in
                      rank order (bit-reproducible).

Rules of use: every rank issues the same collectives through a window in the same order, and a
rank issues them from one stream at a time (stream order is what tells the peers that the rank is
done with its copy; two streams racing through one window would break that).

A `PeerWindow` is one device buffer per rank, mapped into every process of the group with CUDA IPC
(`mk_peer_export` / `mk_peer_open`, handles exchanged with `all_gather_object`).  The kernels keep
their flags and epoch counters in the window's header, so captured CUDA graphs replay correctly.

On by default for NCCL groups of 2..16 ranks (checked against the NCCL path at 2 and 8 GPUs:
forward bit-equal, profiles/r1_peer_exchange.md); `MAXK_PEER_EXCHANGE=0` (or `set_enabled(False)`)
keeps dist.py on NCCL, and so does any failure to map the windows (agreed on by all ranks).  There
is no CPU form: on a gloo group `available()` is False and dist.py stays on its collectives.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import _lib

HEADER_BYTES = 1024          # MK_PEER_HEADER_BYTES
HANDLE_BYTES = 64            # MK_PEER_HANDLE_BYTES
MAX_RANKS = 16               # MK_PEER_MAX_RANKS
_ALIGN = 256

_ENABLED = os.environ.get("MAXK_PEER_EXCHANGE", "1") != "0"   # "0": stay on NCCL collectives
_TIMEOUT_MS = int(os.environ.get("MAXK_PEER_TIMEOUT_MS", "120000"))
# how bank_push reaches the peers: 1 = every row stored straight into all tables (measured),
# 2 = own table first, then each block copies its rows with 16-byte stores (experimental)
_PUSH_MODE = int(os.environ.get("MAXK_PEER_PUSH_MODE", "1"))
_launches = 0


def set_enabled(on: bool) -> None:
    global _ENABLED
    _ENABLED = bool(on)


def enabled() -> bool:
    return _ENABLED


def launch_count() -> int:
    return _launches


def available(group=None) -> bool:
    """The peer kernels need CUDA devices of one box under an NCCL group of at most 16 ranks."""
    if not (dist.is_available() and dist.is_initialized() and torch.cuda.is_available()):
        return False
    return dist.get_backend(group) == "nccl" and 1 <= dist.get_world_size(group) <= MAX_RANKS


def layout(segment_bytes: Sequence[int]) -> Tuple[List[int], int]:
    """Byte offsets (from the window base) of consecutive payload segments, each aligned to 256 B,
    and the window size that holds them."""
    offs, cur = [], HEADER_BYTES
    for b in segment_bytes:
        cur = (cur + _ALIGN - 1) // _ALIGN * _ALIGN
        offs.append(cur)
        cur += int(b)
    return offs, (cur + _ALIGN - 1) // _ALIGN * _ALIGN


class _RawCuda:
    """Device memory torch does not own, described through __cuda_array_interface__."""

    def __init__(self, ptr: int, nbytes: int, owner):
        self.owner = owner
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1",
                                         "data": (int(ptr), False), "version": 2}


class PeerWindow:
    """One rank's view of a window: its own buffer plus the peers' buffers as mapped here."""

    def __init__(self, nbytes: int, world: int, rank: int, device):
        self.nbytes, self.world, self.rank = int(nbytes), int(world), int(rank)
        self.device = torch.device(device)
        self.local = None          # own buffer (int address)
        self.opened: List[int] = []  # peers' buffers mapped with mk_peer_open
        self.ptrs = (ctypes.c_void_p * MAX_RANKS)()
        self._bytes = None

    # ---- construction -------------------------------------------------------------------
    @classmethod
    def _alloc(cls, nbytes, world, rank, device) -> "PeerWindow":
        w = cls(nbytes, world, rank, device)
        p = ctypes.c_void_p(0)
        with torch.cuda.device(w.device):
            rc = _lib.lib().mk_peer_alloc(w.nbytes, ctypes.byref(p))
            if rc == _lib.MK_ECUDA:  # cudaMalloc next to torch's caching allocator: give its cache back once
                torch.cuda.empty_cache()
                rc = _lib.lib().mk_peer_alloc(w.nbytes, ctypes.byref(p))
            _lib.check(rc, "mk_peer_alloc")
        w.local = int(p.value)
        w.ptrs[rank] = w.local
        return w

    @classmethod
    def create(cls, nbytes: int, group=None, device=None) -> Optional["PeerWindow"]:
        """Collective over `group`: allocate, export, exchange handles, map every peer.  A failure
        on any rank (no IPC in this sandbox, out of memory, ...) is agreed on by all ranks, which
        then all return None -- nobody is left waiting in a collective."""
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if world > MAX_RANKS:
            raise RuntimeError(f"peer windows support at most {MAX_RANKS} ranks")
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        L = _lib.lib()
        w, handle, why = None, None, ""
        try:
            w = cls._alloc(nbytes, world, rank, device)
            buf = ctypes.create_string_buffer(HANDLE_BYTES)
            _lib.check(L.mk_peer_export(w.local, buf), "mk_peer_export")
            handle = bytes(buf.raw)
        except Exception as exc:  # noqa: BLE001 -- reported below, after everybody has met
            why = f"rank {rank}: {exc}"
        got: List[Optional[tuple]] = [None] * world
        dist.all_gather_object(got, (handle, int(nbytes), why), group=group)
        ok = all(h is not None and nb == int(nbytes) for h, nb, _ in got)
        if ok:
            try:
                with torch.cuda.device(device):
                    for q, (h, _, _) in enumerate(got):
                        if q == rank:
                            continue
                        p = ctypes.c_void_p(0)
                        _lib.check(L.mk_peer_open(h, ctypes.byref(p)), "mk_peer_open")
                        w.opened.append(int(p.value))
                        w.ptrs[q] = int(p.value)
            except Exception as exc:  # noqa: BLE001
                ok, why = False, f"rank {rank}: {exc}"
        flag = torch.tensor([1 if ok else 0], device=device, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            reasons = "; ".join(r for _, _, r in got if r) or why or "a peer failed to map the window"
            if w is not None:
                w.close()
            if rank == 0:
                import warnings
                warnings.warn(f"peer windows unavailable ({reasons}); staying on NCCL collectives")
            return None
        return w

    @classmethod
    def create_virtual(cls, nbytes: int, world: int, device=None) -> List["PeerWindow"]:
        """`world` windows inside ONE process on one device (every "peer" is a local buffer): the
        kernels and their flag protocol can then be exercised on a single GPU, each virtual rank
        on its own stream (tests/test_gpu_peer.py)."""
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        wins = [cls._alloc(nbytes, world, r, device) for r in range(world)]
        for w in wins:
            for q in range(world):
                w.ptrs[q] = wins[q].local
        return wins

    # ---- access -------------------------------------------------------------------------
    def view(self, offset: int, shape, dtype) -> torch.Tensor:
        """Tensor over this rank's own buffer, `offset` bytes from the window base."""
        if self._bytes is None:
            with torch.cuda.device(self.device):
                self._bytes = torch.as_tensor(_RawCuda(self.local, self.nbytes, self), device=self.device)
        n = 1
        for s in shape:
            n *= int(s)
        nb = n * torch.empty((), dtype=dtype).element_size()
        if offset < HEADER_BYTES or offset + nb > self.nbytes:
            raise ValueError("view outside the window payload")
        return self._bytes[offset:offset + nb].view(dtype).view(*shape)

    def epoch(self) -> Tuple[int, int]:
        """(collectives completed, error word) -- synchronises the current stream."""
        e, err = ctypes.c_uint32(0), ctypes.c_uint32(0)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().mk_peer_epoch(self.local, ctypes.byref(e), ctypes.byref(err),
                                                torch.cuda.current_stream().cuda_stream), "mk_peer_epoch")
        return int(e.value), int(err.value)

    def close(self) -> None:
        L = _lib.lib()
        with torch.cuda.device(self.device):
            for p in self.opened:
                L.mk_peer_close(p)
            self.opened = []
            if self.local is not None:
                self._bytes = None
                L.mk_peer_free(self.local)
                self.local = None


# ---------------------------------------------------------------------------------------
# window cache (one window per use and size, shared by all layers of that shape)
# ---------------------------------------------------------------------------------------
_windows: Dict[tuple, PeerWindow] = {}


def window(kind: str, nbytes: int, group=None) -> Optional[PeerWindow]:
    """Cached window of exactly `nbytes` for `kind`; creating one is a collective, so every rank
    must ask for the same windows in the same order (they do: same model, same shapes).  Returns
    None -- on every rank -- when the windows cannot be set up; the peer path is then switched off
    for the rest of the process and dist.py stays on its NCCL collectives."""
    key = (kind, int(nbytes), id(group) if group is not None else 0, torch.cuda.current_device())
    w = _windows.get(key)
    if w is None:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("peer windows must exist before CUDA graph capture (run one eager step first)")
        w = PeerWindow.create(nbytes, group)
        if w is None:
            set_enabled(False)
            return None
        _windows[key] = w
    return w


def close_all() -> None:
    for w in _windows.values():
        w.close()
    _windows.clear()


# ---------------------------------------------------------------------------------------
# collectives
# ---------------------------------------------------------------------------------------
def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def allgather(win: PeerWindow, locals_: Sequence[torch.Tensor], offsets: Sequence[int],
              grid: int = 0) -> List[torch.Tensor]:
    """Every rank's `locals_[g]` ([R, ...], same shape on all ranks) -> [world*R, ...] at
    `offsets[g]` of every rank's window.  Returns the gathered tensors (views of the window)."""
    global _launches
    n = len(locals_)
    src = (ctypes.c_void_p * n)()
    nbytes = (ctypes.c_int64 * n)()
    offs = (ctypes.c_int64 * n)()
    outs = []
    for g, t in enumerate(locals_):
        if not (t.is_cuda and t.is_contiguous()):
            raise RuntimeError("peer all-gather wants contiguous CUDA tensors")
        b = t.numel() * t.element_size()
        if b % 16:
            raise RuntimeError("peer all-gather segments must be multiples of 16 bytes")
        src[g], nbytes[g], offs[g] = t.data_ptr(), b, int(offsets[g])
        outs.append(win.view(int(offsets[g]), (win.world * t.shape[0],) + tuple(t.shape[1:]), t.dtype))
    with torch.cuda.device(win.device):
        rc = _lib.lib().mk_peer_allgather(win.ptrs, win.world, win.rank, n, src, nbytes, offs, grid,
                                          _TIMEOUT_MS, _stream())
    _lib.check(rc, "mk_peer_allgather")
    _launches += 1
    return outs


def bank_push(win: PeerWindow, sp_data: torch.Tensor, sp_index: torch.Tensor, dim_origin: int,
              offsets: Sequence[int]):
    """Fused banking + all-gather: (full bk_data fp32, full bk_slot int16, full sorted sp_index),
    each [world*R, k], views of this rank's window at `offsets` = (data, slot, index)."""
    global _launches
    r, k = sp_data.shape
    if not (sp_data.is_cuda and sp_data.is_contiguous() and sp_index.is_contiguous()):
        raise RuntimeError("peer bank_push wants contiguous CUDA tensors")
    od, os_, oi = (int(o) for o in offsets)
    with torch.cuda.device(win.device):
        rc = _lib.lib().mk_peer_bank_push(sp_data.data_ptr(), sp_index.data_ptr(), sp_index.element_size(),
                                          win.ptrs, win.world, win.rank, od, os_, oi, r, k, dim_origin,
                                          2 if _PUSH_MODE == 2 else 1, _TIMEOUT_MS, _stream())
    _lib.check(rc, "mk_peer_bank_push")
    _launches += 1
    rows = win.world * r
    return (win.view(od, (rows, k), torch.float32), win.view(os_, (rows, k), torch.int16),
            win.view(oi, (rows, k), sp_index.dtype))


def reduce_scatter(win: PeerWindow, offset: int, rows: int, k: int, grid: int = 0) -> torch.Tensor:
    """Sum over ranks of rows [rank*rows, (rank+1)*rows) of the fp32 [world*rows, k] buffer at
    `offset` of every rank's window -> fp32 [rows, k]."""
    global _launches
    block = rows * k * 4
    if block % 16:
        raise RuntimeError("peer reduce-scatter blocks must be multiples of 16 bytes")
    out = torch.empty((rows, k), dtype=torch.float32, device=win.device)
    with torch.cuda.device(win.device):
        rc = _lib.lib().mk_peer_reduce_scatter(win.ptrs, win.world, win.rank, int(offset), block,
                                               out.data_ptr(), grid, _TIMEOUT_MS, _stream())
    _lib.check(rc, "mk_peer_reduce_scatter")
    _launches += 1
    return out
