"""Peer-memory exchange of the row-partitioned hot path (csrc/peer.cu).

The two exchanges of a sharded layer -- all-gather of the CBSR table in front of the forward
SpGEMM, reduce-scatter of the CBSR gradient behind the backward SSpMM -- over NVLink peer memory
instead of NCCL calls:

  * forward  `begin_push` -> the producer writes the rank's rows into its OWN window ->
             `publish_and_push` (copy engines, side stream: one transfer per peer and table, each
             peer's transfers followed by its `done` flag) -> the forward SpGEMM starts at once and
             waits per source block (`spgemm_forward_banked(wait=...)`) -> `join_push`, `release`.
             The transfer overlaps the kernel that consumes it; no SM copies anything.
  * backward `reduce_scatter`: every rank loads its block from every rank's partial buffer and
             folds it in rank order (bit-reproducible); the SSpMM wrote straight into the window.

Rules of use: every rank issues the same collectives through a window in the same order, and a
rank issues them from one stream at a time (stream order is what tells the peers that the rank is
done with its copy; two streams racing through one window would break that).

A `PeerWindow` is one device buffer per rank, mapped into every process of the group with CUDA IPC
(`mk_peer_export` / `mk_peer_open`, handles exchanged with `all_gather_object`).  The kernels keep
their flags and epoch counters in the window's header, so captured CUDA graphs replay correctly.

On by default for NCCL groups of 2..16 ranks; `MAXK_PEER_EXCHANGE=0` (or `set_enabled(False)`)
keeps dist.py on NCCL, and so does any failure to map the windows (agreed on by all ranks).  There
is no CPU form: on a gloo group `available()` is False and dist.py stays on its collectives.

A kernel that waits for a peer longer than `MAXK_PEER_TIMEOUT_MS` writes the window's error word and
gives up; `check_errors()` (one device synchronisation) turns that into a `PeerTimeoutError`.
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
# side streams (copy engines) the pushes of one collective are spread over; every stream visits its
# share of the peers nearest first
_PUSH_STREAMS = max(1, int(os.environ.get("MAXK_PEER_STREAMS", "2")))
# 1: the forward SpGEMM starts with the pushes and waits per source block; 0: it starts when the whole
# table has arrived (mk_peer_wait_all)
_OVERLAP = os.environ.get("MAXK_PEER_OVERLAP", "1") != "0"
# who moves the rows to the peers:
#   "auto"   (default) through the window's multicast address where there is one and more than two ranks
#            (mk_peer_push_mc: every row stored once, then the forward on the complete table), else "sm_seq";
#   "sm_seq" NVLink stores to every peer as a kernel of its own, then the forward on the complete table;
#   "sm"     pusher CTAs inside the forward SpGEMM kernel (progressive arrival, per-block waiting);
#   "dma"    copy engines on side streams (small transfers pay ~4 us each and concurrent flows interfere).
# Measured at 8 GPUs (profiles/r2/peer_mc8_call20.log, ms per layer, Reddit shape / 20 k-node graph):
# multicast 0.930 / 0.245, NCCL 0.953 / 0.65, "sm" 1.012 / 0.303, "dma" 1.025 / 0.326.
_PUSH = os.environ.get("MAXK_PEER_PUSH", "auto")
_PUSHERS = int(os.environ.get("MAXK_PEER_PUSHERS", "592"))   # pusher CTAs (32 threads each)
# forward cut into source-block phases (own block, the next senders, the rest: one launch each), so
# that whole launches overlap the transfer instead of the CTAs that happen to be resident.  Measured at
# 8 GPUs on the Reddit shape (profiles/r2/peer_phases8_call13.log): the three launches cost 0.13 ms more
# than one (every record zeroes, folds and writes its 8 KB of accumulators three times, and the rows
# are read back twice) and hide less than that of the ~0.12 ms all-gather -- 0.683 ms against 0.638 ms
# with the copy engines, 0.713 against 0.600 with pusher CTAs.  Off by default.
_PHASES = os.environ.get("MAXK_PEER_PHASES", "0") != "0"
# Above this table size, groups of 8 or more ranks use NCCL's all-gather / reduce-scatter instead of the
# peer kernels: the own kernels win where the exchange is latency-bound (20 k-node graph at 8 GPUs:
# 0.30 ms per layer against 0.69 ms) and lose a few per cent where it is bandwidth-bound (Reddit shape,
# 52 MB table: 1.014 against 0.953 ms per layer; products shape 2.75 against 2.68 ms -- NCCL's NVLS
# all-gather sends every row once, the push sends it seven times).  MAXK_PEER_MAX_MB=0: no limit.
_MAX_BYTES = int(os.environ.get("MAXK_PEER_MAX_MB", "32")) << 20
# NVLink multicast (NVLS): windows allocated as symmetric memory bound to a multicast object (torch's
# symmetric-memory allocator does the driver plumbing: cuMemCreate, handle exchange, cuMulticastBindMem).
# The all-gather then stores every row once (mk_peer_push_mc), the reduce-scatter loads one row reduced
# by the switch (mk_peer_reduce_scatter_mc).  "1": use it where the box offers it, fall back to CUDA-IPC
# windows otherwise; "0": CUDA-IPC windows only.
_MULTICAST = os.environ.get("MAXK_PEER_MULTICAST", "1") != "0"
_launches = 0


class PeerTimeoutError(RuntimeError):
    pass


def set_enabled(on: bool) -> None:
    global _ENABLED
    _ENABLED = bool(on)


def enabled() -> bool:
    return _ENABLED


def launch_count() -> int:
    return _launches


def timeout_ms() -> int:
    return _TIMEOUT_MS


def overlap() -> bool:
    return _OVERLAP


_mc_state = {"probed": False, "ok": False}


def set_multicast(on: bool) -> bool:
    """Use the multicast forms of the exchanges where the windows have a multicast address (returns the
    previous setting).  Windows that exist stay what they are; only the kernels change."""
    global _MULTICAST
    was, _MULTICAST = _MULTICAST, bool(on)
    return was


def multicast(group=None) -> bool:
    """Do the windows of this process group come with a multicast address?  Probed once with a small
    symmetric allocation (a collective: every rank calls it at the same point) and agreed on by all ranks."""
    if not (_MULTICAST and available(group)) or dist.get_world_size(group) < 2:
        return False
    if not _mc_state["probed"]:
        _mc_state["probed"] = True
        w = PeerWindow._create_symm(1 << 20, group, None)
        _mc_state["ok"] = w is not None and w.mc != 0
        if w is not None:
            w.close()
    return _mc_state["ok"]


def wanted(world: int, exchange_bytes: int, group=None) -> bool:
    """Peer kernels or NCCL for an exchange of `exchange_bytes` (whole table / whole gradient) over
    `world` ranks -- same answer on every rank (same shapes).  With multicast windows the own kernels
    move no more bytes than NCCL's NVLS collectives, so the size limit only applies without them."""
    if not _ENABLED:
        return False
    if _MAX_BYTES == 0 or world < 8 or exchange_bytes <= _MAX_BYTES:
        return True
    return multicast(group)


def set_max_mb(mb: int) -> int:
    """Size limit of `wanted` in MB (0: none); returns the previous one."""
    global _MAX_BYTES
    was = _MAX_BYTES >> 20
    _MAX_BYTES = max(int(mb), 0) << 20
    return was


def phases() -> bool:
    return _PHASES and _OVERLAP


def push_mode() -> str:
    return _PUSH if _OVERLAP else "dma"


def use_multicast(win: "PeerWindow") -> bool:
    """The multicast all-gather for this window?  Needs a multicast address; with two ranks there is
    nothing to replicate and the plain NVLink stores are faster (profiles/r2/peer_mc2_call19.log)."""
    return bool(win.mc) and _MULTICAST and _PUSH in ("auto", "mc") and (win.world > 2 or _PUSH == "mc")


def publish(win: "PeerWindow", buf: int) -> None:
    """Open the collective whose rows this rank has just written into its own window (no transfer:
    the consumer's pusher CTAs, or `push_dma`, move them)."""
    global _launches
    with torch.cuda.device(win.device):
        _lib.check(_lib.lib().mk_peer_publish(win.local, win.rank, int(buf), _stream()), "mk_peer_publish")
    _launches += 1


def push_sm(win: "PeerWindow", offsets: Sequence[int], bytes_per_rank: Sequence[int]) -> None:
    """The all-gather by NVLink stores as a kernel of its own (after `publish`, current stream)."""
    global _launches
    n = len(offsets)
    offs = (ctypes.c_int64 * n)(*[int(o) for o in offsets])
    nbytes = (ctypes.c_int64 * n)(*[int(b) for b in bytes_per_rank])
    with torch.cuda.device(win.device):
        rc = _lib.lib().mk_peer_push_sm(win.ptrs, win.world, win.rank, n, offs, nbytes, _PUSHERS, _stream())
    _lib.check(rc, "mk_peer_push_sm")
    _launches += 1


def push_mc(win: "PeerWindow", offsets: Sequence[int], bytes_per_rank: Sequence[int]) -> None:
    """The all-gather through the window's multicast address: every row stored once, the switch
    replicates it (after `publish`, current stream)."""
    global _launches
    n = len(offsets)
    offs = (ctypes.c_int64 * n)(*[int(o) for o in offsets])
    nbytes = (ctypes.c_int64 * n)(*[int(b) for b in bytes_per_rank])
    with torch.cuda.device(win.device):
        rc = _lib.lib().mk_peer_push_mc(win.ptrs, win.mc, win.world, win.rank, n, offs, nbytes, 0, _stream())
    _lib.check(rc, "mk_peer_push_mc")
    _launches += 1


def exchange(win: "PeerWindow", rows_per_rank: int, offsets=None, bytes_per_rank=None):
    """The `wait=` argument of the forward kernels for this window; with offsets, the kernel pushes."""
    from .maxk_kernels import ForwardExchange
    pushing = offsets is not None and win.world > 1
    return ForwardExchange(win.local, win.world, win.rank, rows_per_rank, _TIMEOUT_MS,
                           windows=win.ptrs if pushing else None, offsets=offsets,
                           bytes_per_rank=bytes_per_rank, pushers=_PUSHERS if pushing else 0)


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
        self._side = None          # stream of the copy-engine pushes
        self._pushed = None        # event behind the last push
        self._buf = 1              # table buffer of the last forward (alternates 0, 1, 0, ...)
        self.mc = 0                # multicast address of the window (symmetric-memory windows), 0 = none
        self._symm = None          # (tensor, handle) that own a symmetric-memory window

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
    def _create_symm(cls, nbytes: int, group, device) -> Optional["PeerWindow"]:
        """Window in symmetric memory (torch.distributed._symmetric_memory): every peer's copy mapped
        here plus, on NVSwitch boxes, one multicast address for all of them.  Collective; None on every
        rank if any rank fails."""
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        w, why = None, ""
        try:
            import torch.distributed._symmetric_memory as symm
            t = symm.empty(int(nbytes), dtype=torch.uint8, device=device)
            h = symm.rendezvous(t, group if group is not None else dist.group.WORLD)
            t.zero_()
            w = cls(nbytes, world, rank, device)
            ptrs = list(h.buffer_ptrs)
            w.local = int(ptrs[rank])
            for q in range(world):
                w.ptrs[q] = int(ptrs[q])
            w.mc = int(h.multicast_ptr or 0)
            w._symm = (t, h)
        except Exception as exc:  # noqa: BLE001 -- agreed on below
            w, why = None, f"rank {rank}: {type(exc).__name__}: {exc}"
        flag = torch.tensor([1 if w is not None else 0], device=device, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)   # also: every rank's header is zero
        if int(flag.item()) == 0:
            if w is not None:
                w.close()
            return None
        return w

    @classmethod
    def create(cls, nbytes: int, group=None, device=None) -> Optional["PeerWindow"]:
        """Collective over `group`: allocate, export, exchange handles, map every peer.  A failure
        on any rank (no IPC in this sandbox, out of memory, ...) is agreed on by all ranks, which
        then all return None -- nobody is left waiting in a collective.  Symmetric memory with a
        multicast address first (where the box has it), CUDA-IPC windows otherwise."""
        if _MULTICAST and _mc_state.get("ok", False):
            w = cls._create_symm(nbytes, group, device)
            if w is not None:
                return w
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

    def side_streams(self) -> List[torch.cuda.Stream]:
        if self._side is None:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("the push streams must exist before CUDA graph capture (run one eager step first)")
            with torch.cuda.device(self.device):
                self._side = [torch.cuda.Stream(device=self.device)
                              for _ in range(max(1, min(_PUSH_STREAMS, self.world - 1)))]
        return self._side

    def next_buffer(self) -> int:
        self._buf ^= 1
        return self._buf

    def epoch(self) -> Tuple[int, int]:
        """(collectives completed, error word) -- synchronises the current stream."""
        e, err = ctypes.c_uint32(0), ctypes.c_uint32(0)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().mk_peer_epoch(self.local, ctypes.byref(e), ctypes.byref(err),
                                                torch.cuda.current_stream().cuda_stream), "mk_peer_epoch")
        return int(e.value), int(err.value)

    def close(self) -> None:
        if self._symm is not None:     # symmetric memory: torch owns it
            self._bytes = None
            self._symm = None
            self.local = None
            self.mc = 0
            return
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
        multicast(group)          # one collective probe per process: symmetric memory + multicast or CUDA IPC
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


def begin_push(win: PeerWindow, buf: int) -> None:
    """Before the producer overwrites table buffer `buf`: wait until every peer has released it."""
    global _launches
    with torch.cuda.device(win.device):
        rc = _lib.lib().mk_peer_begin_push(win.local, win.world, win.rank, int(buf), _TIMEOUT_MS, _stream())
    _lib.check(rc, "mk_peer_begin_push")
    _launches += 1


def publish_and_push(win: PeerWindow, buf: int, offsets: Sequence[int], bytes_per_rank: Sequence[int]) -> None:
    """The rank's rows of the segments at `offsets` (segment g: `bytes_per_rank[g]` bytes per rank) are
    complete in its own window (current stream): open the collective and let the copy engines carry
    the rows to every peer on the window's side stream, each peer's rows followed by its flag."""
    global _launches
    n = len(offsets)
    offs = (ctypes.c_int64 * n)(*[int(o) for o in offsets])
    nbytes = (ctypes.c_int64 * n)(*[int(b) for b in bytes_per_rank])
    L = _lib.lib()
    main = torch.cuda.current_stream()
    sides = win.side_streams()
    with torch.cuda.device(win.device):
        _lib.check(L.mk_peer_publish(win.local, win.rank, int(buf), main.cuda_stream), "mk_peer_publish")
        ev = torch.cuda.Event()
        ev.record(main)
        win._pushed = []
        for i, side in enumerate(sides):
            side.wait_event(ev)
            _lib.check(L.mk_peer_push_steps(win.ptrs, win.world, win.rank, n, offs, nbytes, 1 + i, len(sides),
                                            side.cuda_stream), "mk_peer_push_steps")
            done = torch.cuda.Event()
            done.record(side)
            win._pushed.append(done)
    _launches += 1


def join_push(win: PeerWindow) -> None:
    """The current stream continues only after the window's pushes have left (joins the side stream;
    needed before the next producer touches the window, and before a CUDA graph capture ends)."""
    if win._pushed:
        cur = torch.cuda.current_stream()
        for ev in win._pushed:
            cur.wait_event(ev)
        win._pushed = None


def wait_all(win: PeerWindow) -> None:
    """For consumers that cannot wait per block: returns (on the stream) when the whole table is in."""
    global _launches
    with torch.cuda.device(win.device):
        rc = _lib.lib().mk_peer_wait_all(win.local, win.world, _TIMEOUT_MS, _stream())
    _lib.check(rc, "mk_peer_wait_all")
    _launches += 1


def release(win: PeerWindow) -> None:
    """The rank has finished reading the table of its current collective (tells the peers)."""
    global _launches
    with torch.cuda.device(win.device):
        rc = _lib.lib().mk_peer_release(win.ptrs, win.world, win.rank, _stream())
    _lib.check(rc, "mk_peer_release")
    _launches += 1


def check_errors() -> None:
    """Raise if a kernel gave up waiting for a peer on any window (synchronises the device)."""
    for key, w in _windows.items():
        ep, err = w.epoch()
        if err:
            raise PeerTimeoutError(f"peer window {key[0]}: a kernel of rank {w.rank} gave up waiting for a peer "
                                   f"in collective {err} (epoch {ep}); results since then are invalid")


def reduce_scatter(win: PeerWindow, offset: int, rows: int, k: int, grid: int = 0) -> torch.Tensor:
    """Sum over ranks of rows [rank*rows, (rank+1)*rows) of the fp32 [world*rows, k] buffer at
    `offset` of every rank's window -> fp32 [rows, k]."""
    global _launches
    block = rows * k * 4
    if block % 16:
        raise RuntimeError("peer reduce-scatter blocks must be multiples of 16 bytes")
    out = torch.empty((rows, k), dtype=torch.float32, device=win.device)
    with torch.cuda.device(win.device):
        if win.mc and _MULTICAST and (win.world > 2 or _PUSH == "mc"):   # summed by the switch: one reduced load per element
            # (two ranks: one remote load per element either way, and the plain loads measured 3 % faster)
            rc = _lib.lib().mk_peer_reduce_scatter_mc(win.ptrs, win.mc, win.world, win.rank, int(offset), block,
                                                      out.data_ptr(), grid, _TIMEOUT_MS, _stream())
        else:
            rc = _lib.lib().mk_peer_reduce_scatter(win.ptrs, win.world, win.rank, int(offset), block,
                                                   out.data_ptr(), grid, _TIMEOUT_MS, _stream())
    _lib.check(rc, "mk_peer_reduce_scatter")
    _launches += 1
    return out


def reduce_scatter_virtual(wins: Sequence[PeerWindow], offset: int, rows: int, k: int, grid: int = 0):
    """Single-process emulation (PeerWindow.create_virtual): the reduce-scatter of ALL virtual ranks
    in one launch -- launches that wait on one another must not share a device.  Tests only."""
    global _launches
    world = len(wins)
    block = rows * k * 4
    outs = [torch.empty((rows, k), dtype=torch.float32, device=wins[0].device) for _ in range(world)]
    optr = (ctypes.c_void_p * world)(*[o.data_ptr() for o in outs])
    with torch.cuda.device(wins[0].device):
        rc = _lib.lib().mk_peer_reduce_scatter_virtual(wins[0].ptrs, world, int(offset), block, optr, grid,
                                                       _TIMEOUT_MS, _stream())
    _lib.check(rc, "mk_peer_reduce_scatter_virtual")
    _launches += 1
    return outs
