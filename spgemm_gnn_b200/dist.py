"""1-D node/row partition of the aggregation over the GPUs of one box.

The reference is single-GPU (README_INTEGRATED.md:382 lists multi-GPU as future work); this
module is the B200 answer BASELINE.json asks for.  One process per GPU, `torch.distributed`
(NCCL over NVLink 5 / NVSwitch; gloo on CPU for the tests):

  * rank p owns the R = ceil(N / P) consecutive rows [p*R, (p+1)*R) of the adjacency (local CSR
    with GLOBAL column ids) and those nodes' features; equal row counts keep every collective on
    NCCL's equal-size fast path, and a random node order (the synthetic shapes are random;
    `random_relabel` does it for others) makes equal rows also nnz-balanced;
  * forward: MaxK -> CBSR is local; ONE all-gather moves only the compact CBSR table
    (N*k*(4+w) bytes: 12.8x smaller than dense at k=32, D=256); the SpGEMM then runs locally
    against the gathered table;
  * backward: the push-form SSpMM produces contributions to every node's CBSR gradient, folded
    by ONE reduce-scatter of N*k*4 bytes (the column ids are already resident from the forward);
  * weights are replicated, their gradients all-reduced in one flat bucket.

The two exchanges have two forms: the `torch.distributed` calls in this file (gloo on CPU, NCCL
with MAXK_PEER_EXCHANGE=0) and the library's own NVLink kernels over CUDA-IPC windows (peer.py,
csrc/peer.cu, the PUSH form of csrc/bank.cu), which NCCL groups use by default.
"""
from __future__ import annotations

from typing import Iterable, Tuple

import torch
import torch.distributed as dist
from torch.autograd import Function

from .graph import CSRGraph, ceil_div


def rows_per_rank(num_nodes: int, world: int) -> int:
    """Rows of the padded table every rank owns: ceil(N / P) rounded up to a multiple of 16, so that a
    rank's block of every table segment (k bytes per row at the narrowest) is a whole number of
    16-byte units -- what the NVLink stores of the exchange move."""
    return ceil_div(ceil_div(num_nodes, world), 16) * 16


def shard_bounds(g: CSRGraph, world: int, balance: str = "rows") -> list:
    """world+1 row offsets of the 1-D partition.  "rows": equal row counts (fine for node orders that
    are random, as the synthetic shapes are).  "nnz": contiguous row ranges with equal numbers of
    stored entries ("balance by nnz, not rows, for skewed graphs") -- no
    relabelling, so whatever locality the node order has is kept."""
    n = g.num_nodes()
    if balance == "rows":
        r = ceil_div(n, world)
        return [min(p * r, n) for p in range(world + 1)]
    if balance == "nnz":
        from .graph import row_partition_bounds
        return row_partition_bounds(g.indptr, world)
    raise ValueError(f"unknown shard balance {balance!r}")


def shard_graph(g: CSRGraph, rank: int, world: int, bounds=None) -> Tuple[CSRGraph, int, int]:
    """Rank's row block of `g`, padded to R rows with empty rows, R = the largest block rounded up to
    a multiple of 16.  The column space is world*R wide: node j of block q sits at table row
    q*R + (j - bounds[q]), so that gathered tables index directly by column id.  With equal-row
    blocks of width R that IS the global id; otherwise the ids are remapped here, once."""
    n = g.num_nodes()
    if bounds is None:
        r = rows_per_rank(n, world)
        bounds = [min(p * r, n) for p in range(world + 1)]
    else:
        r = ceil_div(max(bounds[p + 1] - bounds[p] for p in range(world)), 16) * 16
    r0, r1 = bounds[rank], bounds[rank + 1]
    s = g.row_slice(r0, r1)
    ptr = s.indptr
    if r1 - r0 < r:  # pad with empty rows
        pad = ptr[-1:].expand(r - (r1 - r0))
        ptr = torch.cat([ptr, pad]).contiguous()
    cols = s.indices
    if any(bounds[p] != p * r for p in range(world)):
        b = torch.tensor(bounds[:-1], dtype=torch.int64, device=cols.device)
        c64 = cols.to(torch.int64)
        q = torch.searchsorted(b, c64, right=True) - 1
        cols = (q * r + (c64 - b[q])).to(torch.int32)     # blocks are contiguous: rows stay ascending
    local = CSRGraph(ptr, cols.contiguous(), num_src=world * r)
    local._cache["symmetric"] = False
    return local, r0, r1


def shard_edge_weights(g: CSRGraph, local: CSRGraph, r0: int, r1: int, kind: str) -> torch.Tensor:
    """Per-edge weights of the rank's rows, computed from GLOBAL degrees."""
    lo, hi = int(g.indptr[r0]), int(g.indptr[r1])
    return g.edge_weights(kind)[lo:hi].contiguous()


def random_relabel(g: CSRGraph, seed: int = 97) -> Tuple[CSRGraph, torch.Tensor]:
    """Random node permutation (new = perm[old]) so that equal-row shards are nnz-balanced."""
    from .graph import from_edges
    n = g.num_nodes()
    gen = torch.Generator(device=g.device).manual_seed(seed)
    perm = torch.randperm(n, generator=gen, device=g.device)
    rows = perm[g.row_ids()]
    cols = perm[g.indices.to(torch.int64)]
    return from_edges(rows, cols, n, symmetric=g._cache.get("symmetric", False)), perm


# ---------------------------------------------------------------------------------------
# the two exchanges
# ---------------------------------------------------------------------------------------
def _backend(group) -> str:
    return dist.get_backend(group)


def allgather_cbsr(sp_data: torch.Tensor, sp_index: torch.Tensor, group=None):
    """[R,k] local CBSR -> [P*R,k] table, rows ordered by owner rank == global node id."""
    world = dist.get_world_size(group)
    r, k = sp_data.shape
    full_data = torch.empty((world * r, k), dtype=sp_data.dtype, device=sp_data.device)
    full_index = torch.empty((world * r, k), dtype=sp_index.dtype, device=sp_index.device)
    dist.all_gather_into_tensor(full_data, sp_data.contiguous(), group=group)
    # column ids travel as raw bytes (uint16 has no NCCL dtype in every torch build)
    dist.all_gather_into_tensor(full_index.view(torch.uint8), sp_index.contiguous().view(torch.uint8),
                                group=group)
    return full_data, full_index


def allgather_many(locals_, group=None):
    """Several [R, ...] tensors -> their [P*R, ...] gathers in ONE NCCL launch (grouped call);
    falls back to one call per tensor where the coalescing manager is not available (gloo)."""
    world = dist.get_world_size(group)
    fulls = [torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
             for t in locals_]
    pairs = [(f.view(torch.uint8), t.contiguous().view(torch.uint8)) for f, t in zip(fulls, locals_)]
    done = False
    if _backend(group) == "nccl" and hasattr(dist, "_coalescing_manager"):
        try:
            with dist._coalescing_manager(group=group, device=locals_[0].device, async_ops=False):
                for f, t in pairs:
                    dist.all_gather_into_tensor(f, t, group=group)
            done = True
        except Exception:
            done = False
    if not done:
        for f, t in pairs:
            dist.all_gather_into_tensor(f, t, group=group)
    return fulls


def allgather_rows(local: torch.Tensor, group=None) -> torch.Tensor:
    """[R,k] of any dtype -> [P*R,k] (moved as raw bytes)."""
    world = dist.get_world_size(group)
    full = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype,
                       device=local.device)
    dist.all_gather_into_tensor(full.view(torch.uint8), local.contiguous().view(torch.uint8), group=group)
    return full


def reduce_scatter_rows(full: torch.Tensor, group=None) -> torch.Tensor:
    """[P*R,k] per-rank partial sums -> [R,k] summed rows of this rank."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    r = full.shape[0] // world
    if _backend(group) == "gloo":  # gloo has no reduce_scatter: all-reduce and slice
        buf = full.clone()
        dist.all_reduce(buf, group=group)
        return buf[rank * r:(rank + 1) * r].contiguous()
    out = torch.empty((r, full.shape[1]), dtype=full.dtype, device=full.device)
    dist.reduce_scatter_tensor(out, full.contiguous(), group=group)
    return out


def allreduce_grads(params: Iterable[torch.nn.Parameter], group=None) -> None:
    """Sum the (small, replicated) weight gradients in one flat bucket."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, group=group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


# ---------------------------------------------------------------------------------------
# the sharded hot path as an autograd Function
# ---------------------------------------------------------------------------------------
def _peer_path(group, *row_bytes) -> bool:
    """Peer-memory exchange (peer.py) instead of NCCL calls: not switched off (MAXK_PEER_EXCHANGE=0),
    NCCL group on one box, and every listed per-rank segment a multiple of 16 bytes (same answer on
    every rank: same shapes)."""
    from . import peer
    return peer.enabled() and peer.available(group) and all(b % 16 == 0 for b in row_bytes)


def sharded_forward(sp_data, sp_index, ptr, idx, val, num_rows, dim_origin, group=None,
                    keep_index: bool = True):
    """Exchange + local forward SpGEMM of one rank.  Returns (out [num_rows, D], gathered sorted
    column ids [P*R, k] for the backward).  Where the banked kernels apply, the LOCAL rows are
    banked first and the banked values + cell offsets travel next to the sorted column ids
    (7 bytes per entry instead of 5), so that no rank re-banks rows it does not own.

    On an NCCL group the exchange runs over peer-mapped windows (default; `MAXK_PEER_EXCHANGE=0`
    keeps the NCCL calls below) and OVERLAPS the SpGEMM: the rank's rows are produced straight
    into its own window, the copy engines carry them to the peers on a side stream
    (`peer.publish_and_push`), and the SpGEMM starts at once -- it walks every CSR row in the order
    the source blocks arrive (own rows, rank+1, rank+2, ...) and checks the sender's flag before it
    touches a block.  The window holds two table buffers (alternating) and is shared by all layers
    of a shape, so the column ids are copied out of it when a backward will need them."""
    from . import maxk_kernels, peer
    r, k = sp_data.shape
    ib = sp_index.element_size()
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    part = maxk_kernels.partition(ptr, num_rows)
    e = idx.numel()
    form = ("banked" if maxk_kernels.use_banked(part.num_parts, e, k, dim_origin) else
            "packed" if maxk_kernels.use_packed(part.num_parts, e, k, dim_origin) else "plain")
    # every record walks the source blocks own, rank+1, ..., rank-1: one fixed summation order for the
    # peer form and the NCCL form (their forwards are bit-equal)
    split = (maxk_kernels.block_split(ptr, idx, num_rows, world, rank, r)
             if (form != "plain" and sp_data.is_cuda) else None)
    # banked form: one launch per source-block phase (own block, the next senders, the rest), each adding
    # to the rows of the earlier ones -- same launches, same sums, in the peer form and the NCCL form
    phases = blk = None
    if form == "banked" and sp_data.is_cuda and world > 1 and peer.phases():
        blk = maxk_kernels.block_pointers(ptr, idx, num_rows, world, r)
        if blk is not None:
            phases = maxk_kernels.forward_phases(world, rank)
    pkw = dict(phases=phases, blk=blk, n_blocks=world) if phases is not None else dict(split=split)
    if _peer_path(group, r * k * 4) and peer.wanted(world, world * r * k * (6 + ib if form == "banked" else 4 + ib),
                                                   group):
        rows = world * r
        per_rank = {"banked": [r * k * 4, r * k * 2, r * k * ib], "packed": [r * k * 8, r * k * ib],
                    "plain": [r * k * 4, r * k * ib]}[form]
        offs, total = peer.layout([world * b for b in per_rank] * 2)        # two table buffers
        win = peer.window("table_" + form, total, group)
        if win is not None:
            buf = win.next_buffer()
            o = offs[len(per_rank) * buf: len(per_rank) * (buf + 1)]
            mine = slice(rank * r, (rank + 1) * r)
            # who moves the rows: pusher CTAs inside the forward kernel ("sm": needs 16-byte multiples),
            # or the copy engines on side streams ("dma"), with or without per-block waiting
            aligned = all(b % 16 == 0 for b in per_rank)
            mcast = aligned and world > 1 and peer.use_multicast(win)          # every row stored once (NVLS)
            fused = (not mcast and form != "plain" and peer.push_mode() == "sm" and aligned)
            wait = None
            if mcast:
                pass
            elif fused:
                wait = peer.exchange(win, r, o, per_rank)
            elif peer.overlap() and form != "plain" and peer.push_mode() == "dma":
                wait = peer.exchange(win, r)
            peer.begin_push(win, buf)
            full_index = win.view(o[-1], (rows, k), sp_index.dtype)
            full_index[mine].copy_(sp_index)
            if form == "banked":
                full_data = win.view(o[0], (rows, k), torch.float32)
                full_slot = win.view(o[1], (rows, k), torch.int16)
                maxk_kernels.cbsr_bank(sp_data, sp_index, dim_origin, with_index=False,
                                       out=(full_data[mine], full_slot[mine]))
            elif form == "packed":
                full_pack = win.view(o[0], (rows, k, 2), torch.int32)
                maxk_kernels.cbsr_bank_packed(sp_data, sp_index, dim_origin, out=full_pack[mine])
            else:
                full_data = win.view(o[0], (rows, k), torch.float32)
                full_data[mine].copy_(sp_data)
            if mcast:
                peer.publish(win, buf)
                peer.push_mc(win, o, per_rank)
                peer.wait_all(win)
            elif fused:
                peer.publish(win, buf)
            elif peer.push_mode() in ("sm", "sm_seq", "auto", "mc") and aligned:
                peer.publish(win, buf)           # NVLink stores to every peer as a kernel of its own
                peer.push_sm(win, o, per_rank)
                peer.wait_all(win)
            else:
                peer.publish_and_push(win, buf, o, per_rank)
                if wait is None:
                    peer.wait_all(win)
            if form == "banked":
                if phases is not None and wait is not None:   # only the first launch carries the pushers
                    rest = peer.exchange(win, r)
                    wait = [wait] + [rest] * (len(phases) - 1)
                out = maxk_kernels.spgemm_forward_banked(ptr, idx, val, full_data, full_slot, num_rows, e, k,
                                                         dim_origin, wait=wait, **pkw)
            elif form == "packed":
                out = maxk_kernels.spgemm_forward_packed(ptr, idx, val, full_pack, num_rows, e, k, dim_origin,
                                                         split=split, wait=wait)
            else:
                out, _ = maxk_kernels.spgemm_forward(ptr, idx, val, full_data, full_index, num_rows, e, k,
                                                     dim_origin, allow_banked=False)
            peer.join_push(win)
            kept = full_index.clone() if keep_index else full_index
            peer.release(win)
            return out, kept
    if form == "banked":
        bk_data, _, bk_slot = maxk_kernels.cbsr_bank(sp_data, sp_index, dim_origin, with_index=False)
        full_data, full_slot, full_index = allgather_many([bk_data, bk_slot, sp_index], group)
        out = maxk_kernels.spgemm_forward_banked(ptr, idx, val, full_data, full_slot, num_rows, e, k,
                                                 dim_origin, **pkw)
    elif form == "packed":
        bk_pack = maxk_kernels.cbsr_bank_packed(sp_data, sp_index, dim_origin)
        full_pack, full_index = allgather_many([bk_pack, sp_index], group)
        out = maxk_kernels.spgemm_forward_packed(ptr, idx, val, full_pack, num_rows, e, k, dim_origin, split=split)
    else:
        full_data, full_index = allgather_many([sp_data, sp_index], group)
        out, _ = maxk_kernels.spgemm_forward(ptr, idx, val, full_data, full_index, num_rows, e, k,
                                             dim_origin, allow_banked=False)
    return out, full_index


def sharded_backward(grad_out, full_index, ptr, idx, val, num_rows, dim_origin, group=None):
    """Local push-form SSpMM into a full-height buffer, folded by one reduce-scatter: loads from
    every rank's peer window in fixed rank order (`peer.reduce_scatter`), or NCCL's
    (`MAXK_PEER_EXCHANGE=0`, gloo, or windows that could not be mapped)."""
    from . import maxk_kernels, peer
    n_src, k = full_index.shape
    world = dist.get_world_size(group)
    r = n_src // world
    if _peer_path(group, r * k * 4) and peer.wanted(world, n_src * k * 4, group):
        offs, total = peer.layout([n_src * k * 4])
        win = peer.window("dxs", total, group)
        if win is not None:
            dxs_full = win.view(offs[0], (n_src, k), torch.float32)
            maxk_kernels.spgemm_backward(ptr, idx, val, grad_out, full_index, num_rows, idx.numel(), k,
                                         dim_origin, out=dxs_full)
            return peer.reduce_scatter(win, offs[0], r, k)
    dxs_full = maxk_kernels.spgemm_backward(ptr, idx, val, grad_out, full_index, num_rows,
                                            idx.numel(), k, dim_origin)
    return reduce_scatter_rows(dxs_full, group)


class DistSpGEMMFunction(Function):
    """Local rows of A x Xs where Xs is row-sharded: all-gather(CBSR) -> spgemm_forward;
    backward: spgemm_backward against the gathered column ids -> reduce-scatter."""

    @staticmethod
    def forward(ctx, sp_data, sp_index, ptr, idx, val, num_rows, dim_origin, group):
        out, full_index = sharded_forward(sp_data.contiguous(), sp_index, ptr, idx, val, num_rows,
                                          dim_origin, group, keep_index=ctx.needs_input_grad[0])
        k = sp_data.shape[1]
        ctx.save_for_backward(full_index, ptr, idx, val)
        ctx.meta = (num_rows, k, dim_origin, group)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        full_index, ptr, idx, val = ctx.saved_tensors
        num_rows, k, dim_origin, group = ctx.meta
        dxs = sharded_backward(grad_out.contiguous(), full_index, ptr, idx, val, num_rows,
                               dim_origin, group)
        return dxs, None, None, None, None, None, None, None


class ShardedGraph(CSRGraph):
    """One rank's row block of a graph, usable wherever the layers and models take a `CSRGraph`:
    the aggregation then goes through `DistSpGEMMFunction` (all-gather / reduce-scatter).
    Built from the FULL graph so that the per-edge weights use global degrees."""

    def __init__(self, full: CSRGraph, rank: int, world: int, group=None, balance: str = "rows"):
        self.balance = balance
        self.bounds = None if balance == "rows" else shard_bounds(full, world, balance)
        local, r0, r1 = shard_graph(full, rank, world, self.bounds)
        super().__init__(local.indptr, local.indices, local.num_src)
        self.rank, self.world, self.group = rank, world, group
        self.row_begin, self.row_end = r0, r1
        self.global_nodes = full.num_nodes()
        self.rows_per_rank = local.num_nodes()
        self._weights = {k: shard_edge_weights(full, local, r0, r1, k) for k in ("mean", "both", "sum")}
        self._cache["symmetric"] = False
        # padded rows are empty by construction; only real rows count for GraphConv's check
        self._cache["has_zero_in"] = bool((self.in_degrees()[: r1 - r0] == 0).any())

    def edge_weights(self, kind: str) -> torch.Tensor:
        kind = {"right": "mean", "none": "sum"}.get(kind, kind)
        return self._weights[kind]

    def to(self, device) -> "ShardedGraph":
        """A shard is bound to its rank's device and process group: moving it would silently drop the
        sharding (CSRGraph.to returns a plain graph whose global column ids no longer match a local
        table).  Same device: no-op; anything else is an error."""
        if torch.device(device) == self.device or (torch.device(device).type == self.device.type
                                                   and torch.device(device).index is None):
            return self
        raise RuntimeError("a ShardedGraph cannot be moved: build it from the full graph on the target device")

    def row_slice(self, r0: int, r1: int):
        raise RuntimeError("row_slice of a ShardedGraph is not defined: slice the full graph, then shard it")

    def local_rows(self, t: torch.Tensor) -> torch.Tensor:
        """Rows [row_begin, row_end) of a full-height tensor, zero-padded to rows_per_rank."""
        out = t.new_zeros((self.rows_per_rank,) + tuple(t.shape[1:]))
        out[: self.row_end - self.row_begin] = t[self.row_begin:self.row_end]
        return out


def dist_maxk_aggregate(local: CSRGraph, val: torch.Tensor, feat: torch.Tensor, k: int,
                        group=None) -> torch.Tensor:
    """Sharded MaxK -> CBSR -> all-gather -> SpGEMM for the rank's (padded) rows."""
    from .maxk_layers import MaxKCBSRFunction
    sp_data, sp_index = MaxKCBSRFunction.apply(feat, k)
    return DistSpGEMMFunction.apply(sp_data, sp_index, local.indptr, local.indices, val,
                                    local.num_nodes(), feat.shape[1], group)


# ---------------------------------------------------------------------------------------
# shards on disk (f-4: the `.warp4` successor for the row-partitioned path)
# ---------------------------------------------------------------------------------------
SHARD_MAGIC = "maxk-b200-shard-v1"


def save_shard(sg: "ShardedGraph", path: str) -> None:
    """One rank's shard: local CSR (table-row column ids), per-edge weights from the global degrees,
    the partition bounds and, on a CUDA device, the work records of `mk_partition` -- everything a
    rank needs to start without ever holding the full graph."""
    import numpy as np
    from . import maxk_kernels
    blob = {"magic": np.array(SHARD_MAGIC), "indptr": sg.indptr.cpu().numpy(), "indices": sg.indices.cpu().numpy(),
            "num_src": np.int64(sg.num_src), "rank": np.int64(sg.rank), "world": np.int64(sg.world),
            "row_begin": np.int64(sg.row_begin), "row_end": np.int64(sg.row_end),
            "global_nodes": np.int64(sg.global_nodes), "balance": np.array(sg.balance),
            "bounds": np.array(sg.bounds if sg.bounds is not None else [], dtype=np.int64)}
    for kind, w in sg._weights.items():
        blob["w_" + kind] = w.cpu().numpy()
    if sg.indptr.is_cuda:
        part = maxk_kernels.partition(sg.indptr, sg.num_nodes())
        blob.update(parts=part.parts[: part.num_parts].cpu().numpy(), num_slots=np.int64(part.num_slots),
                    max_nz=np.int64(part.max_nz))
    np.savez(path, **blob)


def load_shard(path: str, device="cpu", group=None) -> "ShardedGraph":
    """Inverse of `save_shard` (the file of THIS rank); stored work records go into the partition cache."""
    import numpy as np
    z = np.load(path if path.endswith(".npz") else path + ".npz")
    if str(z["magic"]) != SHARD_MAGIC:
        raise ValueError(f"{path}: not a {SHARD_MAGIC} file")
    sg = ShardedGraph.__new__(ShardedGraph)
    CSRGraph.__init__(sg, torch.from_numpy(z["indptr"]).to(device), torch.from_numpy(z["indices"]).to(device),
                      int(z["num_src"]))
    sg.rank, sg.world, sg.group = int(z["rank"]), int(z["world"]), group
    sg.row_begin, sg.row_end = int(z["row_begin"]), int(z["row_end"])
    sg.global_nodes, sg.balance = int(z["global_nodes"]), str(z["balance"])
    sg.bounds = [int(v) for v in z["bounds"]] or None
    sg.rows_per_rank = sg.num_nodes()
    sg._weights = {k[2:]: torch.from_numpy(z[k]).to(device) for k in z.files if k.startswith("w_")}
    sg._cache["symmetric"] = False
    sg._cache["has_zero_in"] = bool((sg.in_degrees()[: sg.row_end - sg.row_begin] == 0).any())
    if "parts" in z.files and torch.device(device).type == "cuda":
        from . import maxk_kernels
        maxk_kernels.install_partition(sg.indptr, sg.num_nodes(), int(z["max_nz"]),
                                       torch.from_numpy(z["parts"]).to(device), int(z["num_slots"]))
    return sg
