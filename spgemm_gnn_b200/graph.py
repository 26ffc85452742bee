"""Destination-indexed CSR graph container and synthetic graph shapes.

The reference keeps its graphs in DGL (`maxk_gnn_dgl.py:224-237`) and pulls CSR out of
it inside every layer call (`utils/maxk_layers.py:103-159`, Python loops with `.item()`
per node).  DGL does not exist in this image, so the host side of the hot path is a
plain CSR container that exposes the handful of DGLGraph methods the reference layers
touch (`num_nodes`, `num_edges`, `in_degrees`, `out_degrees`, `adj_tensors`,
`local_scope`, `device`, `to`) and that builds the per-edge weights once per graph
with vectorised torch ops.

Orientation: `graph.update_all(copy_u, sum)` reduces over
IN-edges, so row i of the CSR holds the sources j of the edges j -> i.  That is DGL's
`adj_tensors('csc')`; the reference asks for `'csr'` (source-indexed) which only
coincides on symmetric graphs.  All synthetic shapes here are symmetrised, so both
agree, and the container is explicit about which one it stores.
"""
from __future__ import annotations

import contextlib
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import torch

# (nodes, directed edges incl. both directions) of the real datasets the synthetic
# shapes imitate: spgemm_plot.py:7-12 and run/reddit.log:28 in the reference.
SHAPES: Dict[str, Tuple[int, int]] = {
    "flickr": (89_250, 899_756),
    "reddit": (232_965, 114_615_891),
    "yelp": (716_847, 13_954_819),
    "ogbn-proteins": (132_534, 79_122_504),
    "ogbn-products": (2_449_029, 123_718_280),
}
# in_feats / classes logged by the reference
FEATS: Dict[str, Tuple[int, int]] = {
    "flickr": (500, 7),
    "reddit": (602, 41),
    "yelp": (300, 100),
    "ogbn-proteins": (8, 112),
    "ogbn-products": (100, 47),
}
MAX_DEGREE = 32768


@dataclass
class CSRGraph:
    """Destination-indexed CSR: row i lists the in-neighbours of node i.

    `indptr` int32 [N+1], `indices` int32 [E] (ascending inside a row).  Stands in for
    the `DGLGraph` argument of the reference layers (`utils/maxk_layers.py:82`).
    """

    indptr: torch.Tensor
    indices: torch.Tensor
    num_src: Optional[int] = None  # number of columns; None -> square
    _cache: dict = field(default_factory=dict, repr=False)

    def __post_init__(self):
        if self.indptr.dtype != torch.int32 or self.indices.dtype != torch.int32:
            raise TypeError("CSRGraph wants int32 indptr/indices")
        if self.indptr.dim() != 1 or self.indices.dim() != 1:
            raise ValueError("indptr and indices must be 1-D")
        if self.num_src is None:
            self.num_src = self.num_nodes()
        # the reference layers look for this attribute to choose the kernel path
        # (utils/maxk_layers.py:94, set by maxk_gnn_integrated.py:77-135)
        self._sparse_format = {"ptr": self.indptr, "idx": self.indices}

    # -- the DGLGraph surface the reference layers use ---------------------------------
    def num_nodes(self) -> int:
        return self.indptr.numel() - 1

    number_of_nodes = num_nodes

    def num_edges(self) -> int:
        return self.indices.numel()

    number_of_edges = num_edges

    @property
    def device(self) -> torch.device:
        return self.indptr.device

    def in_degrees(self) -> torch.Tensor:
        d = self._cache.get("in_deg")
        if d is None:
            d = (self.indptr[1:] - self.indptr[:-1]).to(torch.int64)
            self._cache["in_deg"] = d
        return d

    def out_degrees(self) -> torch.Tensor:
        d = self._cache.get("out_deg")
        if d is None:
            d = torch.bincount(self.indices.to(torch.int64), minlength=self.num_src)
            self._cache["out_deg"] = d
        return d

    def adj_tensors(self, fmt: str = "csc"):
        """`'csc'` is what is stored.  `'csr'` (source-indexed) is only handed out when
        the graph was built symmetric, where the two coincide."""
        if fmt == "csc" or (fmt == "csr" and self._cache.get("symmetric", False)):
            return self.indptr, self.indices, torch.arange(self.num_edges(), device=self.device)
        raise ValueError("source-indexed CSR requested from a graph not known to be symmetric")

    @contextlib.contextmanager
    def local_scope(self):
        yield self

    def to(self, device) -> "CSRGraph":
        g = CSRGraph(self.indptr.to(device), self.indices.to(device), self.num_src)
        g._cache["symmetric"] = self._cache.get("symmetric", False)
        return g

    def int(self) -> "CSRGraph":
        return self

    # -- per-edge weights --------------------------------
    def row_ids(self) -> torch.Tensor:
        r = self._cache.get("row_ids")
        if r is None:
            n = self.num_nodes()
            r = torch.repeat_interleave(
                torch.arange(n, device=self.device, dtype=torch.int64), self.in_degrees()
            )
            self._cache["row_ids"] = r
        return r

    def edge_weights(self, kind: str) -> torch.Tensor:
        """fp32 [E] weights.

        * `'mean'`: 1/max(deg_in(i),1) on every edge of row i (SAGE mean,
          `utils/maxk_layers.py:147-157`, there an O(N) `.item()` loop).
        * `'both'`: deg_out(j)^-1/2 * deg_in(i)^-1/2, degrees clamped to 1 (DGL
          `GraphConv(norm='both')`, used at `utils/models.py:252`).
        * `'right'`: deg_in(i)^-1 -- same numbers as `'mean'`.
        * `'sum'` / `'none'`: 1.0 (GIN, `utils/models.py:375`).
        * `'reference_gcn'`: deg_out(j)^-1/2 * deg_in(j)^-1/2 -- BOTH normalisations on the source node,
          which is what the reference's own `MaxKGCNConv` computes (`utils/maxk_layers.py:315-318`
          scales the features by deg_out^-1/2, `:372-376` weights every edge with `norm_right[idx]` =
          deg_in^-1/2 of the SOURCE), unlike the `GraphConv(norm='both')` it trains with.
        """
        w = self._cache.get(("w", kind))
        if w is not None:
            return w
        e = self.num_edges()
        if kind in ("sum", "none"):
            w = torch.ones(e, dtype=torch.float32, device=self.device)
        elif kind in ("mean", "right"):
            inv = 1.0 / self.in_degrees().clamp(min=1).to(torch.float32)
            w = inv[self.row_ids()]
        elif kind == "both":
            di = self.in_degrees().clamp(min=1).to(torch.float32).pow(-0.5)
            do = self.out_degrees().clamp(min=1).to(torch.float32).pow(-0.5)
            w = di[self.row_ids()] * do[self.indices.to(torch.int64)]
        elif kind == "reference_gcn":
            di = self.in_degrees().clamp(min=1).to(torch.float32).pow(-0.5)
            do = self.out_degrees().clamp(min=1).to(torch.float32).pow(-0.5)
            w = (di * do)[self.indices.to(torch.int64)]
        else:
            raise ValueError(f"unknown edge weight kind {kind!r}")
        w = w.contiguous()
        self._cache[("w", kind)] = w
        return w

    def row_slice(self, r0: int, r1: int) -> "CSRGraph":
        """Rows [r0, r1) with GLOBAL column ids: the shard one rank owns under the 1-D
        row partition."""
        lo = int(self.indptr[r0])
        hi = int(self.indptr[r1])
        ptr = (self.indptr[r0 : r1 + 1] - lo).contiguous()
        return CSRGraph(ptr, self.indices[lo:hi].contiguous(), self.num_src)


FILE_MAGIC = "maxk-b200-graph-v1"


def save_graph(g: CSRGraph, path: str, max_nz: int = None) -> None:
    """On-disk form of a graph for the hot path: CSR plus, when a CUDA device is present, the work
    records of `mk_partition` -- the successor of the reference's `<graph>.warp4` file (read by
    `cuda_read_array<int>` on EVERY kernel call there, so@0x24d16; read once here)."""
    import numpy as np
    blob = {"magic": np.array(FILE_MAGIC), "indptr": g.indptr.cpu().numpy(),
            "indices": g.indices.cpu().numpy(), "num_src": np.int64(g.num_src),
            "symmetric": np.bool_(g._cache.get("symmetric", False))}
    if g.indptr.is_cuda:
        from . import maxk_kernels
        part = maxk_kernels.partition(g.indptr, g.num_nodes(), max_nz)
        blob.update(parts=part.parts[: part.num_parts].cpu().numpy(), num_slots=np.int64(part.num_slots),
                    max_nz=np.int64(part.max_nz))
    np.savez(path, **blob)


def load_graph(path: str, device="cpu") -> CSRGraph:
    """Inverse of `save_graph`; stored work records are installed in the partition cache so that
    the first kernel call does not rebuild them."""
    import numpy as np
    z = np.load(path if path.endswith(".npz") else path + ".npz")
    if str(z["magic"]) != FILE_MAGIC:
        raise ValueError(f"{path}: not a {FILE_MAGIC} file")
    g = CSRGraph(torch.from_numpy(z["indptr"]).to(device), torch.from_numpy(z["indices"]).to(device),
                 int(z["num_src"]))
    g._cache["symmetric"] = bool(z["symmetric"])
    if "parts" in z.files and torch.device(device).type == "cuda":
        from . import maxk_kernels
        maxk_kernels.install_partition(g.indptr, g.num_nodes(), int(z["max_nz"]),
                                       torch.from_numpy(z["parts"]).to(device), int(z["num_slots"]))
    return g


def from_edges(dst: torch.Tensor, src: torch.Tensor, num_nodes: int, *, symmetric=False) -> CSRGraph:
    """CSR from an edge list dst<-src.  Duplicate edges are merged, neighbours sorted."""
    key = dst.to(torch.int64) * num_nodes + src.to(torch.int64)
    key = torch.unique(key)  # sorted
    d = torch.div(key, num_nodes, rounding_mode="floor")
    s = (key - d * num_nodes).to(torch.int32)
    counts = torch.bincount(d, minlength=num_nodes)
    ptr = torch.zeros(num_nodes + 1, dtype=torch.int64, device=key.device)
    ptr[1:] = torch.cumsum(counts, 0)
    if int(ptr[-1]) >= 2**31:
        raise ValueError("edge count does not fit int32")
    g = CSRGraph(ptr.to(torch.int32), s.contiguous())
    g._cache["symmetric"] = symmetric
    return g


def from_dgl(g) -> CSRGraph:
    """A DGLGraph (anything with DGL's `edges()` / `num_nodes()` surface) as a `CSRGraph`, so that the
    layers take the object the reference's layers take (utils/maxk_layers.py:82 `forward(graph, feat)`;
    maxk_gnn_integrated.py:77-135 attaches `_sparse_format` the same way).  Rows are DESTINATIONS and
    list their in-neighbours ascending -- what `update_all(copy_u, sum|mean)` sums over, i.e. DGL's
    'csc' form; the reference extracts 'csr' (utils/maxk_layers.py:106), which is the same matrix
    only on the bidirected graphs it trains on.  Parallel edges are kept (DGL counts them).  The
    result is cached on the graph object."""
    if isinstance(g, CSRGraph):
        return g
    hit = getattr(g, "_maxk_csr", None)
    if hit is not None:
        return hit
    src, dst = g.edges()
    n = int(g.num_nodes())
    src, dst = torch.as_tensor(src), torch.as_tensor(dst)
    if src.numel() >= 2**31:
        raise ValueError("edge count does not fit int32")
    key, _ = torch.sort(dst.to(torch.int64) * n + src.to(torch.int64))
    d = torch.div(key, n, rounding_mode="floor")
    ptr = torch.zeros(n + 1, dtype=torch.int64, device=key.device)
    ptr[1:] = torch.cumsum(torch.bincount(d, minlength=n), 0)
    out = CSRGraph(ptr.to(torch.int32), (key - d * n).to(torch.int32).contiguous())
    try:
        g._maxk_csr = out
    except AttributeError:
        pass
    return out


def synthetic_graph(
    num_nodes: int,
    num_edges: int,
    *,
    seed: int = 97,
    sigma: float = 1.0,
    device="cpu",
    self_loops: bool = True,
) -> CSRGraph:
    """Symmetric random graph of a given shape.

    Half-degrees are lognormal(sigma) rescaled so that the symmetrised edge count lands
    near `num_edges`, clipped to [1, min(N-1, 32768)]; endpoints uniform; duplicates
    merged; both directions kept; one self-loop per node (the reference adds them with
    `AddSelfLoop`, maxk_gnn_dgl.py:221-223).  Seed 97 is the reference default
    (utils/config.py:54).
    """
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    n = num_nodes
    target_half = max(num_edges // 2, 1)
    z = torch.randn(n, generator=gen, device=dev, dtype=torch.float64) * sigma
    w = torch.exp(z)
    cap = float(min(n - 1, MAX_DEGREE // 2)) if n > 1 else 1.0
    scale = target_half / float(w.sum())
    # clipping the tail loses mass; two fixed-point passes put it back
    for _ in range(3):
        deg = torch.clamp(torch.round(w * scale), 1.0, cap)
        scale *= target_half / float(deg.sum())
    deg = torch.clamp(torch.round(w * scale), 1.0, cap).to(torch.int64)
    src = torch.repeat_interleave(torch.arange(n, device=dev), deg)
    dst = torch.randint(0, n, (src.numel(),), generator=gen, device=dev)
    rows = torch.cat([src, dst])
    cols = torch.cat([dst, src])
    if self_loops:
        loop = torch.arange(n, device=dev)
        rows = torch.cat([rows, loop])
        cols = torch.cat([cols, loop])
    else:
        keep = rows != cols
        rows, cols = rows[keep], cols[keep]
    return from_edges(rows, cols, n, symmetric=True)


def shaped_graph(name: str, *, scale: float = 1.0, seed: int = 97, device="cpu") -> CSRGraph:
    """One of the BASELINE.json shapes, optionally scaled down (nodes and edges by the
    same factor, so the average degree is kept)."""
    n, e = SHAPES[name]
    n = max(int(round(n * scale)), 4)
    e = max(int(round(e * scale)), n)
    return synthetic_graph(n, e, seed=seed, device=device)


def permute(g: CSRGraph, perm: torch.Tensor) -> CSRGraph:
    """The same graph under the node relabelling new = perm[old] (A' = P A P^T).  Parallel edges are
    kept, neighbours come out ascending.  Node data moves with `x_new[perm] = x_old`."""
    n = g.num_nodes()
    if g.num_src not in (None, n):
        raise ValueError("permute wants a square graph")
    perm = perm.to(device=g.device, dtype=torch.int64)
    if perm.numel() != n:
        raise ValueError("perm must have one entry per node")
    key, _ = torch.sort(perm[g.row_ids()] * n + perm[g.indices.to(torch.int64)])
    d = torch.div(key, n, rounding_mode="floor")
    ptr = torch.zeros(n + 1, dtype=torch.int64, device=key.device)
    ptr[1:] = torch.cumsum(torch.bincount(d, minlength=n), 0)
    out = CSRGraph(ptr.to(torch.int32), (key - d * n).to(torch.int32).contiguous())
    out._cache["symmetric"] = g._cache.get("symmetric", False)
    return out


def reorder(g: CSRGraph, method: str = "rcm") -> Tuple[CSRGraph, torch.Tensor]:
    """Locality-improving node order for graphs that HAVE locality (co-purchase, social, road
    graphs; the uniform-random synthetic shapes have none to find).  Returns `(graph, perm)` with
    new = perm[old], like `dist.random_relabel`.

    "rcm": reverse Cuthill-McKee on the symmetrised pattern (scipy) -- neighbours get nearby ids, so
           the CBSR rows one CSR row gathers, and the `dXs` rows it reduces into, share cache lines
           and L2 residency; contiguous nnz-balanced shards (`row_partition_bounds`) then also cut
           few edges.
    "degree": descending in-degree -- the hub rows every CTA keeps gathering sit together at the
           head of the table (and the work records come out longest first without a second copy).
    Neither changes any result beyond the relabelling (tests/test_host.py)."""
    n = g.num_nodes()
    if method == "degree":
        order = torch.argsort(g.in_degrees().to(torch.int64), descending=True, stable=True)
    elif method == "rcm":
        import numpy as np
        import scipy.sparse as sp
        from scipy.sparse.csgraph import reverse_cuthill_mckee
        ptr = g.indptr.cpu().numpy().astype(np.int64)
        idx = g.indices.cpu().numpy()
        a = sp.csr_matrix((np.ones(idx.shape[0], dtype=np.int8), idx, ptr), shape=(n, n))
        order = torch.from_numpy(np.ascontiguousarray(
            reverse_cuthill_mckee(a, symmetric_mode=bool(g._cache.get("symmetric", False))).astype(np.int64)))
    else:
        raise ValueError(f"unknown reorder method {method!r}")
    order = order.to(g.device)                      # order[new] = old
    perm = torch.empty(n, dtype=torch.int64, device=g.device)
    perm[order] = torch.arange(n, dtype=torch.int64, device=g.device)
    return permute(g, perm), perm


def row_partition_bounds(indptr: torch.Tensor, world: int) -> list:
    """nnz-balanced contiguous row ranges, one per rank (balance by
    nnz, not rows).  Returns world+1 row offsets."""
    n = indptr.numel() - 1
    total = int(indptr[-1])
    bounds = [0]
    ptr64 = indptr.to(torch.int64)
    for p in range(1, world):
        tgt = total * p // world
        r = int(torch.searchsorted(ptr64, torch.tensor(tgt, device=indptr.device), right=False))
        r = min(max(r, bounds[-1]), n)
        bounds.append(r)
    bounds.append(n)
    return bounds


def index_dtype_for(dim_origin: int) -> torch.dtype:
    """uint8 column ids up to 256 columns (the reference's only mode),
    uint16 above (hidden 384 of scripts_train/yelp_maxk.sh)."""
    if dim_origin <= 256:
        return torch.uint8
    if dim_origin <= 65536:
        return torch.uint16
    raise ValueError("dim_origin above 65536 is not supported")


def ceil_div(a: int, b: int) -> int:
    return -(-a // b)


__all__ = [
    "CSRGraph",
    "SHAPES",
    "FEATS",
    "from_edges",
    "save_graph",
    "load_graph",
    "synthetic_graph",
    "shaped_graph",
    "row_partition_bounds",
    "permute",
    "reorder",
    "index_dtype_for",
    "ceil_div",
]
