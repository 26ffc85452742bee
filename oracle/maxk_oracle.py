"""CPU ORACLE (test infrastructure, NOT product code) -- numpy restatement of the
MaxK-GNN aggregation hot path of julius-sk/spgemm-gnn.

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs
of `bench.py` may import this package.  The product (`spgemm_gnn_b200`, `maxk_kernels`)
never does and fails loudly when its CUDA library is missing.

Parity status:
  * MaxK forward/backward and the CBSR extraction are PINNED against the reference's own
    Python (`utils/models.py::MaxK`, `utils/maxk_layers.py::MaxKFunction` fallback branch
    and `MaxKSAGEConv._extract_sparse_format`), run in the build container with a stub
    `dgl` module by `tests/golden/make_golden.py`; the vectors are committed under
    `tests/golden/`.
  * SpGEMM forward, the per-edge weights and the SSpMM backward are ANCHORED ON THE REFERENCE'S OWN
    CALL SITES: `tests/golden/make_golden_layers.py` runs the reference's `MaxKSAGEConv` /
    `MaxKGCNConv` code (utils/maxk_layers.py:47-447) through its DGL branch and through its
    `maxk_kernels.spgemm_forward` call site (with this oracle as the kernel stand-in and a
    message-passing stand-in for DGL's `update_all`), checks that the two agree, and commits the
    recorded call arguments, the layer outputs and the layer backward under `tests/golden/`.
  * What nothing in /root/reference can pin: DGL's own float32 arithmetic (`update_all` is library
    code; conda `dglteam/label/cu121`, version unpinned, README.md:47; absent from the image) and
    the native kernels' arithmetic (sm_80-only binary without sources).  There the formulas below
    restate the SASS decode in SURVEY.md section 2.3 (K3, K4) and are cross-checked against dense
    linear algebra (`A @ dense(Xs)`, `(A^T @ dY)` sampled) and scipy.  Partitioning (a-5) is
    oracle-anchored only: PARITY UNPINNED for it.

All sums are carried in float64.
"""
from __future__ import annotations

import numpy as np

try:  # scipy is only a second opinion / speed-up, never required
    import scipy.sparse as _sp
except Exception:  # pragma: no cover
    _sp = None


# ---------------------------------------------------------------------------------------
# a-1  MaxK -> CBSR      (utils/models.py:12-20, utils/maxk_layers.py:16-34, K1)
# ---------------------------------------------------------------------------------------
def order_key(x: np.ndarray) -> np.ndarray:
    """uint32 key whose unsigned order is the selection order of the build contract
    (SURVEY.md section 8 a-1): larger value first, every NaN above +inf (torch.topk
    convention), -0.0 and +0.0 equal."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    bits = x.view(np.uint32)
    neg = (bits >> np.uint32(31)).astype(bool)
    key = np.where(neg, ~bits, bits | np.uint32(0x80000000)).astype(np.uint32)
    key = np.where(bits == np.uint32(0x80000000), np.uint32(0x80000000), key)  # -0 == +0
    key = np.where(np.isnan(x), np.uint32(0xFFFFFFFF), key)
    return key


def index_dtype_for(dim_origin: int):
    return np.uint8 if dim_origin <= 256 else np.uint16


def maxk_cbsr(x: np.ndarray, k: int):
    """Exact top-k per row, emitted as CBSR.

    Selection: by value descending; ties -> LOWER column wins.  Layout: `sp_data`
    fp32 [N,k] bit-copies of the selected inputs, `sp_index` uint8 (uint16 when D>256)
    [N,k], entries of a row in ASCENDING column order -- the layout K1 writes
     and `_extract_sparse_format` (utils/maxk_layers.py:236-257)
    rebuilds with `torch.nonzero`.
    """
    x = np.ascontiguousarray(x, dtype=np.float32)
    n, d = x.shape
    if not (1 <= k <= d):
        raise ValueError("k must be between 1 and input dimension")
    key = order_key(x).astype(np.int64)
    # stable sort on descending key: equal keys keep ascending column order
    order = np.argsort(-key, axis=1, kind="stable")[:, :k]
    cols = np.sort(order, axis=1)
    sp_data = np.take_along_axis(x, cols, axis=1)
    return sp_data, cols.astype(index_dtype_for(d))


def cbsr_to_dense(sp_data: np.ndarray, sp_index: np.ndarray, dim_origin: int, dtype=np.float64):
    """dense[i, sp_index[i,t]] += sp_data[i,t] (accumulating, so padded duplicate
    `(0.0, idx 0)` entries of utils/maxk_layers.py:245-257 stay harmless)."""
    n, k = sp_data.shape
    out = np.zeros((n, dim_origin), dtype=dtype)
    rows = np.repeat(np.arange(n), k)
    np.add.at(out, (rows, sp_index.reshape(-1).astype(np.int64)), sp_data.reshape(-1).astype(dtype))
    return out


def maxk_dense_forward(x: np.ndarray, k: int):
    """`MaxK.forward` (utils/models.py:14-20): topk -> 0/1 mask -> multiply.  Returns
    (output, mask).  Uses the tie rule of `maxk_cbsr`; on tie-free rows it is the same
    set torch.topk picks."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    _, cols = maxk_cbsr(x, k)
    mask = np.zeros_like(x)
    np.put_along_axis(mask, cols.astype(np.int64), 1.0, axis=1)
    return x * mask, mask


def maxk_dense_backward(grad_output: np.ndarray, mask: np.ndarray):
    """`MaxK.backward` (utils/models.py:23-26): grad * mask."""
    return grad_output * mask


# ---------------------------------------------------------------------------------------
# a-2  CBSR gradient -> dense   (K2: maxk_backward_cuda host loop)
# ---------------------------------------------------------------------------------------
def cbsr_scatter(grad: np.ndarray, sp_index: np.ndarray, dim_origin: int):
    """zeros [N,D]; for j ascending: dense[i, sp_index[i,j]] = grad[i,j]  (copy, the
    last write to a column wins -- K2 in SURVEY.md section 2.3)."""
    n, k = grad.shape
    out = np.zeros((n, dim_origin), dtype=grad.dtype)
    idx = sp_index.astype(np.int64)
    for j in range(k):
        out[np.arange(n), idx[:, j]] = grad[:, j]
    return out


def cbsr_gather(dense: np.ndarray, sp_index: np.ndarray):
    """values of a dense [N,D] matrix at the CBSR positions -> [N,k]."""
    return np.take_along_axis(dense, sp_index.astype(np.int64), axis=1)


# ---------------------------------------------------------------------------------------
# a-5  warp partitions   (generate_meta.py -> .warp4 records, SURVEY.md K3 host notes)
# ---------------------------------------------------------------------------------------
def partition_rows(indptr: np.ndarray, max_nz: int) -> np.ndarray:
    """int32 [P,4] records {row, loc, len, slot}: every CSR row cut into chunks of at
    most `max_nz` stored entries, rows ascending, chunks ascending.

    Differences from the reference file format, both deliberate: (1) an empty row still
    gets one record with len 0, because the product writes every output row itself
    instead of relying on a zero-filled output; (2) the unused 4th word carries the
    partial-sum slot: -1 for a row with a single chunk (written straight to the
    output), otherwise a running index into the partial buffer that a fixed-order
    reduction folds (no float atomics, unlike K3's RED write-back).
    """
    indptr = np.asarray(indptr, dtype=np.int64)
    n = indptr.size - 1
    recs = []
    slot = 0
    for r in range(n):
        lo, hi = int(indptr[r]), int(indptr[r + 1])
        deg = hi - lo
        chunks = max(1, -(-deg // max_nz))
        for c in range(chunks):
            loc = lo + c * max_nz
            ln = max(0, min(max_nz, hi - loc))
            if chunks == 1:
                recs.append((r, loc, ln, -1))
            else:
                recs.append((r, loc, ln, slot))
                slot += 1
    return np.asarray(recs, dtype=np.int32).reshape(-1, 4)


# ---------------------------------------------------------------------------------------
# a-3  forward SpGEMM    (K3: spmm_kernel_opt2_sparse_v3)
# ---------------------------------------------------------------------------------------
def spgemm_fwd(indptr, indices, val, sp_data, sp_index, dim_origin: int) -> np.ndarray:
    """Y[i, sp_index[j,t]] += val[e] * sp_data[j,t] for every stored e=(i<-j), t<k.
    float64 [N_rows, D]."""
    indptr = np.asarray(indptr, dtype=np.int64)
    indices = np.asarray(indices, dtype=np.int64)
    n = indptr.size - 1
    xs = cbsr_to_dense(sp_data, sp_index, dim_origin)
    if _sp is not None:
        a = _sp.csr_matrix(
            (np.asarray(val, dtype=np.float64), indices, indptr), shape=(n, sp_data.shape[0])
        )
        return np.asarray(a @ xs)
    out = np.zeros((n, dim_origin))
    for i in range(n):
        lo, hi = indptr[i], indptr[i + 1]
        out[i] = (np.asarray(val[lo:hi], dtype=np.float64)[:, None] * xs[indices[lo:hi]]).sum(0)
    return out


def spgemm_fwd_loops(indptr, indices, val, sp_data, sp_index, dim_origin: int) -> np.ndarray:
    """The K3 pseudocode literally, edge by edge (small cases only)."""
    n = len(indptr) - 1
    k = sp_data.shape[1]
    out = np.zeros((n, dim_origin))
    for i in range(n):
        for e in range(int(indptr[i]), int(indptr[i + 1])):
            nz = int(indices[e])
            v = float(val[e])
            for t in range(k):
                out[i, int(sp_index[nz, t])] += v * float(sp_data[nz, t])
    return out


# ---------------------------------------------------------------------------------------
# a-4  backward SSpMM    (K4: spmm_kernel_opt2_sparse_backward_v3)
# ---------------------------------------------------------------------------------------
def sspmm_bwd(indptr, indices, val, grad_out, sp_index, num_src=None) -> np.ndarray:
    """dXs[j,t] = sum over stored e=(r<-j) of val[e] * dY[r, sp_index[j,t]].
    float64 [N_src, k]."""
    indptr = np.asarray(indptr, dtype=np.int64)
    indices = np.asarray(indices, dtype=np.int64)
    n = indptr.size - 1
    ns = sp_index.shape[0] if num_src is None else num_src
    dy = np.asarray(grad_out, dtype=np.float64)
    if _sp is not None:
        a = _sp.csr_matrix((np.asarray(val, dtype=np.float64), indices, indptr), shape=(n, ns))
        g = np.asarray(a.T @ dy)
    else:
        g = np.zeros((ns, dy.shape[1]))
        rows = np.repeat(np.arange(n), np.diff(indptr))
        np.add.at(g, indices, np.asarray(val, dtype=np.float64)[:, None] * dy[rows])
    return np.take_along_axis(g, sp_index.astype(np.int64), axis=1)


def sspmm_bwd_loops(indptr, indices, val, grad_out, sp_index) -> np.ndarray:
    """The K4 pseudocode literally (small cases only)."""
    n = len(indptr) - 1
    ns, k = sp_index.shape
    out = np.zeros((ns, k))
    for r in range(n):
        for e in range(int(indptr[r]), int(indptr[r + 1])):
            nz = int(indices[e])
            v = float(val[e])
            for t in range(k):
                out[nz, t] += v * float(grad_out[r, int(sp_index[nz, t])])
    return out


def abs_bound_fwd(indptr, indices, val, sp_data, sp_index, dim_origin: int) -> np.ndarray:
    """sum of |terms| per output element: the scale the 1e-5 relative tolerance of
    BASELINE.json is measured against ('fp32 tolerance')."""
    return spgemm_fwd(indptr, indices, np.abs(val), np.abs(sp_data), sp_index, dim_origin)


def abs_bound_bwd(indptr, indices, val, grad_out, sp_index) -> np.ndarray:
    return sspmm_bwd(indptr, indices, np.abs(val), np.abs(grad_out), sp_index)


# ---------------------------------------------------------------------------------------
# a-7  edge weights  (utils/maxk_layers.py:147-159, DGL GraphConv norm='both')
# ---------------------------------------------------------------------------------------
def edge_weights(indptr, indices, kind: str, num_src=None) -> np.ndarray:
    indptr = np.asarray(indptr, dtype=np.int64)
    indices = np.asarray(indices, dtype=np.int64)
    n = indptr.size - 1
    ns = n if num_src is None else num_src
    deg_in = np.diff(indptr)
    rows = np.repeat(np.arange(n), deg_in)
    if kind in ("sum", "none"):
        return np.ones(indices.size, dtype=np.float32)
    if kind in ("mean", "right"):
        return (np.float32(1.0) / np.maximum(deg_in, 1).astype(np.float32))[rows]
    if kind == "both":
        deg_out = np.bincount(indices, minlength=ns)
        di = np.maximum(deg_in, 1).astype(np.float32) ** np.float32(-0.5)
        do = np.maximum(deg_out, 1).astype(np.float32) ** np.float32(-0.5)
        return (di[rows] * do[indices]).astype(np.float32)
    raise ValueError(kind)


# ---------------------------------------------------------------------------------------
# dense restatement of one layer, the way the reference actually trains
# (utils/models.py:157-163: MaxK then graph.update_all(copy_u, mean|sum))
# ---------------------------------------------------------------------------------------
def layer_dense_forward(indptr, indices, val, x, k):
    """Y = A @ (x * mask).  float64."""
    xm, mask = maxk_dense_forward(x, k)
    indptr = np.asarray(indptr, dtype=np.int64)
    n = indptr.size - 1
    rows = np.repeat(np.arange(n), np.diff(indptr))
    y = np.zeros((n, x.shape[1]))
    np.add.at(y, rows, np.asarray(val, np.float64)[:, None] * xm.astype(np.float64)[np.asarray(indices)])
    return y, mask


def layer_dense_backward(indptr, indices, val, grad_y, mask):
    """dX = (A^T @ dY) * mask.  float64."""
    indptr = np.asarray(indptr, dtype=np.int64)
    n = indptr.size - 1
    rows = np.repeat(np.arange(n), np.diff(indptr))
    g = np.zeros((mask.shape[0], grad_y.shape[1]))
    np.add.at(g, np.asarray(indices), np.asarray(val, np.float64)[:, None] * np.asarray(grad_y, np.float64)[rows])
    return g * mask


# ---------------------------------------------------------------------------------------
# banked CBSR (product-internal format, csrc/bank.cu): validity check + conflict count
# ---------------------------------------------------------------------------------------
def banked_rows(dim_origin: int) -> int:
    return (dim_origin + 7) // 8 + 8 * ((dim_origin + 63) // 64)


def check_banked(sp_data, sp_index, bk_data, bk_index, bk_slot, dim_origin: int):
    """Asserts that (bk_data, bk_index) is a per-row permutation of (sp_data, sp_index) and
    that every bk_slot is one of the two legal cells of its column.  Returns the mean number
    of shared-memory wavefronts per 8-lane step (1.0 = conflict-free)."""
    n, k = sp_data.shape
    cap = k // 8
    ra = (dim_origin + 7) // 8
    order = np.argsort(bk_index.astype(np.int64), axis=1, kind="stable")
    assert np.array_equal(np.take_along_axis(bk_index, order, 1), sp_index)
    assert np.array_equal(np.take_along_axis(bk_data, order, 1).view(np.uint32), sp_data.view(np.uint32))
    c = bk_index.astype(np.int64)
    slot = bk_slot.astype(np.int64) & 0xFFFF
    slot_a = 32 * (c >> 3) + (c & 7)
    slot_b = 32 * (ra + ((c >> 6) << 3) + (c & 7)) + ((c >> 3) & 7)
    assert np.all((slot == slot_a) | (slot == slot_b))
    assert slot.max() < 32 * banked_rows(dim_origin)
    bank = slot & 7  # inside the group's octet
    assert np.all(((slot >> 3) & 3) == 0)  # group offset is added by the kernel
    steps = bank.reshape(n, 8, cap)        # position p = cap * t + q
    wf = np.zeros((n, cap))
    for q in range(cap):
        b = steps[:, :, q]
        wf[:, q] = np.max(np.stack([(b == x).sum(1) for x in range(8)], 1), 1)
    return float(wf.mean()), wf
