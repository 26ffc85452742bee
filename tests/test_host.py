"""Host-side logic: CSR container, synthetic shapes, edge weights, row partition (no GPU)."""
import numpy as np
import pytest
import torch

from oracle import maxk_oracle as mo
from spgemm_gnn_b200 import graph as G


def test_synthetic_graph_is_symmetric_sorted_with_self_loops():
    g = G.synthetic_graph(500, 6000, seed=97)
    ptr, idx = g.indptr.numpy(), g.indices.numpy()
    n = g.num_nodes()
    assert ptr[0] == 0 and ptr[-1] == idx.size and np.all(np.diff(ptr) >= 1)
    rows = np.repeat(np.arange(n), np.diff(ptr))
    a = set(zip(rows.tolist(), idx.tolist()))
    assert len(a) == idx.size                                  # no duplicate edges
    assert all((c, r) in a for r, c in a)                      # symmetric
    assert all((i, i) in a for i in range(n))                  # self-loops
    for r in range(n):
        assert np.all(np.diff(idx[ptr[r]:ptr[r + 1]]) > 0)     # ascending neighbours
    assert 0.6 * 6000 < idx.size < 1.4 * 6000
    g2 = G.synthetic_graph(500, 6000, seed=97)
    assert torch.equal(g.indices, g2.indices)                 # seeded


@pytest.mark.parametrize("kind", ["mean", "both", "sum", "right"])
def test_edge_weights_match_oracle(kind):
    g = G.synthetic_graph(300, 3000, seed=3)
    w = g.edge_weights(kind).numpy()
    ref = mo.edge_weights(g.indptr.numpy(), g.indices.numpy(), kind)
    np.testing.assert_allclose(w, ref, rtol=1e-6)
    if kind == "mean":       # rows of a mean adjacency sum to one
        rows = g.row_ids().numpy()
        np.testing.assert_allclose(np.bincount(rows, weights=w), 1.0, rtol=1e-5)


def test_row_partition_bounds_balance_nnz_and_slices_keep_global_columns():
    g = G.synthetic_graph(1000, 20000, seed=9)
    for world in (1, 2, 4, 8):
        b = G.row_partition_bounds(g.indptr, world)
        assert b[0] == 0 and b[-1] == 1000 and len(b) == world + 1 and b == sorted(b)
        nnz = [int(g.indptr[b[p + 1]] - g.indptr[b[p]]) for p in range(world)]
        assert max(nnz) <= 1.3 * g.num_edges() / world + 2000
        cat = torch.cat([g.row_slice(b[p], b[p + 1]).indices for p in range(world)])
        assert torch.equal(cat, g.indices)
        for p in range(world):
            s = g.row_slice(b[p], b[p + 1])
            assert s.num_src == 1000 and s.indptr[0] == 0 and int(s.indptr[-1]) == s.num_edges()


def test_shapes_and_index_dtype():
    assert G.SHAPES["reddit"] == (232965, 114615891)
    assert G.index_dtype_for(256) == torch.uint8 and G.index_dtype_for(384) == torch.uint16
    g = G.shaped_graph("flickr", scale=0.02)
    assert abs(g.num_nodes() - 1785) <= 1
    with pytest.raises(TypeError):
        G.CSRGraph(torch.zeros(3, dtype=torch.int64), torch.zeros(0, dtype=torch.int32))


def test_dgl_surface_used_by_reference_layers():
    g = G.synthetic_graph(50, 300, seed=1)
    assert g.number_of_nodes() == 50 and g.num_edges() == g.indices.numel()
    assert hasattr(g, "_sparse_format")                      # utils/maxk_layers.py:94
    ptr, idx, eid = g.adj_tensors("csr")                     # symmetric: csr == csc
    assert ptr is g.indptr and idx is g.indices and eid.numel() == g.num_edges()
    with g.local_scope():
        assert int(g.in_degrees().sum()) == g.num_edges() == int(g.out_degrees().sum())
    assert torch.equal(g.in_degrees(), g.out_degrees())      # symmetric graph


def test_graph_file_round_trip(tmp_path):
    g = G.synthetic_graph(200, 1500, seed=4)
    path = str(tmp_path / "g.npz")
    G.save_graph(g, path)
    h = G.load_graph(path)
    assert torch.equal(g.indptr, h.indptr) and torch.equal(g.indices, h.indices)
    assert h.num_src == g.num_src and h._cache["symmetric"] is True
    import numpy as np
    np.savez(str(tmp_path / "bad.npz"), magic=np.array("nope"))
    with pytest.raises(ValueError):
        G.load_graph(str(tmp_path / "bad.npz"))


def test_extract_sparse_format_matches_reference_method(golden):
    """Vectorised stand-in of MaxKSAGEConv._extract_sparse_format against the reference's own
    outputs (tests/golden), including the (0.0, idx 0) padding rows."""
    from spgemm_gnn_b200.maxk_layers import extract_sparse_format, MaxKSAGEConv
    d, i = extract_sparse_format(torch.from_numpy(golden["pad_x"]), int(golden["pad_k"]))
    assert np.array_equal(d.numpy(), golden["pad_sp_data"]) and np.array_equal(i.numpy(), golden["pad_sp_index"])
    for ci in range(int(golden["num_cases"])):
        pre = f"c{ci}_"
        if pre + "sp_index" not in golden.files:
            continue
        conv = MaxKSAGEConv(8, 8, maxk=int(golden[pre + "k"]))
        d, i = conv._extract_sparse_format(torch.from_numpy(golden[pre + "maxk_out"]))
        assert np.array_equal(d.numpy(), golden[pre + "sp_data"])
        assert np.array_equal(i.numpy(), golden[pre + "sp_index"])


def test_synthetic_task_shapes_on_cpu():
    from spgemm_gnn_b200.train import synthetic_task
    g, feats, labels, tr, va, te, fin, ncls = synthetic_task("flickr", 0.02, "cpu")
    n = g.num_nodes()
    assert feats.shape == (n, 500) and (fin, ncls) == (500, 7) and labels.max() < 7
    assert int(tr.sum() + va.sum() + te.sum()) == n and 0.55 < tr.float().mean() < 0.77


def test_edge_weights_match_the_reference_layer_loop(golden_layers):
    """CSRGraph.edge_weights against the per-node `.item()` loop of the reference
    (utils/maxk_layers.py:148-157), as recorded by tests/golden/make_golden_layers.py."""
    from conftest import layer_cases
    from spgemm_gnn_b200.graph import CSRGraph
    for name, c, _, _ in layer_cases(golden_layers):
        if not name.startswith("sage"):
            continue
        g = CSRGraph(torch.from_numpy(c["ptr"]), torch.from_numpy(c["idx"]))
        kind = "sum" if name == "sage_sum" else "mean"
        got = g.edge_weights(kind).numpy()
        np.testing.assert_allclose(got, c["val"], rtol=1e-7)


def test_aligned_linear_is_the_same_layer():
    """MAXK_ALIGN_GEMM: zero-padded features and on-the-fly padded weights give the output and the
    gradients of the plain nn.Linear (Reddit's 602 inputs / 41 classes are not multiples of 4)."""
    import torch.nn as nn
    from spgemm_gnn_b200 import models
    torch.manual_seed(0)
    lin = nn.Linear(602, 41).double()
    x = torch.randn(50, 602, dtype=torch.float64)
    gy = torch.randn(50, 41, dtype=torch.float64)
    y0 = lin(x)
    y0.backward(gy)
    g0 = (lin.weight.grad.clone(), lin.bias.grad.clone())
    lin.zero_grad()
    xp = models.pad_features(x)
    assert xp.shape == (50, 608) and not xp[:, 602:].any() and models.pad_features(xp) is xp
    was = models.align_gemm()
    models.set_align_gemm(False)
    try:
        assert torch.equal(models.aligned_linear(lin, x), y0)         # no padding asked for, none needed
        models.set_align_gemm(True)
        y1 = models.aligned_linear(lin, xp)
        assert y1.shape == (50, 41)
        y1.backward(gy)
    finally:
        models.set_align_gemm(was)
    torch.testing.assert_close(y1, y0, rtol=1e-12, atol=1e-12)
    torch.testing.assert_close(lin.weight.grad, g0[0], rtol=1e-12, atol=1e-12)
    torch.testing.assert_close(lin.bias.grad, g0[1], rtol=1e-12, atol=1e-12)
    with pytest.raises(RuntimeError, match="expects 602"):
        models.aligned_linear(lin, x[:, :600])


def test_tools_and_drivers_compile():
    """Every script the GPU runs use (tools/, bench.py, __graft_entry__.py) at least byte-compiles."""
    import glob
    import os
    import py_compile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = glob.glob(os.path.join(root, "tools", "*.py")) + [os.path.join(root, "bench.py"),
                                                              os.path.join(root, "__graft_entry__.py")]
    assert len(files) >= 8
    for f in files:
        py_compile.compile(f, doraise=True)


def test_forward_phases_cover_every_block_once_in_arrival_order():
    """`forward_phases`: the rotated order rank, rank+1, ... cut into contiguous phases; every block in
    exactly one phase, own block first."""
    from spgemm_gnn_b200.maxk_kernels import forward_phases
    for world in (1, 2, 3, 4, 5, 8, 16):
        for rank in range(world):
            seen = []
            for a0, a1, b0, b1 in forward_phases(world, rank):
                assert 0 <= a0 <= a1 <= world and 0 <= b0 <= b1 <= world
                seen += list(range(a0, a1)) + list(range(b0, b1))
            assert seen == [(rank + s) % world for s in range(world)], (world, rank, seen)
    assert forward_phases(8, 5) == [(5, 6, 0, 0), (6, 8, 0, 1), (1, 5, 0, 0)]


def test_from_dgl_builds_the_in_edge_csr_and_caches_it():
    """`graph.from_dgl` on a stand-in with DGL's surface (DGL is not in the image): rows are destinations,
    in-neighbours ascending, parallel edges kept, result cached on the object."""
    import torch
    from spgemm_gnn_b200.graph import CSRGraph, from_dgl

    class FakeDGL:
        def __init__(self, src, dst, n):
            self._s, self._d, self._n = torch.tensor(src), torch.tensor(dst), n

        def edges(self):
            return self._s, self._d

        def num_nodes(self):
            return self._n

    g = FakeDGL([2, 0, 1, 2, 2, 3], [0, 0, 0, 1, 1, 3], 5)       # 2->1 twice
    c = from_dgl(g)
    assert c.indptr.tolist() == [0, 3, 5, 5, 6, 6]
    assert c.indices.tolist() == [0, 1, 2, 2, 2, 3]
    assert from_dgl(g) is c and from_dgl(c) is c
    assert c.in_degrees().tolist() == [3, 2, 0, 1, 0]
    assert isinstance(c, CSRGraph)


def test_reference_gcn_weights_are_the_reference_layers_own(golden_layers):
    """`edge_weights('reference_gcn')`: what the reference's MaxKGCNConv applies in total to an edge --
    its recorded per-edge weight `norm_right[idx]` (deg_in of the SOURCE, utils/maxk_layers.py:372-376)
    times the deg_out^-1/2 it folds into the features (:315-318) -- on the graph of the golden call."""
    import torch
    from spgemm_gnn_b200.graph import CSRGraph
    z = golden_layers
    g = CSRGraph(torch.from_numpy(z["gcn_both_call_ptr"].astype(np.int32)),
                 torch.from_numpy(z["gcn_both_call_idx"].astype(np.int32)))
    idx = z["gcn_both_call_idx"].astype(np.int64)
    do = g.out_degrees().clamp(min=1).float().pow(-0.5).numpy()
    want = z["gcn_both_call_val"].astype(np.float32) * do[idx]
    got = g.edge_weights("reference_gcn").numpy()
    np.testing.assert_allclose(got, want, rtol=1e-6, atol=0)
    # and it is not GraphConv's 'both' unless the graph is regular
    assert not np.allclose(g.edge_weights("both").numpy(), got)


def test_reorder_is_a_relabelling_and_rcm_restores_locality():
    """`graph.reorder`: the permuted graph is P A P^T (the oracle's forward commutes with it), RCM brings a
    shuffled banded graph back to a narrow band, "degree" puts the rows in descending-degree order."""
    rng = np.random.default_rng(5)
    n, half = 600, 6
    r = np.repeat(np.arange(n), 2 * half + 1)
    c = r + np.tile(np.arange(-half, half + 1), n)
    ok = (c >= 0) & (c < n)
    band = G.from_edges(torch.from_numpy(r[ok]), torch.from_numpy(c[ok]), n, symmetric=True)
    shuffle = torch.from_numpy(rng.permutation(n))
    g = G.permute(band, shuffle)
    assert g.num_edges() == band.num_edges()

    def bandwidth(h):
        return int((h.row_ids() - h.indices.to(torch.int64)).abs().max())

    assert bandwidth(band) == half and bandwidth(g) > n // 2
    g_rcm, perm = G.reorder(g, "rcm")
    assert bandwidth(g_rcm) <= 2 * half + 1
    assert sorted(perm.tolist()) == list(range(n))

    k, d = 4, 16
    x = rng.standard_normal((n, d)).astype(np.float32)
    for h, p in ((g_rcm, perm), G.reorder(g, "degree")):
        p = p.numpy()
        x_new = np.empty_like(x)
        x_new[p] = x
        sd0, si0 = mo.maxk_cbsr(x, k)
        sd1, si1 = mo.maxk_cbsr(x_new, k)
        out0 = mo.spgemm_fwd(g.indptr.numpy(), g.indices.numpy(), g.edge_weights("mean").numpy(),
                                 sd0, si0, d)
        out1 = mo.spgemm_fwd(h.indptr.numpy(), h.indices.numpy(), h.edge_weights("mean").numpy(),
                                 sd1, si1, d)
        np.testing.assert_allclose(out1[p], out0, rtol=1e-6, atol=1e-7)
    g_deg, _ = G.reorder(g, "degree")
    deg = g_deg.in_degrees()
    assert bool((deg[:-1] >= deg[1:]).all())
    with pytest.raises(ValueError):
        G.reorder(g, "nope")


def test_linear_is_nn_linear_with_the_same_state_dict_and_cpu_gradients():
    """`maxk_layers.Linear` / `weight_grad`: on the CPU (and below the row threshold) it IS nn.Linear -- same
    parameter names, same outputs, same gradients bit for bit; the row-blocked form is a CUDA-only path."""
    from spgemm_gnn_b200 import maxk_layers as ML
    torch.manual_seed(3)
    ref = torch.nn.Linear(12, 7)
    ours = ML.Linear(12, 7)
    assert list(ours.state_dict().keys()) == list(ref.state_dict().keys())
    ours.load_state_dict(ref.state_dict())
    x = torch.randn(50, 12)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ya, yb = ref(xa), ours(xb)
    assert torch.equal(ya, yb)
    g = torch.randn(50, 7)
    ya.backward(g)
    yb.backward(g)
    assert torch.equal(xa.grad, xb.grad) and torch.equal(ref.weight.grad, ours.weight.grad)
    assert torch.equal(ref.bias.grad, ours.bias.grad)
    assert torch.equal(ML.weight_grad(g, x), g.t().mm(x))
    assert isinstance(ours, torch.nn.Linear)
