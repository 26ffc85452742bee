"""The oracle against the reference's own outputs (tests/golden) and against itself
(numpy vs C vs literal loops vs dense linear algebra vs torch)."""
import numpy as np
import pytest
import torch

from oracle import c_oracle, maxk_oracle as mo, ref_torch
from spgemm_gnn_b200.graph import synthetic_graph


def _cases(golden):
    for ci in range(int(golden["num_cases"])):
        pre = f"c{ci}_"
        yield pre, golden[pre + "x"], int(golden[pre + "k"])


def test_maxk_dense_matches_reference_python(golden):
    """utils/models.py::MaxK and utils/maxk_layers.py::MaxKFunction (fallback) outputs."""
    for pre, x, k in _cases(golden):
        out, mask = mo.maxk_dense_forward(x, k)
        assert np.array_equal(out, golden[pre + "maxk_out"])
        gin = mo.maxk_dense_backward(golden[pre + "grad_out"], mask)
        assert np.array_equal(gin, golden[pre + "maxk_grad_in"])


def test_cbsr_layout_matches_reference_extract(golden):
    """MaxKSAGEConv._extract_sparse_format: ascending columns, uint8, values bit-copied."""
    seen = 0
    for pre, x, k in _cases(golden):
        if pre + "sp_data" not in golden.files:
            continue
        seen += 1
        for impl in (mo.maxk_cbsr, c_oracle.maxk_cbsr):
            sp_data, sp_index = impl(x, k)
            assert sp_index.dtype == np.uint8
            assert np.array_equal(sp_index, golden[pre + "sp_index"])
            assert np.array_equal(sp_data.view(np.uint32), golden[pre + "sp_data"].view(np.uint32))
    assert seen >= 4


def test_spgemm_as_the_reference_layers_call_it(golden_layers):
    """The reference's own call sites (utils/maxk_layers.py:166-171, 380-385) with the arguments its
    own code builds: the oracle's SpGEMM + the layer's epilogue reproduces what the same layer
    computes through `graph.update_all(copy_u, mean|sum)` (utils/maxk_layers.py:208-222, 392-405)."""
    from conftest import layer_cases
    seen = 0
    for name, c, y, add in layer_cases(golden_layers):
        seen += 1
        assert c["ptr"].dtype == np.int32 and c["sp_index"].dtype == np.uint8
        assert c["ptr"].size == c["n"] + 1 and c["idx"].size == c["e"] == c["val"].size
        assert c["sp_data"].shape == (c["n"], c["k"]) and y.shape == (c["n"], c["d"])
        for impl in (mo.spgemm_fwd, c_oracle.spgemm_fwd):
            out = impl(c["ptr"], c["idx"], c["val"], c["sp_data"], c["sp_index"], c["d"]) + add
            np.testing.assert_allclose(out, y, rtol=0, atol=2e-6 * np.abs(y).max())
        # the CBSR rows the reference extracts are what the MaxK contract says: ascending, distinct
        assert (np.diff(c["sp_index"].astype(np.int64), axis=1) > 0).all()
        # and its per-edge weights are the ones this repo's graph code produces (section 8 a-7)
        kind = {"sage_mean": "mean", "sage_sum": "sum", "sage_mean_wide": "mean"}.get(name)
        if kind:
            np.testing.assert_allclose(mo.edge_weights(c["ptr"], c["idx"], kind), c["val"], rtol=1e-7)
    assert seen == 4


def test_sspmm_against_the_backward_of_the_reference_layer(golden_layers):
    """Backward of the reference's DGL branch (autograd through `update_all`, then
    `MaxKFunction.backward` = grad * mask, utils/maxk_layers.py:37-45): the gradient at the input
    of the MaxK is the oracle's SSpMM scattered to the kept columns, and zero elsewhere."""
    from conftest import layer_cases
    seen = 0
    for name, c, _, _ in layer_cases(golden_layers):
        if f"{name}_dy" not in golden_layers.files:
            continue
        seen += 1
        dy, want = golden_layers[f"{name}_dy"], golden_layers[f"{name}_grad_maxk_in"]
        for impl in (mo.sspmm_bwd, c_oracle.sspmm_bwd):
            dxs = impl(c["ptr"], c["idx"], c["val"], dy, c["sp_index"])
            dense = mo.cbsr_scatter(dxs, c["sp_index"], c["d"])
            np.testing.assert_allclose(dense, want, rtol=0, atol=2e-6 * np.abs(want).max())
        assert np.count_nonzero(want) <= c["n"] * c["k"]
    assert seen == 3


def test_gin_layer_of_the_reference_end_to_end(golden_layers):
    """utils/integrated_models.py::MaxKGINConv: (1 + eps) * feat + sum over in-neighbours of MaxK(feat);
    no GEMM precedes the MaxK, so the oracle reproduces the input of the MLP from the raw features."""
    gl = golden_layers
    n, d_in, _, k = (int(v) for v in gl["gin_dims"])
    feat, ptr, idx = gl["gin_feat"], gl["gin_ptr"], gl["gin_idx"]
    eps = float(gl["gin_sd_eps"][0])
    assert eps == 0.25
    for topk, fwd in ((mo.maxk_cbsr, mo.spgemm_fwd), (c_oracle.maxk_cbsr, c_oracle.spgemm_fwd)):
        sp_data, sp_index = topk(feat, k)
        neigh = fwd(ptr, idx, np.ones(idx.size, np.float32), sp_data, sp_index, d_in)
        pre = (1 + eps) * feat.astype(np.float64) + neigh
        np.testing.assert_allclose(pre, gl["gin_pre_mlp"], rtol=0, atol=2e-6 * np.abs(pre).max())


def test_sage_layer_of_the_reference_end_to_end(golden_layers):
    """utils/maxk_layers.py::MaxKSAGEConv with scaled permutation matrices as weights (exact GEMMs):
    fc_self(x) + mean over in-neighbours of MaxK(fc_neigh(x)), from the raw features."""
    gl = golden_layers
    n, d, k = (int(v) for v in gl["sagex_dims"])
    feat, ptr, idx = gl["sagex_feat"], gl["sagex_ptr"], gl["sagex_idx"]
    w_self, w_neigh = gl["sagex_sd_fc_self.weight"], gl["sagex_sd_fc_neigh.weight"]
    assert (np.count_nonzero(w_self, axis=1) == 1).all() and (np.count_nonzero(w_neigh, axis=1) == 1).all()
    h_self, h_neigh = feat @ w_self.T, feat @ w_neigh.T          # one product per output: exact
    val = mo.edge_weights(ptr, idx, "mean")
    for topk, fwd in ((mo.maxk_cbsr, mo.spgemm_fwd), (c_oracle.maxk_cbsr, c_oracle.spgemm_fwd)):
        sp_data, sp_index = topk(h_neigh, k)
        y = h_self.astype(np.float64) + fwd(ptr, idx, val, sp_data, sp_index, d)
        np.testing.assert_allclose(y, gl["sagex_y"], rtol=0, atol=2e-6 * np.abs(y).max())


def test_padding_convention_is_harmless(golden):
    """Rows with fewer than k non-zeros are padded with (0.0, idx 0) by the reference; the
    accumulating dense view ignores them."""
    x = golden["pad_x"].astype(np.float64)
    k = int(golden["pad_k"])
    dense = mo.cbsr_to_dense(golden["pad_sp_data"], golden["pad_sp_index"], x.shape[1])
    want = x.copy()
    for r in range(x.shape[0]):        # more than k non-zeros: the reference keeps the first k
        nzc = np.nonzero(x[r])[0]
        want[r, nzc[k:]] = 0.0
    assert np.array_equal(dense, want)
    assert golden["pad_sp_index"].dtype == np.uint8 and golden["pad_sp_index"][2].tolist() == [0] * k


@pytest.mark.parametrize("n,d,k", [(64, 256, 32), (50, 64, 8), (30, 384, 16), (17, 100, 7), (9, 32, 32)])
def test_numpy_and_c_topk_agree_with_torch_on_tie_free_rows(n, d, k):
    rng = np.random.default_rng(97 + d + k)
    x = rng.standard_normal((n, d)).astype(np.float32)
    a_data, a_idx = mo.maxk_cbsr(x, k)
    b_data, b_idx = c_oracle.maxk_cbsr(x, k)
    assert np.array_equal(a_idx, b_idx) and np.array_equal(a_data, b_data)
    assert a_idx.dtype == (np.uint8 if d <= 256 else np.uint16)
    t_idx = torch.from_numpy(x).topk(k, dim=1)[1].sort(dim=1)[0].numpy()
    assert np.array_equal(a_idx.astype(np.int64), t_idx)
    assert np.all(np.diff(a_idx.astype(np.int64), axis=1) > 0)


def test_topk_tie_nan_zero_rules():
    nan, inf = np.nan, np.inf
    x = np.array([
        [1, 3, 3, 3, 0, 3, -1, 2],          # ties on the threshold -> lower columns win
        [0, 0, 0, 0, 0, 0, 0, 0],           # all equal
        [-0.0, 0.0, -0.0, 0.0, -1, -1, -1, -1],   # -0 == +0
        [nan, 1, inf, nan, -inf, 5, 4, nan],      # NaN above +inf
        [-inf, -inf, -inf, -5, -inf, -inf, -inf, -inf],
    ], dtype=np.float32)
    want = {
        2: [[1, 2], [0, 1], [0, 1], [0, 3], [0, 3]],
        3: [[1, 2, 3], [0, 1, 2], [0, 1, 2], [0, 3, 7], [0, 1, 3]],
        5: [[1, 2, 3, 5, 7], [0, 1, 2, 3, 4], [0, 1, 2, 3, 4], [0, 2, 3, 5, 7], [0, 1, 2, 3, 4]],
    }
    for k, cols in want.items():
        for impl in (mo.maxk_cbsr, c_oracle.maxk_cbsr):
            data, idx = impl(x, k)
            assert idx.tolist() == cols, (k, impl.__module__)
            assert np.array_equal(data.view(np.uint32),
                                  np.take_along_axis(x, idx.astype(np.int64), 1).view(np.uint32))


def _layer_inputs(n=300, avg_deg=12, d=64, k=16, seed=5, kind="mean"):
    g = synthetic_graph(n, n * avg_deg, seed=seed)
    ptr, idx = g.indptr.numpy(), g.indices.numpy()
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d)).astype(np.float32)
    dy = rng.standard_normal((n, d)).astype(np.float32)
    val = mo.edge_weights(ptr, idx, kind)
    return g, ptr, idx, val, x, dy


@pytest.mark.parametrize("kind", ["mean", "both", "sum"])
def test_spgemm_and_sspmm_formulas_cross_check(kind):
    g, ptr, idx, val, x, dy = _layer_inputs(kind=kind)
    k, d = 16, x.shape[1]
    assert np.allclose(val, g.edge_weights(kind).numpy(), rtol=1e-6)
    sp_data, sp_index = mo.maxk_cbsr(x, k)
    y = mo.spgemm_fwd(ptr, idx, val, sp_data, sp_index, d)
    y_c = c_oracle.spgemm_fwd(ptr, idx, val, sp_data, sp_index, d)
    y_dense, mask = mo.layer_dense_forward(ptr, idx, val, x, k)     # A @ (x*mask)
    np.testing.assert_allclose(y, y_dense, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(y_c, y_dense, rtol=1e-12, atol=1e-12)
    dxs = mo.sspmm_bwd(ptr, idx, val, dy, sp_index)
    dxs_c = c_oracle.sspmm_bwd(ptr, idx, val, dy, sp_index)
    dx_dense = mo.layer_dense_backward(ptr, idx, val, dy, mask)    # (A^T @ dY) * mask
    np.testing.assert_allclose(mo.cbsr_scatter(dxs, sp_index, d), dx_dense, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(dxs_c, dxs, rtol=1e-12, atol=1e-12)
    # adjointness: <A Xs, dY> == <Xs, dXs>
    assert np.isclose((y * dy).sum(), (sp_data.astype(np.float64) * dxs).sum(), rtol=1e-10)


def test_kernel_pseudocode_loops_agree_with_vectorised_forms():
    g, ptr, idx, val, x, dy = _layer_inputs(n=60, avg_deg=5, d=32, k=8, seed=3)
    sp_data, sp_index = mo.maxk_cbsr(x, 8)
    np.testing.assert_allclose(mo.spgemm_fwd_loops(ptr, idx, val, sp_data, sp_index, 32),
                               mo.spgemm_fwd(ptr, idx, val, sp_data, sp_index, 32), rtol=1e-13)
    np.testing.assert_allclose(mo.sspmm_bwd_loops(ptr, idx, val, dy, sp_index),
                               mo.sspmm_bwd(ptr, idx, val, dy, sp_index), rtol=1e-13, atol=1e-15)


def test_oracle_layer_matches_torch_reference_path():
    """numpy oracle == torch.topk + torch.sparse.mm autograd (the path the reference trains)."""
    g, ptr, idx, val, x, dy = _layer_inputs(n=200, avg_deg=9, d=64, k=16, seed=11)
    adj = ref_torch.csr_matrix(g.indptr, g.indices, torch.from_numpy(val), g.num_src)
    y_t, dx_t = ref_torch.layer_forward_backward(adj, torch.from_numpy(x), torch.from_numpy(dy), 16)
    y, mask = mo.layer_dense_forward(ptr, idx, val, x, 16)
    dx = mo.layer_dense_backward(ptr, idx, val, dy, mask)
    np.testing.assert_allclose(y_t.numpy(), y, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(dx_t.numpy(), dx, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("max_nz", [1, 3, 64, 1000])
def test_partition_records(max_nz):
    ptr = np.array([0, 0, 1, 4, 4, 68, 132, 133, 333], dtype=np.int32)  # degrees 0,1,3,0,64,64,1,200
    recs = mo.partition_rows(ptr, max_nz)
    recs_c, slots_c = c_oracle.partition_rows(ptr, max_nz)
    assert np.array_equal(recs, recs_c)
    # every stored entry is covered exactly once, in order, records never exceed max_nz
    cover = np.concatenate([np.arange(l, l + n) for _, l, n, _ in recs] or [np.zeros(0, int)])
    assert np.array_equal(cover, np.arange(ptr[-1]))
    assert recs[:, 2].max() <= max_nz
    # every row has at least one record (empty rows: len 0), rows ascending
    assert np.array_equal(np.unique(recs[:, 0]), np.arange(len(ptr) - 1))
    assert np.all(np.diff(recs[:, 0]) >= 0)
    # slot: -1 for single-record rows, else consecutive
    multi = recs[recs[:, 3] >= 0]
    assert np.array_equal(multi[:, 3], np.arange(len(multi)))
    assert slots_c == len(multi)
    rows, counts = np.unique(recs[:, 0], return_counts=True)
    for r, c in zip(rows, counts):
        assert np.all((recs[recs[:, 0] == r, 3] >= 0) == (c > 1))
    if max_nz == 64:  # the reference's WARP_MAX_NZ: ceil(deg/64) records per non-empty row
        assert len(recs) == 2 + 1 + 1 + 1 + 1 + 1 + 4


def test_scatter_gather_round_trip():
    rng = np.random.default_rng(1)
    x = rng.standard_normal((40, 96)).astype(np.float32)
    data, idx = mo.maxk_cbsr(x, 12)
    dense = mo.cbsr_scatter(data, idx, 96)
    assert np.array_equal(dense, c_oracle.cbsr_scatter(data, idx, 96))
    assert np.array_equal(mo.cbsr_gather(dense, idx), data)
    assert np.array_equal(c_oracle.cbsr_gather(dense, idx), data)
    out, _ = mo.maxk_dense_forward(x, 12)
    assert np.array_equal(dense, out)


def test_row_sample_checker_agrees_with_the_whole_problem_oracle():
    """The sampler the full-size GPU tests use (conftest.oracle_sample_check) cuts sub-problems out
    of the graph; fed with the whole-problem oracle's own results it must pass, and it must notice a
    single wrong element."""
    import pytest
    import torch
    from conftest import oracle_sample_check, small_graph
    from oracle import c_oracle
    g = small_graph(600, 20)
    n, d, k = g.num_nodes(), 64, 8
    rng = np.random.default_rng(3)
    x = rng.standard_normal((n, d)).astype(np.float32)
    dy = rng.standard_normal((n, d)).astype(np.float32)
    val = g.edge_weights("both")
    wd, wi = c_oracle.maxk_cbsr(x, k)
    ptr, idx = g.indptr.numpy(), g.indices.numpy()
    out = torch.from_numpy(c_oracle.spgemm_fwd(ptr, idx, val.numpy(), wd, wi, d).astype(np.float32))
    dxs = torch.from_numpy(c_oracle.sspmm_bwd(ptr, idx, val.numpy(), dy, wi).astype(np.float32))
    args = (g, val, torch.from_numpy(wd), torch.from_numpy(wi))
    oracle_sample_check(*args, out, torch.from_numpy(dy), dxs, d, k, n_sample=600)
    bad = out.clone()
    bad[17, int(wi[g.indices[g.indptr[17]]][0])] += 1e-2
    with pytest.raises(AssertionError):
        oracle_sample_check(*args, bad, torch.from_numpy(dy), dxs, d, k, n_sample=600)
    badb = dxs.clone()
    badb[5, 3] += 1e-2
    with pytest.raises(AssertionError):
        oracle_sample_check(*args, out, torch.from_numpy(dy), badb, d, k, n_sample=600)
