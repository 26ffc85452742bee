"""Parity of the CUDA path (through `maxk_kernels` -> ctypes -> C ABI) with the CPU oracle.

Bars (BASELINE.json north_star): column ids and CBSR layout bit-exact; SpGEMM / SSpMM values
within 1e-5 relative, measured against sum|terms| of each output element (the fp32 sums are
formed in a different order than the float64 oracle's; SURVEY.md section 7 'fp32 tolerance').
"""
import numpy as np
import pytest
import torch

from conftest import oracle_sample_check

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5


@pytest.fixture(scope="module")
def mk():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from spgemm_gnn_b200 import build as _build
    _build.build()                      # no-op when the in-tree library matches the sources
    import maxk_kernels
    from spgemm_gnn_b200 import _lib
    assert _lib.lib().mk_device_ok() == 0, "not an sm_100 device"
    return maxk_kernels


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def assert_rel(got, want64, bound, what):
    got = got.detach().cpu().numpy().astype(np.float64)
    err = np.abs(got - want64)
    lim = REL_TOL * bound + 1e-30
    worst = float((err / lim).max()) if err.size else 0.0
    assert worst <= 1.0, f"{what}: error is {worst:.3g} x the 1e-5*sum|terms| bound"


# ---------------------------------------------------------------------------------------
# a-1 / a-2  MaxK -> CBSR, scatter, gather
# ---------------------------------------------------------------------------------------
TOPK_CASES = [(257, 256, 32), (100, 256, 8), (100, 256, 16), (64, 256, 64), (33, 256, 256),
              (50, 64, 8), (50, 32, 32), (40, 96, 1), (37, 100, 7), (31, 130, 20),
              (40, 384, 16), (20, 512, 32), (20, 768, 64), (12, 1024, 32), (9, 1500, 40),
              (5, 4100, 33), (1, 256, 32), (3, 49152, 40), (3, 50000, 33), (2, 65536, 64)]


@pytest.mark.parametrize("n,d,k", TOPK_CASES)
def test_topk_cbsr_bit_exact(mk, n, d, k):
    from oracle import c_oracle
    rng = np.random.default_rng(n * 7 + d + k)
    x = rng.standard_normal((n, d)).astype(np.float32)
    data, idx = mk.maxk_forward_cbsr(dev(x), k)
    want_data, want_idx = c_oracle.maxk_cbsr(x, k)
    assert idx.dtype == (torch.uint8 if d <= 256 else torch.uint16)
    got_idx = idx.cpu().numpy() if d <= 256 else idx.view(torch.int16).cpu().numpy().view(np.uint16)
    assert np.array_equal(got_idx, want_idx)
    assert np.array_equal(data.cpu().numpy().view(np.uint32), want_data.view(np.uint32))
    assert torch.equal(mk.maxk_forward(dev(x), k), data)          # reference signature: values only


def test_topk_ties_nan_signed_zero(mk):
    from oracle import maxk_oracle as mo
    nan, inf = np.nan, np.inf
    rows = [
        [1, 3, 3, 3, 0, 3, -1, 2], [0] * 8, [-0.0, 0.0, -0.0, 0.0, -1, -1, -1, -1],
        [nan, 1, inf, nan, -inf, 5, 4, nan], [-inf, -inf, -inf, -5, -inf, -inf, -inf, -inf],
    ]
    x = np.array(rows, dtype=np.float32)
    for k in (1, 2, 3, 5, 8):
        data, idx = mk.maxk_forward_cbsr(dev(x), k)
        wd, wi = mo.maxk_cbsr(x, k)
        assert np.array_equal(idx.cpu().numpy(), wi), k
        assert np.array_equal(data.cpu().numpy().view(np.uint32), wd.view(np.uint32)), k
    # heavy ties at realistic width: quantised values, many duplicates of the threshold
    rng = np.random.default_rng(0)
    for d, k in ((256, 32), (384, 16), (1500, 64)):
        xq = np.round(rng.standard_normal((64, d)) * 2).astype(np.float32) / 2
        xq[3] = 0.0
        xq[5, ::3] = np.nan
        data, idx = mk.maxk_forward_cbsr(dev(xq), k)
        wd, wi = mo.maxk_cbsr(xq, k)
        gi = idx.cpu().numpy() if d <= 256 else idx.view(torch.int16).cpu().numpy().view(np.uint16)
        assert np.array_equal(gi, wi), (d, k)
        assert np.array_equal(data.cpu().numpy().view(np.uint32), wd.view(np.uint32)), (d, k)


def _tie_heavy_rows(rng, n, d):
    """Row families that stress the interpolation search of topk_tile.cu: thresholds inside big tie
    groups (ReLU-like zeros, clipped maxima, integers, constants), extreme scales, infinities."""
    base = rng.standard_normal((n, d)).astype(np.float32)
    fam = [
        base, np.maximum(base, 0), np.maximum(base - 1.5, 0), np.minimum(base, 0.5),
        rng.integers(-3, 4, (n, d)).astype(np.float32), np.ones((n, d), np.float32),
        (base * 1e-30).astype(np.float32), (base * np.exp(rng.standard_normal((n, d)) * 5)).astype(np.float32),
        -np.abs(base), np.where(base > 1.0, np.inf, base).astype(np.float32),
        np.where(base < -1.0, -np.inf, np.float32(-0.0) * base).astype(np.float32),
        rng.standard_cauchy((n, d)).astype(np.float32), np.float32(1e-42) * rng.integers(0, 5, (n, d)).astype(np.float32),
    ]
    return np.concatenate(fam, axis=0)


@pytest.mark.parametrize("d,k", [(256, 32), (256, 8), (256, 64), (128, 32), (100, 7), (384, 16), (512, 64),
                                 (1024, 40), (64, 64), (32, 1)])
def test_topk_tile_kernel_on_tie_heavy_rows(mk, d, k):
    """Second-generation top-k kernel (interpolation search, tie short cuts, NaN fallback) against the
    oracle, bit for bit, on rows built to break a value-space search."""
    from oracle import c_oracle
    rng = np.random.default_rng(d * 100 + k)
    x = _tie_heavy_rows(rng, 40, d)
    x[7, : d // 2] = np.nan
    x[11, 3] = np.nan
    wd, wi = c_oracle.maxk_cbsr(x, k)
    outs = [mk.maxk_forward_cbsr(dev(x), k)]                         # topk.cu (bitwise search)
    if mk.banked_supported(k, d):                                    # topk_tile.cu (interpolation search)
        outs.append(mk.maxk_forward_cbsr_banked(dev(x), k, want_data=True)[:2])
    for data, idx in outs:
        gi = idx.cpu().numpy() if d <= 256 else idx.view(torch.int16).cpu().numpy().view(np.uint16)
        assert np.array_equal(gi, wi)
        assert np.array_equal(data.cpu().numpy().view(np.uint32), wd.view(np.uint32))


@pytest.mark.parametrize("n,d,k", [(1000, 256, 32), (999, 256, 64), (4097, 256, 16), (333, 128, 8), (257, 384, 32),
                                   (100, 512, 64), (200_000, 256, 32), (31, 64, 8)])
def test_fused_topk_and_banking_is_bit_identical_to_the_two_kernels(mk, n, d, k):
    """mk_topk_cbsr_bank (f-3) == mk_topk_cbsr followed by mk_cbsr_bank / mk_cbsr_bank_packed, bit for
    bit: sorted column ids, sorted values, banked values, cell offsets -- 200,000 rows make every warp
    walk more than one 32-row tile."""
    rng = np.random.default_rng(n + d + k)
    x = rng.standard_normal((n, d)).astype(np.float32)
    if n < 5000:
        x[: n // 3] = np.maximum(x[: n // 3] - 1.0, 0)          # ties inside the tile
    xt = dev(x)
    sd, si = mk.maxk_forward_cbsr(xt, k)
    bd, _, bs = mk.cbsr_bank(sd, si, d, with_index=False)
    fd, fi, fbd, fbs = mk.maxk_forward_cbsr_banked(xt, k, want_data=True)
    assert torch.equal(fi.view(torch.uint8), si.view(torch.uint8))
    assert torch.equal(fd.view(torch.int32), sd.view(torch.int32))
    assert torch.equal(fbd.view(torch.int32), bd.view(torch.int32))
    assert torch.equal(fbs, bs)
    nd, ni, _, _ = mk.maxk_forward_cbsr_banked(xt, k)           # sorted values not requested
    assert nd is None and torch.equal(ni.view(torch.uint8), si.view(torch.uint8))
    if mk.packed_supported(k, d):
        want = mk.cbsr_bank_packed(sd, si, d)
        _, pi, pk, none = mk.maxk_forward_cbsr_banked(xt, k, packed=True)
        assert none is None and torch.equal(pk, want) and torch.equal(pi.view(torch.uint8), si.view(torch.uint8))


@pytest.mark.parametrize("n,deg,d,k,max_nz,with_self,with_bias", [
    (2500, 120, 256, 32, 1024, True, False), (2500, 120, 256, 32, 64, True, True), (1500, 150, 256, 16, 1024, True, False),
    (1500, 150, 128, 8, 32, False, True), (900, 200, 512, 64, 1024, True, True), (700, 90, 384, 32, 1024, False, False)])
def test_forward_with_layernorm_epilogue_is_bit_identical(mk, n, deg, d, k, max_nz, with_self, with_bias):
    """mk_spgemm_fwd_banked_ln (f-3: y = LayerNorm(h_self + A x Xs + bias) inside the forward SpGEMM) ==
    the banked / packed forward followed by mk_add_layernorm_fwd, bit for bit -- y, z, mean, rstd --
    rows of several records (fold + epilogue kernel) included; inference form without z."""
    from conftest import small_graph
    g = small_graph(n, deg, seed=n + k, device="cuda")
    e = g.num_edges()
    rng = np.random.default_rng(d + k)
    x = dev(rng.standard_normal((n, d)).astype(np.float32))
    hs = dev(rng.standard_normal((n, d)).astype(np.float32)) if with_self else None
    bias = dev(rng.standard_normal(d).astype(np.float32)) if with_bias else None
    gamma = dev(rng.standard_normal(d).astype(np.float32))
    beta = dev(rng.standard_normal(d).astype(np.float32))
    val = g.edge_weights("mean")
    sd, si = mk.maxk_forward_cbsr(x, k)
    mk.set_max_nz(max_nz)
    try:
        mk.clear_partition_cache()
        if k >= 32:
            table, _, slot = mk.cbsr_bank(sd, si, d, with_index=False)
            agg = mk.spgemm_forward_banked(g.indptr, g.indices, val, table, slot, n, e, k, d)
        else:
            table, slot = mk.cbsr_bank_packed(sd, si, d), None
            agg = mk.spgemm_forward_packed(g.indptr, g.indices, val, table, n, e, k, d)
        if with_self:
            wy, wz, wm, wr = mk.add_layernorm_forward(hs, agg, bias, gamma, beta, 1e-5)
        else:
            wy, wz, wm, wr = mk.add_layernorm_forward(agg, None, bias, gamma, beta, 1e-5)
        y, z, m, r = mk.spgemm_forward_ln(g.indptr, g.indices, val, table, slot, n, e, k, d, hs, bias, gamma, beta, 1e-5)
        assert torch.equal(z, wz) and torch.equal(m, wm) and torch.equal(r, wr) and torch.equal(y, wy)
        y2, z2, m2, r2 = mk.spgemm_forward_ln(g.indptr, g.indices, val, table, slot, n, e, k, d, hs, bias, gamma,
                                              beta, 1e-5, keep_stats=False)
        assert z2 is None and m2 is None and r2 is None and torch.equal(y2, wy)
    finally:
        mk.set_max_nz(1024)
        mk.clear_partition_cache()


def test_sage_layer_with_the_fused_epilogue_equals_the_separate_kernels(mk):
    """MaxKSAGEConv through MaxKAggregateLNFunction (epilogue inside the SpGEMM) against the same layer
    on the separate kernels: forward bit-equal, parameter and input gradients to summation order."""
    from conftest import small_graph
    from spgemm_gnn_b200 import maxk_layers as ml
    g = small_graph(3000, 110, seed=12, device="cuda")
    torch.manual_seed(5)
    layer = ml.MaxKSAGEConv(256, 256, "mean", feat_drop=0.0, norm=torch.nn.LayerNorm(256), maxk=32).cuda()
    x0 = torch.randn(g.num_nodes(), 256, device="cuda")
    dy = torch.randn(g.num_nodes(), 256, device="cuda")
    res = {}
    for fused in (True, False):
        was = ml.FUSED_LN_EPILOGUE
        ml.FUSED_LN_EPILOGUE = fused
        try:
            layer.zero_grad()
            x = x0.clone().requires_grad_(True)
            y = layer(g, x)
            y.backward(dy)
            res[fused] = (y.detach(), x.grad.clone(), [p.grad.clone() for p in layer.parameters()],
                          type(y.grad_fn).__name__)
        finally:
            ml.FUSED_LN_EPILOGUE = was
    assert res[True][3].startswith("MaxKAggregateLNFunction") and not res[False][3].startswith("MaxKAggregateLN")
    assert torch.equal(res[True][0], res[False][0])
    for a, b in [(res[True][1], res[False][1])] + list(zip(res[True][2], res[False][2])):
        assert float((a - b).abs().max() / b.abs().max().clamp_min(1e-20)) < 2e-5


@pytest.mark.parametrize("k", [32, 16])
def test_fused_aggregate_function_equals_the_unfused_path(mk, k):
    """`maxk_aggregate` through MaxKAggregateFunction (top-k + banking fused) gives the forward and the
    input gradient of MaxKCBSRFunction + SpGEMMFunction."""
    from conftest import small_graph
    from spgemm_gnn_b200 import maxk_layers as ml
    g = small_graph(3000, 120, seed=4, device="cuda")
    rng = np.random.default_rng(k)
    x = dev(rng.standard_normal((g.num_nodes(), 256)).astype(np.float32))
    dy = dev(rng.standard_normal((g.num_nodes(), 256)).astype(np.float32))
    a = x.clone().requires_grad_(True)
    was = ml.FUSED_TOPK_BANK
    ml.FUSED_TOPK_BANK = True
    try:
        ya = ml.maxk_aggregate(g, a, k, "mean")
    finally:
        ml.FUSED_TOPK_BANK = was
    assert type(ya.grad_fn).__name__.startswith("MaxKAggregateFunction")
    ya.backward(dy)
    b = x.clone().requires_grad_(True)
    sd, si = ml.MaxKCBSRFunction.apply(b, k)
    yb = ml.aggregate_cbsr(g, sd, si, "mean", 256)
    yb.backward(dy)
    assert torch.equal(ya, yb)
    scale = b.grad.abs().max()
    assert float((a.grad - b.grad).abs().max() / scale) < 1e-5   # REDG order is free in both
    assert torch.equal(a.grad != 0, b.grad != 0)


def test_golden_vectors_through_the_cuda_path(mk, golden):
    """Outputs of the reference's own Python (tests/golden/make_golden.py)."""
    from spgemm_gnn_b200.maxk_layers import MaxKFunction
    for ci in range(int(golden["num_cases"])):
        pre = f"c{ci}_"
        x, k = golden[pre + "x"], int(golden[pre + "k"])
        xt = dev(x).requires_grad_(True)
        y = MaxKFunction.apply(xt, k)                       # utils/maxk_layers.py::MaxKFunction
        assert np.array_equal(y.detach().cpu().numpy(), golden[pre + "maxk_out"])
        y.backward(dev(golden[pre + "grad_out"]))
        assert np.array_equal(xt.grad.cpu().numpy(), golden[pre + "maxk_grad_in"])
        if pre + "sp_index" in golden.files:              # _extract_sparse_format layout
            data, idx = mk.maxk_forward_cbsr(dev(x), k)
            assert np.array_equal(idx.cpu().numpy(), golden[pre + "sp_index"])
            assert np.array_equal(data.cpu().numpy(), golden[pre + "sp_data"])


@pytest.mark.parametrize("n,d,k", [(100, 256, 32), (33, 100, 7), (20, 384, 16), (7, 1500, 40), (5, 9000, 33),
                                   (3, 49152, 64), (2, 51204, 9), (2, 65536, 40)])
def test_scatter_and_gather(mk, n, d, k):
    from oracle import maxk_oracle as mo
    rng = np.random.default_rng(d)
    x = rng.standard_normal((n, d)).astype(np.float32)
    g = rng.standard_normal((n, k)).astype(np.float32)
    _, wi = mo.maxk_cbsr(x, k)
    idx = dev(wi.astype(np.uint8)) if d <= 256 else dev(wi.view(np.int16)).view(torch.uint16)
    dense = mk.cbsr_scatter(dev(g), idx, d)
    assert np.array_equal(dense.cpu().numpy(), mo.cbsr_scatter(g, wi, d))
    assert np.array_equal(mk.cbsr_gather(dev(x), idx).cpu().numpy(), mo.cbsr_gather(x, wi))
    # reference signature maxk_backward(grad_output, indices): D inferred as indices.max()+1,
    # int64 indices accepted (what utils/maxk_layers.py:23 saves)
    d_inf = max(int(wi.max()) + 1, k)
    got = mk.maxk_backward(dev(g), dev(wi.astype(np.int64)))
    assert got.shape == (n, d_inf)
    assert np.array_equal(got.cpu().numpy(), mo.cbsr_scatter(g, wi, d)[:, :d_inf])


# ---------------------------------------------------------------------------------------
# a-5  partition
# ---------------------------------------------------------------------------------------
def _edge_case_ptr():
    deg = np.array([0, 1, 63, 64, 65, 0, 0, 128, 129, 20000, 1, 0, 5000, 7], dtype=np.int64)
    rng = np.random.default_rng(1)
    deg = np.concatenate([deg, rng.integers(0, 300, 3000)])
    return np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)


@pytest.mark.parametrize("max_nz", [1, 64, 100, 1024, 1 << 20])
def test_partition_matches_oracle(mk, max_nz):
    from oracle import c_oracle
    ptr = _edge_case_ptr()
    if max_nz == 1:
        ptr = ptr[:200]
        ptr = ptr - ptr[0]
    want, slots = c_oracle.partition_rows(ptr, max_nz)
    p = mk.partition(dev(ptr), ptr.size - 1, max_nz)
    assert (p.num_parts, p.num_slots) == (len(want), slots)
    assert np.array_equal(p.parts.cpu().numpy()[: p.num_parts], want)


# ---------------------------------------------------------------------------------------
# a-3 / a-4  SpGEMM forward, SSpMM backward
# ---------------------------------------------------------------------------------------
def _problem(n, avg_deg, d, k, seed, kind="mean", extra_ptr=None):
    from oracle import maxk_oracle as mo
    from spgemm_gnn_b200.graph import synthetic_graph
    g = synthetic_graph(n, n * avg_deg, seed=seed)
    ptr, idx = g.indptr.numpy(), g.indices.numpy()
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d)).astype(np.float32)
    dy = rng.standard_normal((n, d)).astype(np.float32)
    val = mo.edge_weights(ptr, idx, kind)
    return ptr, idx, val, x, dy


def _to_index_tensor(wi, d):
    return dev(wi) if d <= 256 else dev(wi.view(np.int16)).view(torch.uint16)


LAYER_CASES = [
    # n, avg_deg, d, k, max_nz
    (2000, 40, 256, 32, 1024), (2000, 40, 256, 16, 1024), (2000, 40, 256, 8, 1024),
    (2000, 40, 256, 64, 1024), (1500, 30, 256, 32, 64), (1500, 30, 64, 16, 7),
    (1000, 20, 384, 16, 256), (1000, 20, 384, 64, 256), (800, 25, 256, 7, 128),
    (800, 25, 100, 20, 128), (600, 10, 512, 128, 1024), (600, 10, 96, 96, 1024),
    (500, 12, 1500, 40, 200), (700, 15, 128, 4, 50),
]


@pytest.mark.parametrize("n,avg_deg,d,k,max_nz", LAYER_CASES)
@pytest.mark.parametrize("kind", ["mean", "both"])
def test_spgemm_forward_and_sspmm_backward(mk, n, avg_deg, d, k, max_nz, kind):
    from oracle import c_oracle, maxk_oracle as mo
    ptr, idx, val, x, dy = _problem(n, avg_deg, d, k, seed=n + d + k, kind=kind)
    wd, wi = c_oracle.maxk_cbsr(x, k)
    mk.set_max_nz(max_nz)
    try:
        tptr, tidx, tval = dev(ptr), dev(idx), dev(val)
        sp_data, sp_index = dev(wd), _to_index_tensor(wi, d)
        out, ret_index = mk.spgemm_forward(tptr, tidx, tval, sp_data, sp_index, n, idx.size, k, d)
        assert ret_index is sp_index and out.shape == (n, d) and out.dtype == torch.float32
        want = c_oracle.spgemm_fwd(ptr, idx, val, wd, wi, d)
        bound = c_oracle.spgemm_fwd(ptr, idx, np.abs(val), np.abs(wd), wi, d)
        assert_rel(out, want, bound, "spgemm_forward")
        # no atomics on the forward path: bit-reproducible
        out2, _ = mk.spgemm_forward(tptr, tidx, tval, sp_data, sp_index, n, idx.size, k, d)
        assert torch.equal(out, out2)

        dxs = mk.spgemm_backward(tptr, tidx, tval, dev(dy), sp_index, n, idx.size, k, d)
        assert dxs.shape == (n, k)
        want_b = c_oracle.sspmm_bwd(ptr, idx, val, dy, wi)
        bound_b = c_oracle.sspmm_bwd(ptr, idx, np.abs(val), np.abs(dy), wi)
        assert_rel(dxs, want_b, bound_b, "spgemm_backward")
    finally:
        mk.set_max_nz(1024)


@pytest.mark.parametrize("nt", [1, 2, 4])
def test_backward_bulk_reduction_variant_and_out_buffer(mk, nt):
    """The experimental backward (part of the reductions as bulk shared->global reductions) and the
    `out=` form of spgemm_backward compute what the shipped kernel computes."""
    from oracle import c_oracle
    n, d, k = 3000, 256, 32
    ptr, idx, val, x, dy = _problem(n, 60, d, k, seed=nt, kind="mean")
    wd, wi = c_oracle.maxk_cbsr(x, k)
    tptr, tidx, tval, sp_index = dev(ptr), dev(idx), dev(val), dev(wi)
    want_b = c_oracle.sspmm_bwd(ptr, idx, val, dy, wi)
    bound_b = c_oracle.sspmm_bwd(ptr, idx, np.abs(val), np.abs(dy), wi)
    buf = torch.full((n, k), float("nan"), device="cuda")
    mk.set_backward_tma(nt)
    try:
        dxs = mk.spgemm_backward(tptr, tidx, tval, dev(dy), sp_index, n, idx.size, k, d, out=buf)
    finally:
        mk.set_backward_tma(0)
    assert dxs is buf
    assert_rel(dxs, want_b, bound_b, f"spgemm_backward tma={nt}")
    with pytest.raises(RuntimeError, match="out must be float32"):
        mk.spgemm_backward(tptr, tidx, tval, dev(dy), sp_index, n, idx.size, k, d, out=buf[:, :8].contiguous())


def test_degree_edge_cases_and_empty_rows(mk):
    """degrees 0 / 1 / 64 / 65 / > 10^4 in one graph (SURVEY.md section 4 implication)."""
    from oracle import c_oracle
    rng = np.random.default_rng(4)
    n, d, k = 12000, 256, 32
    deg = np.zeros(n, dtype=np.int64)
    deg[:8] = [0, 1, 64, 65, 11000, 0, 1025, 2048]
    deg[8:] = rng.integers(0, 6, n - 8)
    ptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int32)
    idx = np.concatenate([np.sort(rng.choice(n, dg, replace=False)) for dg in deg]).astype(np.int32)
    val = rng.standard_normal(idx.size).astype(np.float32)
    x = rng.standard_normal((n, d)).astype(np.float32)
    dy = rng.standard_normal((n, d)).astype(np.float32)
    wd, wi = c_oracle.maxk_cbsr(x, k)
    for max_nz in (64, 1024):
        mk.set_max_nz(max_nz)
        try:
            out, _ = mk.spgemm_forward(dev(ptr), dev(idx), dev(val), dev(wd), dev(wi), n, idx.size, k, d)
            want = c_oracle.spgemm_fwd(ptr, idx, val, wd, wi, d)
            bound = c_oracle.spgemm_fwd(ptr, idx, np.abs(val), np.abs(wd), wi, d)
            assert_rel(out, want, bound, f"fwd max_nz={max_nz}")
            assert float(out[0].abs().max()) == 0.0 and float(out[5].abs().max()) == 0.0
            dxs = mk.spgemm_backward(dev(ptr), dev(idx), dev(val), dev(dy), dev(wi), n, idx.size, k, d)
            want_b = c_oracle.sspmm_bwd(ptr, idx, val, dy, wi)
            bound_b = c_oracle.sspmm_bwd(ptr, idx, np.abs(val), np.abs(dy), wi)
            assert_rel(dxs, want_b, bound_b, f"bwd max_nz={max_nz}")
        finally:
            mk.set_max_nz(1024)


def test_identity_adjacency_and_reference_padding(mk, golden):
    """A = I  =>  Y == dense(Xs) exactly; zero-padded CBSR rows (utils/maxk_layers.py:245-257)
    produce the plain dense view (they race in the reference kernel)."""
    from oracle import maxk_oracle as mo
    x = golden["pad_x"]
    n, d = x.shape
    k = int(golden["pad_k"])
    ptr = dev(np.arange(n + 1, dtype=np.int32))
    idx = dev(np.arange(n, dtype=np.int32))
    val = torch.ones(n, device="cuda")
    out, _ = mk.spgemm_forward(ptr, idx, val, dev(golden["pad_sp_data"]), dev(golden["pad_sp_index"]),
                               n, n, k, d)
    want = mo.cbsr_to_dense(golden["pad_sp_data"], golden["pad_sp_index"], d)
    assert np.array_equal(out.cpu().numpy().astype(np.float64), want)


def test_reference_layer_call_sites_through_the_cuda_path(mk, golden_layers):
    """The arguments the reference's own MaxKSAGEConv / MaxKGCNConv code passes to
    `maxk_kernels.spgemm_forward` (recorded by tests/golden/make_golden_layers.py), through the CUDA
    entry point of the same name: + the layer's epilogue == what the layer computes through
    `graph.update_all` in the reference."""
    from conftest import layer_cases
    from oracle import c_oracle
    from spgemm_gnn_b200.graph import CSRGraph
    from spgemm_gnn_b200.maxk_layers import aggregate_cbsr
    for name, c, y, add in layer_cases(golden_layers):
        args = (dev(c["ptr"]), dev(c["idx"]), dev(c["val"]), dev(c["sp_data"]), dev(c["sp_index"]))
        out, _ = mk.spgemm_forward(*args, c["n"], c["e"], c["k"], c["d"])
        bound = c_oracle.spgemm_fwd(c["ptr"], c["idx"], np.abs(c["val"]), np.abs(c["sp_data"]),
                                    c["sp_index"], c["d"]) + np.abs(add)
        # the golden output is float32 from float64 sums: half an ulp of y on top of the kernel bar
        assert_rel(out + dev(np.broadcast_to(add, y.shape).astype(np.float32)), y.astype(np.float64),
                   bound + 0.01 * np.abs(y), name)
        if name == "gcn_both":
            # the reference's MaxKGCNConv as a whole (both degree factors on the source node): this repo's
            # `reference_gcn` edge weights applied to the un-scaled MaxK output give the layer's output
            g = CSRGraph(args[0], args[1])
            do = g.out_degrees().clamp(min=1).float().pow(0.5)            # undo the feature scaling of :315-318
            raw = (args[3] * do[:, None]).contiguous()
            agg = aggregate_cbsr(g, raw, args[4], "reference_gcn", c["d"])
            assert float((agg - out).abs().max()) <= 1e-5 * float(out.abs().max())
        if name.startswith("sage"):   # the layer's own graph code builds the same weights (a-7)
            g = CSRGraph(args[0], args[1])
            kind = "sum" if name == "sage_sum" else "mean"
            agg = aggregate_cbsr(g, args[3], args[4], kind, c["d"])
            assert torch.equal(agg, out) or torch.allclose(agg, out, rtol=1e-6, atol=0)
            # backward of the reference layer (autograd through update_all, then grad * mask):
            # spgemm_backward + CBSR scatter == the gradient at the input of the MaxK
            dy, want = golden_layers[f"{name}_dy"], golden_layers[f"{name}_grad_maxk_in"]
            dxs = mk.spgemm_backward(args[0], args[1], args[2], dev(dy), args[4], c["n"], c["e"],
                                     c["k"], c["d"])
            bound_b = c_oracle.sspmm_bwd(c["ptr"], c["idx"], np.abs(c["val"]), np.abs(dy), c["sp_index"])
            want_k = np.take_along_axis(want, c["sp_index"].astype(np.int64), axis=1)
            assert_rel(dxs, want_k.astype(np.float64), bound_b + 0.01 * np.abs(want_k), name + " backward")
            dense = mk.cbsr_scatter(dxs, args[4], c["d"])
            assert int((dense != 0).sum()) <= c["n"] * c["k"]
            assert torch.equal(mk.cbsr_gather(dense, args[4]), dxs)


def test_gin_layer_matches_the_reference_class_end_to_end(mk, golden_layers):
    """This repo's MaxKGINConv with the reference's state dict against the output of the reference's
    own class (utils/integrated_models.py:221-270, run by tests/golden/make_golden_layers.py): MaxK on
    the raw features (bit-exact selection), sum over in-neighbours on the kernels, then the MLP."""
    from spgemm_gnn_b200.graph import CSRGraph
    from spgemm_gnn_b200.maxk_layers import MaxKGINConv
    gl = golden_layers
    n, d_in, d_out, k = (int(v) for v in gl["gin_dims"])
    conv = MaxKGINConv(d_in, d_out, learn_eps=True, maxk=k)
    conv.load_state_dict({key[len("gin_sd_"):]: torch.from_numpy(gl[key]) for key in gl.files
                          if key.startswith("gin_sd_")})
    conv = conv.cuda().eval()
    g = CSRGraph(dev(gl["gin_ptr"]), dev(gl["gin_idx"]))
    seen = []
    hook = conv.mlp.register_forward_pre_hook(lambda _m, inp: seen.append(inp[0].detach()))
    tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False      # the MLP is compared in plain fp32
    try:
        with torch.no_grad():
            y = conv(g, dev(gl["gin_feat"]))
    finally:
        torch.backends.cuda.matmul.allow_tf32 = tf32
        hook.remove()
    want_pre, want_y = gl["gin_pre_mlp"], gl["gin_y"]
    assert y.shape == (n, d_out) and len(seen) == 1
    np.testing.assert_allclose(seen[0].cpu().numpy(), want_pre, rtol=0, atol=1e-5 * np.abs(want_pre).max())
    # two fp32 GEMMs (K = 128, 64) on different hardware: summation order only
    np.testing.assert_allclose(y.cpu().numpy(), want_y, rtol=0, atol=1e-4 * np.abs(want_y).max())


def test_sage_layer_matches_the_reference_class_end_to_end(mk, golden_layers):
    """This repo's MaxKSAGEConv with the reference's state dict against the output of the reference's
    own class (utils/maxk_layers.py:47-222).  The weights are scaled permutation matrices, so both
    Linear layers are exact on any hardware and the MaxK selects the same entries as on the CPU."""
    from spgemm_gnn_b200.graph import CSRGraph
    from spgemm_gnn_b200.maxk_layers import MaxKSAGEConv
    gl = golden_layers
    n, d, k = (int(v) for v in gl["sagex_dims"])
    conv = MaxKSAGEConv(d, d, aggregator_type="mean", maxk=k)
    conv.load_state_dict({key[len("sagex_sd_"):]: torch.from_numpy(gl[key]) for key in gl.files
                          if key.startswith("sagex_sd_")})
    conv = conv.cuda().eval()
    g = CSRGraph(dev(gl["sagex_ptr"]), dev(gl["sagex_idx"]))
    tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False      # exactness of the two Linear layers needs fp32
    try:
        with torch.no_grad():
            y = conv(g, dev(gl["sagex_feat"]))
    finally:
        torch.backends.cuda.matmul.allow_tf32 = tf32
    want = gl["sagex_y"]
    assert y.shape == (n, d)
    np.testing.assert_allclose(y.cpu().numpy(), want, rtol=0, atol=1e-5 * np.abs(want).max())


def test_rectangular_shard_with_global_columns(mk):
    """Row shard of a bigger graph: n_rows < n_src (the 1-D partition of SURVEY.md section 8e)."""
    from oracle import c_oracle, maxk_oracle as mo
    from spgemm_gnn_b200.graph import synthetic_graph
    g = synthetic_graph(3000, 90000, seed=8)
    s = g.row_slice(1000, 1700)
    ptr, idx = s.indptr.numpy(), s.indices.numpy()
    rng = np.random.default_rng(8)
    x = rng.standard_normal((3000, 256)).astype(np.float32)
    dy = rng.standard_normal((700, 256)).astype(np.float32)
    val = mo.edge_weights(ptr, idx, "mean", num_src=3000)
    wd, wi = c_oracle.maxk_cbsr(x, 32)
    out, _ = mk.spgemm_forward(dev(ptr), dev(idx), dev(val), dev(wd), dev(wi), 700, idx.size, 32, 256)
    assert_rel(out, c_oracle.spgemm_fwd(ptr, idx, val, wd, wi, 256),
               c_oracle.spgemm_fwd(ptr, idx, np.abs(val), np.abs(wd), wi, 256), "shard fwd")
    dxs = mk.spgemm_backward(dev(ptr), dev(idx), dev(val), dev(dy), dev(wi), 700, idx.size, 32, 256)
    assert dxs.shape == (3000, 32)
    assert_rel(dxs, c_oracle.sspmm_bwd(ptr, idx, val, dy, wi),
               c_oracle.sspmm_bwd(ptr, idx, np.abs(val), np.abs(dy), wi), "shard bwd")


def test_error_behaviour_matches_reference_binding(mk):
    x = torch.randn(8, 32, device="cuda")
    with pytest.raises(RuntimeError, match="k must be between 1 and input dimension"):
        mk.maxk_forward(x, 0)
    with pytest.raises(RuntimeError, match="k must be between 1 and input dimension"):
        mk.maxk_forward(x, 33)
    with pytest.raises(RuntimeError, match="Input must be 2D tensor"):
        mk.maxk_forward(x.view(-1), 4)
    with pytest.raises(RuntimeError, match="input must be contiguous"):
        mk.maxk_forward(x.t(), 4)
    ptr = torch.arange(9, dtype=torch.int64, device="cuda")
    idx = torch.arange(8, dtype=torch.int32, device="cuda")
    val = torch.ones(8, device="cuda")
    d, i = mk.maxk_forward_cbsr(x, 4)
    with pytest.raises(RuntimeError, match="ptr must be int32"):
        mk.spgemm_forward(ptr, idx, val, d, i, 8, 8, 4, 32)
    with pytest.raises(RuntimeError, match="val must be float32"):
        mk.spgemm_forward(ptr.int(), idx, val.double(), d, i, 8, 8, 4, 32)
    with pytest.raises(RuntimeError, match="sp_data must be float32"):
        mk.spgemm_forward(ptr.int(), idx, val, d.half(), i, 8, 8, 4, 32)
    with pytest.raises(RuntimeError, match="grad_output must be float32"):
        mk.spgemm_backward(ptr.int(), idx, val, x.half(), i, 8, 8, 4, 32)


# ---------------------------------------------------------------------------------------
# layers against the torch restatement of the reference's training path
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["mean", "both", "sum"])
def test_autograd_layer_matches_torch_reference(mk, kind):
    from oracle import ref_torch
    from spgemm_gnn_b200.graph import synthetic_graph
    from spgemm_gnn_b200.maxk_layers import maxk_aggregate
    g = synthetic_graph(1500, 45000, seed=21)
    gen = torch.Generator().manual_seed(21)
    x = torch.randn(1500, 256, generator=gen)
    dy = torch.randn(1500, 256, generator=gen)
    adj = ref_torch.csr_matrix(g.indptr, g.indices, g.edge_weights(kind), g.num_src)
    y_ref, dx_ref = ref_torch.layer_forward_backward(adj.to(torch.float64), x.double(), dy.double(), 32)
    gg = g.to("cuda")
    xc = x.cuda().requires_grad_(True)
    y = maxk_aggregate(gg, xc, 32, kind)
    y.backward(dy.cuda())
    scale_y = float(y_ref.abs().max())
    scale_g = float(dx_ref.abs().max())
    assert float((y.cpu().double() - y_ref).abs().max()) <= 1e-5 * scale_y
    assert float((xc.grad.cpu().double() - dx_ref).abs().max()) <= 1e-5 * scale_g
    # the gradient is zero exactly where the reference's mask is zero
    assert torch.equal(xc.grad.cpu() == 0, dx_ref == 0)



# ---------------------------------------------------------------------------------------
# full BASELINE size: properties that need no CPU oracle, plus an oracle check on a row sample
# ---------------------------------------------------------------------------------------
def test_reddit_shape_properties(mk):
    """Reddit-shaped synthetic graph (232,965 nodes, ~114M edges), D=256, k=32: adjointness
    <A Xs, dY> == <Xs, dXs>, exact linearity in val (power-of-two scaling), agreement with the
    dense cuSPARSE SpMM on the masked matrix, bit-reproducible forward."""
    from spgemm_gnn_b200.graph import shaped_graph
    g = shaped_graph("reddit", device="cuda")
    n, e = g.num_nodes(), g.num_edges()
    assert n == 232965 and 0.9e8 < e < 1.3e8
    gen = torch.Generator(device="cuda").manual_seed(97)
    x = torch.randn(n, 256, device="cuda", generator=gen)
    dy = torch.randn(n, 256, device="cuda", generator=gen)
    val = g.edge_weights("mean")
    sp_data, sp_index = mk.maxk_forward_cbsr(x, 32)
    # top-k: ascending distinct columns, values are the k largest of the row
    cols = sp_index.to(torch.int32)
    assert bool((cols[:, 1:] > cols[:, :-1]).all())
    kth = torch.topk(x, 32, dim=1)[0][:, -1]
    assert bool((sp_data.min(dim=1)[0] == kth).all())
    assert torch.equal(torch.gather(x, 1, cols.long()), sp_data)

    out, _ = mk.spgemm_forward(g.indptr, g.indices, val, sp_data, sp_index, n, e, 32, 256)
    out2, _ = mk.spgemm_forward(g.indptr, g.indices, val, sp_data, sp_index, n, e, 32, 256)
    assert torch.equal(out, out2)
    out4, _ = mk.spgemm_forward(g.indptr, g.indices, val * 4.0, sp_data, sp_index, n, e, 32, 256)
    assert torch.equal(out4, out * 4.0)
    dxs = mk.spgemm_backward(g.indptr, g.indices, val, dy, sp_index, n, e, 32, 256)
    lhs = float((out.double() * dy.double()).sum())
    rhs = float((sp_data.double() * dxs.double()).sum())
    # both sides are sums of ~6e7 signed terms: compare against the sum of their magnitudes
    scale = float((out.double().abs() * dy.double().abs()).sum())
    assert abs(lhs - rhs) <= 1e-8 * scale
    oracle_sample_check(g, val, sp_data, sp_index, out, dy, dxs, 256, 32)

    # mean aggregation of a constant CBSR table reproduces the constant (rows of A sum to 1)
    ones = torch.ones_like(sp_data)
    same_idx = sp_index[:1].expand(n, 32).contiguous()
    o1, _ = mk.spgemm_forward(g.indptr, g.indices, val, ones, same_idx, n, e, 32, 256)
    picked = o1[:, same_idx[0].long()]
    assert float((picked - 1.0).abs().max()) < 1e-4
    assert float(o1.sum()) == pytest.approx(32.0 * n, rel=1e-4)

    # independent check at full size: cuSPARSE dense SpMM on the masked matrix
    xm = mk.cbsr_scatter(sp_data, sp_index, 256)
    adj = torch.sparse_csr_tensor(g.indptr.long(), g.indices.long(), val, size=(n, n))
    ref = torch.sparse.mm(adj, xm)
    assert float((out - ref).abs().max()) <= 2e-5 * float(ref.abs().max())
    gref = torch.sparse.mm(adj.t().to_sparse_csr(), dy)
    gpick = torch.gather(gref, 1, cols.long())
    assert float((dxs - gpick).abs().max()) <= 2e-5 * float(gpick.abs().max())


# ---------------------------------------------------------------------------------------
# banked CBSR path (product-internal conflict-free format)
# ---------------------------------------------------------------------------------------
def _np_index(t, d):
    return t.cpu().numpy() if d <= 256 else t.view(torch.int16).cpu().numpy().view(np.uint16)


@pytest.mark.parametrize("n,d,k", [(500, 256, 32), (300, 256, 16), (300, 256, 8), (200, 256, 64),
                                   (200, 64, 8), (150, 384, 16), (100, 512, 64), (120, 128, 32)])
def test_banked_form_is_a_valid_permutation_with_few_conflicts(mk, n, d, k):
    from oracle import c_oracle, maxk_oracle as mo
    rng = np.random.default_rng(d + k)
    x = rng.standard_normal((n, d)).astype(np.float32)
    wd, wi = c_oracle.maxk_cbsr(x, k)
    bd, bi, bs = mk.cbsr_bank(dev(wd), _to_index_tensor(wi, d), d)
    mean_wf, _ = mo.check_banked(wd, wi, bd.cpu().numpy(), _np_index(bi, d),
                                 bs.cpu().numpy().view(np.uint16), d)
    # unbanked layouts see ~3.5 wavefronts per access on these inputs
    assert mean_wf < (1.6 if k >= 32 else 2.4), mean_wf
    # padded / duplicate rows must not break the permutation property
    zd = np.zeros((4, k), np.float32)
    zi = np.zeros((4, k), wi.dtype)
    bd, bi, bs = mk.cbsr_bank(dev(zd), _to_index_tensor(zi, d), d)
    assert float(bd.abs().max()) == 0.0 and int(bi.to(torch.int32).max()) == 0


BANKED_CASES = [(2000, 40, 256, 32, 1024), (2000, 40, 256, 16, 1024), (2000, 40, 256, 8, 1024),
                (1500, 30, 256, 64, 64), (1000, 20, 384, 16, 256), (800, 25, 64, 8, 7),
                (600, 10, 512, 64, 1024), (700, 15, 128, 32, 50)]


@pytest.mark.parametrize("n,avg_deg,d,k,max_nz", BANKED_CASES)
def test_banked_forward_and_backward(mk, n, avg_deg, d, k, max_nz):
    from oracle import c_oracle
    ptr, idx, val, x, dy = _problem(n, avg_deg, d, k, seed=n + d + k, kind="mean")
    wd, wi = c_oracle.maxk_cbsr(x, k)
    mk.set_max_nz(max_nz)
    try:
        tptr, tidx, tval = dev(ptr), dev(idx), dev(val)
        bd, bi, bs = mk.maxk_forward_banked(dev(x), k)
        out = mk.spgemm_forward_banked(tptr, tidx, tval, bd, bs, n, idx.size, k, d)
        want = c_oracle.spgemm_fwd(ptr, idx, val, wd, wi, d)
        bound = c_oracle.spgemm_fwd(ptr, idx, np.abs(val), np.abs(wd), wi, d)
        assert_rel(out, want, bound, "spgemm_forward_banked")
        assert torch.equal(out, mk.spgemm_forward_banked(tptr, tidx, tval, bd, bs, n, idx.size, k, d))
        dxs_b = mk.spgemm_backward_banked(tptr, tidx, tval, dev(dy), bs, n, idx.size, k, d)
        # banked entry order -> dense -> compare at the sorted positions
        dense = mk.cbsr_scatter(dxs_b, bi, d)
        got = mk.cbsr_gather(dense, _to_index_tensor(wi, d))
        want_b = c_oracle.sspmm_bwd(ptr, idx, val, dy, wi)
        bound_b = c_oracle.sspmm_bwd(ptr, idx, np.abs(val), np.abs(dy), wi)
        assert_rel(got, want_b, bound_b, "spgemm_backward_banked")
    finally:
        mk.set_max_nz(1024)


def test_graph_file_carries_the_work_records(mk, tmp_path):
    """f-4: the .warp4 successor -- records saved with the graph are installed, not rebuilt."""
    from spgemm_gnn_b200 import graph as G
    g = G.synthetic_graph(3000, 200000, seed=12, device="cuda")
    p0 = mk.partition(g.indptr, g.num_nodes(), 256)
    path = str(tmp_path / "g.npz")
    G.save_graph(g, path, max_nz=256)
    mk.clear_partition_cache()
    h = G.load_graph(path, device="cuda")
    before = mk.launch_count()
    p1 = mk.partition(h.indptr, h.num_nodes(), 256)
    assert mk.launch_count() == before                     # served from the installed records
    assert p1.num_slots == p0.num_slots and torch.equal(p1.parts[: p1.num_parts], p0.parts[: p0.num_parts])


@pytest.mark.parametrize("n_blocks", [2, 3, 7])
def test_column_blocked_backward_matches_oracle(mk, n_blocks):
    """The destination-blocked record order of the backward (large graphs) changes nothing but
    the order of the atomic sums."""
    from oracle import c_oracle
    n, d, k = 3000, 256, 32
    ptr, idx, val, x, dy = _problem(n, 40, d, k, seed=77, kind="mean")
    wd, wi = c_oracle.maxk_cbsr(x, k)
    tptr, tidx = dev(ptr), dev(idx)
    part = mk.partition_blocked(tptr, tidx, n, n, n_blocks, 64)
    recs = part.parts[: part.num_parts].cpu().numpy()
    # every stored entry exactly once; inside a record all columns belong to one block
    width = -(-n // n_blocks)
    cover = np.concatenate([np.arange(l, l + ln) for _, l, ln, _ in recs])
    assert np.array_equal(np.sort(cover), np.arange(idx.size))
    for r, l, ln, _ in recs[:: max(len(recs) // 200, 1)]:
        assert ptr[r] <= l and l + ln <= ptr[r + 1] and ln > 0
        blocks = idx[l:l + ln] // width
        assert blocks.min() == blocks.max()
    first_block = np.array([idx[l] // width for _, l, _, _ in recs])
    assert np.all(np.diff(first_block) >= 0)                       # blocks in order
    from spgemm_gnn_b200 import _lib
    dxs = torch.empty((n, k), dtype=torch.float32, device="cuda")
    rc = _lib.lib().mk_sspmm_bwd(part.parts.data_ptr(), part.num_parts, tidx.data_ptr(), dev(val).data_ptr(),
                                 dev(dy).data_ptr(), dev(wi).data_ptr(), 1, dxs.data_ptr(), n, n, k, d,
                                 torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    want = c_oracle.sspmm_bwd(ptr, idx, val, dy, wi)
    bound = c_oracle.sspmm_bwd(ptr, idx, np.abs(val), np.abs(dy), wi)
    assert_rel(dxs, want, bound, "blocked backward")
    # unsorted rows: falls back to the plain record list
    perm_idx = idx.copy()
    perm_idx[ptr[5]:ptr[6]] = perm_idx[ptr[5]:ptr[6]][::-1]
    p2 = mk.partition_blocked(tptr, dev(perm_idx), n, n, n_blocks, 64)
    assert p2.num_parts == mk.partition(tptr, n, 64).num_parts


@pytest.mark.parametrize("shape,d,k", [("ogbn-products", 256, 32), ("ogbn-proteins", 256, 64), ("yelp", 384, 16)])
def test_other_baseline_shapes_full_size(mk, shape, d, k):
    """BASELINE.json configs 4 and 5 (and the Yelp width, uint16 column ids) at full size:
    top-k against torch.topk, forward against dense cuSPARSE SpMM, backward through adjointness."""
    from spgemm_gnn_b200.graph import shaped_graph
    g = shaped_graph(shape, device="cuda")
    n, e = g.num_nodes(), g.num_edges()
    gen = torch.Generator(device="cuda").manual_seed(97)
    x = torch.randn(n, d, device="cuda", generator=gen)
    dy = torch.randn(n, d, device="cuda", generator=gen)
    val = g.edge_weights("both")
    sp_data, sp_index = mk.maxk_forward_cbsr(x, k)
    assert sp_index.dtype == (torch.uint8 if d <= 256 else torch.uint16)
    cols = (sp_index if d <= 256 else sp_index.view(torch.int16)).to(torch.int64) & 0xFFFF
    tv, ti = torch.topk(x, k, dim=1)
    assert torch.equal(cols, ti.sort(dim=1)[0])                      # tie-free: same set, ascending
    assert torch.equal(torch.gather(x, 1, cols), sp_data)
    out, _ = mk.spgemm_forward(g.indptr, g.indices, val, sp_data, sp_index, n, e, k, d)
    out2, _ = mk.spgemm_forward(g.indptr, g.indices, val, sp_data, sp_index, n, e, k, d)
    assert torch.equal(out, out2)
    xm = mk.cbsr_scatter(sp_data, sp_index, d)
    adj = torch.sparse_csr_tensor(g.indptr.long(), g.indices.long(), val, size=(n, n))
    ref = torch.sparse.mm(adj, xm)
    assert float((out - ref).abs().max()) <= 2e-5 * float(ref.abs().max())
    del ref, xm, adj
    dxs = mk.spgemm_backward(g.indptr, g.indices, val, dy, sp_index, n, e, k, d)
    lhs = float((out.double() * dy.double()).sum())
    rhs = float((sp_data.double() * dxs.double()).sum())
    scale = float((out.double().abs() * dy.double().abs()).sum())
    assert abs(lhs - rhs) <= 1e-8 * scale
    oracle_sample_check(g, val, sp_data, sp_index, out, dy, dxs, d, k)


def test_empty_and_extreme_inputs(mk):
    """n = 0, E = 0, k = 1, k = D, one row holding every stored entry."""
    from oracle import c_oracle
    dev0 = "cuda"
    # n = 0
    d0, i0 = mk.maxk_forward_cbsr(torch.empty((0, 64), device=dev0), 8)
    assert d0.shape == (0, 8) and i0.shape == (0, 8)
    assert mk.cbsr_scatter(d0, i0, 64).shape == (0, 64)
    # graph without stored entries: output rows are zero, gradient is zero
    n, d, k = 50, 64, 8
    x = torch.randn(n, d, device=dev0)
    sd, si = mk.maxk_forward_cbsr(x, k)
    ptr = torch.zeros(n + 1, dtype=torch.int32, device=dev0)
    idx = torch.zeros(0, dtype=torch.int32, device=dev0)
    val = torch.zeros(0, device=dev0)
    out, _ = mk.spgemm_forward(ptr, idx, val, sd, si, n, 0, k, d)
    assert out.shape == (n, d) and float(out.abs().max()) == 0.0
    dxs = mk.spgemm_backward(ptr, idx, val, torch.randn(n, d, device=dev0), si, n, 0, k, d)
    assert dxs.shape == (n, k) and float(dxs.abs().max()) == 0.0
    # k = 1 and k = D (CBSR == dense row, identity permutation)
    rng = np.random.default_rng(3)
    xh = rng.standard_normal((40, 32)).astype(np.float32)
    for kk in (1, 32):
        gd, gi = mk.maxk_forward_cbsr(dev(xh), kk)
        wd, wi = c_oracle.maxk_cbsr(xh, kk)
        assert np.array_equal(gi.cpu().numpy(), wi) and np.array_equal(gd.cpu().numpy(), wd)
    # one row holds all 70,000 stored entries (more than 65,535, many records, partial folding)
    n, d, k = 5000, 256, 32
    xh = rng.standard_normal((n, d)).astype(np.float32)
    wd, wi = c_oracle.maxk_cbsr(xh, k)
    e = 70000
    ptr_h = np.zeros(n + 1, np.int32)
    ptr_h[8:] = e                                    # row 7 owns everything
    idx_h = rng.integers(0, n, e).astype(np.int32)
    idx_h.sort()
    val_h = rng.standard_normal(e).astype(np.float32)
    dy_h = rng.standard_normal((n, d)).astype(np.float32)
    out, _ = mk.spgemm_forward(dev(ptr_h), dev(idx_h), dev(val_h), dev(wd), dev(wi), n, e, k, d)
    want = c_oracle.spgemm_fwd(ptr_h, idx_h, val_h, wd, wi, d)
    bound = c_oracle.spgemm_fwd(ptr_h, idx_h, np.abs(val_h), np.abs(wd), wi, d)
    assert_rel(out, want, bound, "single huge row fwd")
    dxs = mk.spgemm_backward(dev(ptr_h), dev(idx_h), dev(val_h), dev(dy_h), dev(wi), n, e, k, d)
    assert_rel(dxs, c_oracle.sspmm_bwd(ptr_h, idx_h, val_h, dy_h, wi),
               c_oracle.sspmm_bwd(ptr_h, idx_h, np.abs(val_h), np.abs(dy_h), wi), "single huge row bwd")


@pytest.mark.parametrize("n,d", [(1000, 256), (333, 64), (257, 384), (100, 1024), (64, 100)])
@pytest.mark.parametrize("with_b,with_bias", [(True, True), (True, False), (False, True), (False, False)])
def test_fused_add_layernorm_matches_torch(mk, n, d, with_b, with_bias):
    """f-3 epilogue against a plain PyTorch float64 reference of the same op (fp32 kernel:
    tolerance 2e-5 relative to the largest element)."""
    import torch.nn.functional as F
    from spgemm_gnn_b200.maxk_layers import add_layer_norm
    gen = torch.Generator().manual_seed(n + d)
    a = torch.randn(n, d, generator=gen) * 3 + 1
    b = torch.randn(n, d, generator=gen) if with_b else None
    bias = torch.randn(d, generator=gen) if with_bias else None
    gy = torch.randn(n, d, generator=gen)
    norm = torch.nn.LayerNorm(d)
    with torch.no_grad():
        norm.weight.copy_(torch.randn(d, generator=gen))
        norm.bias.copy_(torch.randn(d, generator=gen))
    # float64 reference
    ra = a.double().requires_grad_(True)
    rb = b.double().requires_grad_(True) if with_b else None
    rbias = bias.double().requires_grad_(True) if with_bias else None
    rw, rbeta = norm.weight.detach().double().requires_grad_(True), norm.bias.detach().double().requires_grad_(True)
    z = ra + (rb if with_b else 0) + (rbias if with_bias else 0)
    ry = F.layer_norm(z, (d,), rw, rbeta, norm.eps)
    ry.backward(gy.double())
    # CUDA path
    norm = norm.cuda()
    ca = a.cuda().requires_grad_(True)
    cb = b.cuda().requires_grad_(True) if with_b else None
    cbias = bias.cuda().requires_grad_(True) if with_bias else None
    y = add_layer_norm(ca, cb, cbias, norm)
    y.backward(gy.cuda())

    def close(got, want, what):
        err = float((got.detach().cpu().double() - want.detach()).abs().max())
        assert err <= 2e-5 * float(want.detach().abs().max()) + 1e-12, (what, err)

    close(y, ry, "y")
    close(ca.grad, ra.grad, "grad a")
    if with_b:
        close(cb.grad, rb.grad, "grad b")
    if with_bias:
        close(cbias.grad, rbias.grad, "grad bias")
    close(norm.weight.grad, rw.grad, "grad gamma")
    close(norm.bias.grad, rbeta.grad, "grad beta")


# ---------------------------------------------------------------------------------------
# the knobs of the row-partitioned / load-balanced forward (mk_spgemm_fwd_banked_ex)
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("k", [32, 64, 16])
def test_forward_arrival_order_walk_and_sorted_records(mk, k):
    """`split` (walk [split, end) before [begin, split)) and the longest-first record order change the
    order of execution, not the result: against the float64 oracle at the 1e-5 bar, bit-equal between
    record orders (each row is still summed by one warp in one fixed order), and with a pre-opened
    wait window the waiting kernel is bit-equal to the plain one."""
    from oracle import c_oracle
    from spgemm_gnn_b200 import maxk_kernels as mkk
    from conftest import small_graph
    g = small_graph(3000, 70, seed=8, device="cuda")
    n, e, d = g.num_nodes(), g.num_edges(), 256
    rng = np.random.default_rng(k)
    x = rng.standard_normal((n, d)).astype(np.float32)
    val = g.edge_weights("both")
    sd, si = mk.maxk_forward_cbsr(dev(x), k)
    bd, _, bs = mk.cbsr_bank(sd, si, d, with_index=False)
    mk.set_max_nz(64)
    try:
        mk.clear_partition_cache()
        world, rank = 4, 1
        r = -(-n // world)
        split = mk.block_split(g.indptr, g.indices, n, world, rank, r)
        # split really is the first entry of the row with column >= rank * r
        ptr, idx = g.indptr.cpu().numpy(), g.indices.cpu().numpy()
        want_split = np.array([ptr[i] + np.searchsorted(idx[ptr[i]:ptr[i + 1]], rank * r) for i in range(n)])
        assert np.array_equal(split.cpu().numpy(), want_split)
        base = mk.spgemm_forward_banked(g.indptr, g.indices, val, bd, bs, n, e, k, d)
        walked = mk.spgemm_forward_banked(g.indptr, g.indices, val, bd, bs, n, e, k, d, split=split)
        want = c_oracle.spgemm_fwd(ptr, idx, val.cpu().numpy(), sd.cpu().numpy(), si.cpu().numpy(), d)
        bound = c_oracle.spgemm_fwd(ptr, idx, np.abs(val.cpu().numpy()), np.abs(sd.cpu().numpy()), si.cpu().numpy(), d)
        assert_rel(base, want, bound, "banked forward, sorted records")
        assert_rel(walked, want, bound, "banked forward, arrival-order walk")
        was = mkk._EXEC_SORTED
        mkk._EXEC_SORTED = False
        mk.clear_partition_cache()
        try:
            row_order = mk.spgemm_forward_banked(g.indptr, g.indices, val, bd, bs, n, e, k, d)
        finally:
            mkk._EXEC_SORTED = was
        assert torch.equal(base, row_order)
    finally:
        mk.set_max_nz(1024)
        mk.clear_partition_cache()


@pytest.mark.parametrize("world,rank,k,max_nz", [(8, 5, 32, 64), (4, 0, 32, 1024), (2, 1, 64, 64), (3, 2, 32, 16)])
def test_forward_in_source_block_phases(mk, world, rank, k, max_nz):
    """The forward cut into source-block phases (mk_spgemm_fwd_banked_phase: own block, the next
    senders, the rest; one launch each, later ones adding to the rows of the earlier ones) covers every
    stored entry exactly once: float64 oracle at the 1e-5 bar, rows of several records included, and
    repeatable bit for bit."""
    from oracle import c_oracle
    from conftest import small_graph
    g = small_graph(2600, 90, seed=21, device="cuda")
    n, e, d = g.num_nodes(), g.num_edges(), 256
    rng = np.random.default_rng(world * 10 + rank)
    x = rng.standard_normal((n, d)).astype(np.float32)
    val = g.edge_weights("mean")
    sd, si = mk.maxk_forward_cbsr(dev(x), k)
    bd, _, bs = mk.cbsr_bank(sd, si, d, with_index=False)
    mk.set_max_nz(max_nz)
    try:
        mk.clear_partition_cache()
        r = -(-n // world)
        blk = mk.block_pointers(g.indptr, g.indices, n, world, r)
        phases = mk.forward_phases(world, rank)
        got = mk.spgemm_forward_banked(g.indptr, g.indices, val, bd, bs, n, e, k, d, phases=phases, blk=blk,
                                       n_blocks=world)
        again = mk.spgemm_forward_banked(g.indptr, g.indices, val, bd, bs, n, e, k, d, phases=phases, blk=blk,
                                         n_blocks=world)
        ptr, idx = g.indptr.cpu().numpy(), g.indices.cpu().numpy()
        want = c_oracle.spgemm_fwd(ptr, idx, val.cpu().numpy(), sd.cpu().numpy(), si.cpu().numpy(), d)
        bound = c_oracle.spgemm_fwd(ptr, idx, np.abs(val.cpu().numpy()), np.abs(sd.cpu().numpy()), si.cpu().numpy(), d)
        assert_rel(got, want, bound, f"phased forward, {len(phases)} phases")
        assert torch.equal(got, again)
        # a single phase over all blocks is the one-launch forward in the `split` order, bit for bit
        whole = [(rank, world, 0, rank)]
        split = mk.block_split(g.indptr, g.indices, n, world, rank, r)
        one = mk.spgemm_forward_banked(g.indptr, g.indices, val, bd, bs, n, e, k, d, phases=whole, blk=blk,
                                       n_blocks=world)
        ref = mk.spgemm_forward_banked(g.indptr, g.indices, val, bd, bs, n, e, k, d, split=split)
        assert torch.equal(one, ref)
    finally:
        mk.set_max_nz(1024)
        mk.clear_partition_cache()


@pytest.mark.parametrize("n,deg,d,k", [(1500, 150, 256, 8), (1500, 150, 256, 16), (800, 200, 128, 8),
                                       (600, 180, 384, 16), (700, 120, 64, 16)])
def test_packed_banked_forward_k8_k16(mk, n, deg, d, k):
    """k = 8, 16 on long records go through the packed banked table (8-byte entries): the packed
    table is the banked table re-encoded, and the forward meets the oracle at the 1e-5 bar; the
    reference-named spgemm_forward takes that route by itself."""
    from oracle import c_oracle
    from conftest import small_graph
    g = small_graph(n, deg, seed=k + d, device="cuda")
    e = g.num_edges()
    rng = np.random.default_rng(n + k)
    x = rng.standard_normal((n, d)).astype(np.float32)
    x[7] = 0.0                                          # zero row: every entry skipped
    val = g.edge_weights("mean")
    sd, si = mk.maxk_forward_cbsr(dev(x), k)
    assert mk.packed_supported(k, d)
    pack = mk.cbsr_bank_packed(sd, si, d)
    bd, bi, bs = mk.cbsr_bank(sd, si, d)
    pk = pack.cpu().numpy().view(np.uint32)
    assert np.array_equal(pk[:, :, 0], bd.cpu().numpy().view(np.uint32))
    assert np.array_equal(pk[:, :, 1] & 0xffff, bs.cpu().numpy().view(np.uint16).astype(np.uint32))
    assert np.array_equal(pk[:, :, 1] >> 16, _np_index(bi, d).astype(np.uint32))
    out = mk.spgemm_forward_packed(g.indptr, g.indices, val, pack, n, e, k, d)
    ptr, idx = g.indptr.cpu().numpy(), g.indices.cpu().numpy()
    wd, wi = sd.cpu().numpy(), _np_index(si, d)
    want = c_oracle.spgemm_fwd(ptr, idx, val.cpu().numpy(), wd, wi, d)
    bound = c_oracle.spgemm_fwd(ptr, idx, np.abs(val.cpu().numpy()), np.abs(wd), wi, d)
    assert_rel(out, want, bound, "packed banked forward")
    part = mk.partition(g.indptr, n)
    assert mk.use_packed(part.num_parts, e, k, d)
    auto, _ = mk.spgemm_forward(g.indptr, g.indices, val, sd, si, n, e, k, d)
    assert torch.equal(auto, out)
    plain, _ = mk.spgemm_forward(g.indptr, g.indices, val, sd, si, n, e, k, d, allow_banked=False)
    assert_rel(plain, want, bound, "plain forward")


@pytest.mark.parametrize("n,deg,d,k,tile_mb", [(3000, 40, 256, 32, 0), (2500, 60, 256, 16, 0), (2000, 30, 128, 8, 0),
                                               (1500, 50, 384, 64, 0), (4000, 12, 256, 32, 0)])
def test_tiled_column_blocked_backward(mk, n, deg, d, k, tile_mb):
    """mk_sspmm_bwd_tiled (the backward for CBSR gradients larger than L2): forced on with column
    blocks of a few hundred sources, it meets the oracle at the 1e-5 bar and agrees with the plain
    kernel -- including empty rows, a row that holds a large share of all stored entries, and a row
    count that is not a multiple of the 8-row tile."""
    from oracle import c_oracle
    from spgemm_gnn_b200 import maxk_kernels as mkk
    from conftest import small_graph
    g = small_graph(n - 3, deg, seed=n + k, device="cuda")          # n-3 rows: ragged last tile
    nn, e = g.num_nodes(), g.num_edges()
    rng = np.random.default_rng(n + d)
    x = rng.standard_normal((nn, d)).astype(np.float32)
    dy = rng.standard_normal((nn, d)).astype(np.float32)
    val = g.edge_weights("both")
    sd, si = mk.maxk_forward_cbsr(dev(x), k)
    ptr, idx, v = g.indptr.cpu().numpy(), g.indices.cpu().numpy(), val.cpu().numpy()
    wi = _np_index(si, d)
    want = c_oracle.sspmm_bwd(ptr, idx, v, dy, wi)
    bound = c_oracle.sspmm_bwd(ptr, idx, np.abs(v), np.abs(dy), wi)
    plain = mk.spgemm_backward(g.indptr, g.indices, val, dev(dy), si, nn, e, k, d)
    saved = (mkk._BWD_TILED, mkk._BWD_TILE_MB)
    try:
        mkk._BWD_TILED = "1"
        for blocks in (2, 5, 13):
            mkk._blkptr_cache.clear()
            mkk.backward_tiles = lambda *_a, _b=blocks: _b            # this many column blocks
            tiled = mkk.spgemm_backward(g.indptr, g.indices, val, dev(dy), si, nn, e, k, d)
            assert_rel(tiled, want, bound, f"tiled backward, {blocks} column blocks")
            assert float((tiled - plain).abs().max()) <= 2e-5 * float(plain.abs().max())
    finally:
        mkk._BWD_TILED, mkk._BWD_TILE_MB = saved
        mkk.backward_tiles = _ORIG_BACKWARD_TILES
        mkk._blkptr_cache.clear()
    # unsorted column ids: the blocked form refuses (falls back to the plain kernel), same result
    perm_idx = g.indices.clone()
    lo, hi = int(g.indptr[5]), int(g.indptr[6])
    perm_idx[lo:hi] = perm_idx[lo:hi].flip(0)
    if hi - lo > 1:
        assert mkk.block_pointers(g.indptr, perm_idx, nn, 4, -(-nn // 4)) is None


from spgemm_gnn_b200 import maxk_kernels as _mkk_for_tiles  # noqa: E402
_ORIG_BACKWARD_TILES = _mkk_for_tiles.backward_tiles
