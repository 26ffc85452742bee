import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "maxk_reference.npz"))


@pytest.fixture(scope="session")
def golden_layers():
    """Outputs and kernel-call arguments of the reference's own MaxKSAGEConv / MaxKGCNConv code
    (tests/golden/make_golden_layers.py)."""
    return np.load(os.path.join(ROOT, "tests", "golden", "layers_reference.npz"))


def layer_cases(gl):
    """(name, call-argument dict, expected layer output, the part added to the aggregation)."""
    for name in [str(s) for s in gl["names"]]:
        call = {k: gl[f"{name}_call_{k}"] for k in ("ptr", "idx", "val", "sp_data", "sp_index")}
        n, e, k, d = (int(v) for v in gl[f"{name}_call_dims"])
        call.update(n=n, e=e, k=k, d=d)
        add = gl[f"{name}_h_self"] if name.startswith("sage") else gl[f"{name}_bias"][None, :]
        yield name, call, gl[f"{name}_y"], add


@pytest.fixture(scope="session")
def built_lib():
    """The product library, built in-tree if stale (nvcc cross-compiles without a GPU)."""
    from spgemm_gnn_b200 import build as _b
    return _b.build()


def small_graph(n, avg_deg, seed=97, device="cpu"):
    from spgemm_gnn_b200.graph import synthetic_graph
    return synthetic_graph(n, int(n * avg_deg), seed=seed, device=device)


def _assert_rel(got, want64, bound, what, tol=1e-5):
    got = got.detach().cpu().numpy().astype(np.float64)
    err = np.abs(got - want64)
    worst = float((err / (tol * bound + 1e-30)).max()) if err.size else 0.0
    assert worst <= 1.0, f"{what}: error is {worst:.3g} x the {tol:g}*sum|terms| bound"


def oracle_sample_check(g, val, sp_data, sp_index, out, dy, dxs, d, k, n_sample=2000, seed=5):
    """Full-size parity at the north_star bar (1e-5 * sum|terms| per element, float64 C oracle):
    forward rows = `n_sample` random rows + the 10 highest-degree rows; backward = the CBSR-gradient
    rows of `n_sample` random source nodes + the 10 with the most in-edges.  The sub-problems are cut
    out on the GPU (the rows' stored entries; every stored entry that points at a sampled source)
    and handed to oracle/maxk_oracle.c."""
    import torch
    from oracle import c_oracle
    dev = g.indptr.device
    n = g.num_nodes()
    gen = torch.Generator().manual_seed(seed)
    deg = g.in_degrees()
    rows = torch.unique(torch.cat([torch.randperm(n, generator=gen)[:n_sample],
                                   torch.topk(deg, min(10, n))[1].cpu()])).to(dev)
    ptr64 = g.indptr.to(torch.int64)
    lens = ptr64[rows + 1] - ptr64[rows]
    sub_ptr = torch.zeros(rows.numel() + 1, dtype=torch.int64, device=dev)
    sub_ptr[1:] = torch.cumsum(lens, 0)
    eid = (torch.repeat_interleave(ptr64[rows] - sub_ptr[:-1], lens)
           + torch.arange(int(sub_ptr[-1]), device=dev))
    wd = sp_data.cpu().numpy()
    wi = sp_index.cpu().numpy() if d <= 256 else sp_index.view(torch.int16).cpu().numpy().view(np.uint16)
    s_ptr, s_idx = sub_ptr.to(torch.int32).cpu().numpy(), g.indices[eid].cpu().numpy()
    s_val = val[eid].cpu().numpy()
    want = c_oracle.spgemm_fwd(s_ptr, s_idx, s_val, wd, wi, d)
    bound = c_oracle.spgemm_fwd(s_ptr, s_idx, np.abs(s_val), np.abs(wd), wi, d)
    _assert_rel(out[rows], want, bound, f"forward, {rows.numel()} sampled rows incl. max degree {int(deg.max())}")

    # backward: every stored entry (r <- j) with j in the sample, relabelled to a compact problem
    out_deg = g.out_degrees()
    srcs = torch.unique(torch.cat([torch.randperm(g.num_src, generator=gen)[:n_sample],
                                   torch.topk(out_deg, min(10, g.num_src))[1].cpu()])).to(dev)
    mark = torch.zeros(g.num_src, dtype=torch.bool, device=dev)
    mark[srcs] = True
    eid = mark[g.indices.to(torch.int64)].nonzero().squeeze(1)
    r_of_e = g.row_ids()[eid]
    rset, rinv = torch.unique(r_of_e, return_inverse=True)       # ascending, like eid
    cinv = torch.searchsorted(srcs, g.indices[eid].to(torch.int64))
    b_ptr = torch.zeros(rset.numel() + 1, dtype=torch.int64, device=dev)
    b_ptr[1:] = torch.cumsum(torch.bincount(rinv, minlength=rset.numel()), 0)
    b_val = val[eid].cpu().numpy()
    b_dy = dy[rset].cpu().numpy()
    b_wi = np.ascontiguousarray(wi[srcs.cpu().numpy()])
    b_ptr_np, b_idx_np = b_ptr.to(torch.int32).cpu().numpy(), cinv.to(torch.int32).cpu().numpy()
    want_b = c_oracle.sspmm_bwd(b_ptr_np, b_idx_np, b_val, b_dy, b_wi)
    bound_b = c_oracle.sspmm_bwd(b_ptr_np, b_idx_np, np.abs(b_val), np.abs(b_dy), b_wi)
    _assert_rel(dxs[srcs], want_b, bound_b, f"backward, {srcs.numel()} sampled sources ({eid.numel()} stored entries)")
    g._cache.pop("row_ids", None)


