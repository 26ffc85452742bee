"""The reference's real FFI: `maxk_kernels` as a compiled pybind11 / ATen extension
(spgemm_gnn_b200/binding/maxk_bindings.cpp, the module setup.py:25-31 of the reference builds) over
the same C ABI the ctypes shim uses.  Same parity bars as tests/test_gpu_parity.py, same inputs;
the two bindings must agree bit for bit (they launch the same kernels)."""
import time

import numpy as np
import pytest
import torch

from conftest import _assert_rel, small_graph


@pytest.fixture(scope="module")
def ext(built_lib):
    from spgemm_gnn_b200 import build as _b
    _b.build_binding()
    from spgemm_gnn_b200 import maxk_kernels_ext
    return maxk_kernels_ext


def test_binding_imports_and_keeps_the_reference_messages(ext):
    """No GPU needed: the module loads, reports the library's ABI version and rejects CPU tensors with
    the TORCH_CHECK strings of the reference binary."""
    from spgemm_gnn_b200 import _lib
    assert ext.abi_version() == _lib.lib().mk_version()
    x = torch.randn(4, 16)
    with pytest.raises(RuntimeError, match="input must be a CUDA tensor"):
        ext.maxk_forward(x, 4)
    with pytest.raises(RuntimeError, match="grad_output must be a CUDA tensor"):
        ext.maxk_backward(x, torch.zeros(4, 16, dtype=torch.uint8))
    i32 = torch.zeros(5, dtype=torch.int32)
    with pytest.raises(RuntimeError, match="ptr must be a CUDA tensor"):
        ext.spgemm_forward(i32, i32, x, x, x, 4, 4, 16, 16)
    with pytest.raises(RuntimeError, match="ptr must be a CUDA tensor"):
        ext.spgemm_backward(i32, i32, x, x, x, 4, 4, 16, 16)


@pytest.mark.gpu
@pytest.mark.parametrize("n,deg,d,k", [(2000, 150, 256, 32), (1500, 150, 256, 16), (2000, 30, 256, 32),
                                       (900, 140, 384, 16), (700, 20, 100, 7), (1200, 160, 256, 64)])
def test_binding_parity_and_equality_with_the_ctypes_shim(ext, n, deg, d, k):
    from oracle import c_oracle
    import maxk_kernels as mk
    g = small_graph(n, deg, seed=n + k, device="cuda")
    e = g.num_edges()
    rng = np.random.default_rng(k * 7 + d)
    x = rng.standard_normal((n, d)).astype(np.float32)
    dy = rng.standard_normal((n, d)).astype(np.float32)
    xc, dyc = torch.from_numpy(x).cuda(), torch.from_numpy(dy).cuda()
    val = g.edge_weights("both")
    sd, si = ext.maxk_forward_cbsr(xc, k)
    wd, wi = c_oracle.maxk_cbsr(x, k)                                   # a-1: bit-exact
    gi = si.cpu().numpy() if d <= 256 else si.view(torch.int16).cpu().numpy().view(np.uint16)
    assert np.array_equal(gi, wi) and np.array_equal(sd.cpu().numpy().view(np.uint32), wd.view(np.uint32))
    assert torch.equal(ext.maxk_forward(xc, k), sd)
    out, si2 = ext.spgemm_forward(g.indptr, g.indices, val, sd, si, n, e, k, d)   # a-3
    assert si2.data_ptr() == si.data_ptr()
    ptr, idx, v = g.indptr.cpu().numpy(), g.indices.cpu().numpy(), val.cpu().numpy()
    want = c_oracle.spgemm_fwd(ptr, idx, v, wd, wi, d)
    bound = c_oracle.spgemm_fwd(ptr, idx, np.abs(v), np.abs(wd), wi, d)
    _assert_rel(out, want, bound, "pybind spgemm_forward")
    dxs = ext.spgemm_backward(g.indptr, g.indices, val, dyc, si, n, e, k, d)      # a-4
    want_b = c_oracle.sspmm_bwd(ptr, idx, v, dy, wi)
    bound_b = c_oracle.sspmm_bwd(ptr, idx, np.abs(v), np.abs(dy), wi)
    _assert_rel(dxs, want_b, bound_b, "pybind spgemm_backward")
    dense = ext.maxk_backward(dxs, si.to(torch.int64))                            # a-2, int64 ids as the reference saves
    d_inf = dense.shape[1]
    assert d_inf == max(int(wi.max()) + 1, k)
    assert torch.equal(dense, mk.cbsr_scatter(dxs, si, d)[:, :d_inf])
    out_c, _ = mk.spgemm_forward(g.indptr, g.indices, val, sd, si, n, e, k, d)   # same kernels, same order
    assert torch.equal(out, out_c)


@pytest.mark.gpu
def test_binding_host_overhead(ext, capsys):
    """Per-call host cost of the two bindings on a tiny input (launch-bound regime): printed for the
    record, asserted only to be sane."""
    import maxk_kernels as mk
    x = torch.randn(64, 64, device="cuda")
    res = {}
    for name, fn in (("ctypes", mk.maxk_forward), ("pybind", ext.maxk_forward)):
        for _ in range(50):
            fn(x, 8)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(2000):
            fn(x, 8)
        host = (time.perf_counter() - t0) / 2000
        torch.cuda.synchronize()
        res[name] = host * 1e6
    with capsys.disabled():
        print(f"\n[binding overhead] maxk_forward host time per call: ctypes {res['ctypes']:.1f} us, pybind {res['pybind']:.1f} us")
    assert res["pybind"] < 200 and res["ctypes"] < 400
