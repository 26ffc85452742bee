"""Property-based parity cases (hypothesis): random shapes and adversarial value patterns for the
MaxK selection, on the CPU between the two oracles and on the GPU against them."""
import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st, HealthCheck

from oracle import c_oracle, maxk_oracle as mo

SPECIAL = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 1e-38, -1e-38, 3.0, 3.0, 3.0], dtype=np.float32)


def _matrix(seed, n, d, mode):
    rng = np.random.default_rng(seed)
    if mode == "normal":
        return rng.standard_normal((n, d)).astype(np.float32)
    if mode == "quantised":          # many exact ties
        return (np.round(rng.standard_normal((n, d)) * 2) / 2).astype(np.float32)
    if mode == "special":            # NaN / inf / signed zeros / denormals / repeated values
        return SPECIAL[rng.integers(0, SPECIAL.size, (n, d))]
    x = rng.standard_normal((n, d)).astype(np.float32)   # "constant rows"
    x[::2] = x[::2, :1]
    return x


shapes = st.tuples(st.integers(1, 40), st.sampled_from([8, 32, 33, 64, 100, 256, 260, 384, 1030]))
modes = st.sampled_from(["normal", "quantised", "special", "constant"])


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(shape=shapes, mode=modes, seed=st.integers(0, 2**20), kfrac=st.floats(0.0, 1.0))
def test_numpy_and_c_oracle_agree_on_any_input(shape, mode, seed, kfrac):
    n, d = shape
    k = max(1, min(d, int(round(kfrac * d))))
    x = _matrix(seed, n, d, mode)
    a_data, a_idx = mo.maxk_cbsr(x, k)
    b_data, b_idx = c_oracle.maxk_cbsr(x, k)
    assert np.array_equal(a_idx, b_idx)
    assert np.array_equal(a_data.view(np.uint32), b_data.view(np.uint32))
    assert np.all(np.diff(a_idx.astype(np.int64), axis=1) > 0)            # ascending, distinct
    # the kept set dominates the dropped set in the contract's order
    key = mo.order_key(x).astype(np.int64)
    kept = np.zeros((n, d), bool)
    np.put_along_axis(kept, a_idx.astype(np.int64), True, axis=1)
    lo_kept = np.where(kept, key, np.iinfo(np.int64).max).min(1)
    hi_drop = np.where(~kept, key, -1).max(1)
    assert np.all(lo_kept >= hi_drop)


@pytest.mark.gpu
@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(shape=shapes, mode=modes, seed=st.integers(0, 2**20), kfrac=st.floats(0.0, 1.0))
def test_cuda_topk_bit_exact_on_any_input(shape, mode, seed, kfrac):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import maxk_kernels as mk
    n, d = shape
    k = max(1, min(d, int(round(kfrac * d))))
    x = _matrix(seed, n, d, mode)
    want_data, want_idx = c_oracle.maxk_cbsr(x, k)
    data, idx = mk.maxk_forward_cbsr(torch.from_numpy(x).cuda(), k)
    got_idx = idx.cpu().numpy() if d <= 256 else idx.view(torch.int16).cpu().numpy().view(np.uint16)
    assert np.array_equal(got_idx, want_idx)
    assert np.array_equal(data.cpu().numpy().view(np.uint32), want_data.view(np.uint32))
    # scatter / gather round trip on the same positions
    dense = mk.cbsr_scatter(data, idx, d)
    assert torch.equal(mk.cbsr_gather(dense, idx).view(torch.int32), data.view(torch.int32))


@pytest.mark.gpu
@settings(max_examples=15, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(seed=st.integers(0, 2**20), k=st.sampled_from([8, 16, 32, 64]), d=st.sampled_from([64, 128, 256, 384, 512]),
       mode=st.sampled_from(["normal", "quantised"]))
def test_banked_form_invariants_on_any_input(seed, k, d, mode):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import maxk_kernels as mk
    if k > d:
        return
    x = _matrix(seed, 64, d, mode)
    wd, wi = c_oracle.maxk_cbsr(x, k)
    ti = torch.from_numpy(wi).cuda() if d <= 256 else torch.from_numpy(wi.view(np.int16)).cuda().view(torch.uint16)
    bd, bi, bs = mk.cbsr_bank(torch.from_numpy(wd).cuda(), ti, d)
    gi = bi.cpu().numpy() if d <= 256 else bi.view(torch.int16).cpu().numpy().view(np.uint16)
    mean_wf, _ = mo.check_banked(wd, wi, bd.cpu().numpy(), gi, bs.cpu().numpy().view(np.uint16), d)
    assert 1.0 <= mean_wf <= 3.0
