#!/usr/bin/env python
"""Forward SpGEMM + add + LayerNorm as separate kernels against the epilogue inside the SpGEMM
(mk_spgemm_fwd_banked_ln), training form (z, mean, rstd kept) and inference form."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import maxk_kernels as mk
from spgemm_gnn_b200.graph import shaped_graph

wl = sys.argv[1] if len(sys.argv) > 1 else "reddit"
k, d = 32, 256
g = shaped_graph(wl, device="cuda")
n, e = g.num_nodes(), g.num_edges()
val = g.edge_weights("mean")
gen = torch.Generator(device="cuda").manual_seed(97)
x = torch.randn(n, d, device="cuda", generator=gen)
hs = torch.randn(n, d, device="cuda", generator=gen)
gamma, beta = torch.randn(d, device="cuda", generator=gen), torch.randn(d, device="cuda", generator=gen)
sd, si = mk.maxk_forward_cbsr(x, k)
bd, _, bs = mk.cbsr_bank(sd, si, d, with_index=False)


def t(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def separate():
    agg = mk.spgemm_forward_banked(g.indptr, g.indices, val, bd, bs, n, e, k, d)
    return mk.add_layernorm_forward(hs, agg, None, gamma, beta, 1e-5)


f = t(lambda: mk.spgemm_forward_banked(g.indptr, g.indices, val, bd, bs, n, e, k, d))
s = t(separate)
ft = t(lambda: mk.spgemm_forward_ln(g.indptr, g.indices, val, bd, bs, n, e, k, d, hs, None, gamma, beta, 1e-5))
fi = t(lambda: mk.spgemm_forward_ln(g.indptr, g.indices, val, bd, bs, n, e, k, d, hs, None, gamma, beta, 1e-5,
                                    keep_stats=False))
print(f"{wl} k={k} d={d}: forward alone {f:.3f} ms | forward + add_layernorm kernels {s:.3f} ms | "
      f"epilogue inside the forward: training form {ft:.3f} ms, inference form {fi:.3f} ms")
