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
ks = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [32]   # 8, 16: the packed table
tag = os.path.basename(os.environ.get("MAXK_LIB", "product"))
d = 256
g = shaped_graph(wl, device="cuda")
n, e = g.num_nodes(), g.num_edges()
val = g.edge_weights("mean")
gen = torch.Generator(device="cuda").manual_seed(97)
x = torch.randn(n, d, device="cuda", generator=gen)
hs = torch.randn(n, d, device="cuda", generator=gen)
gamma, beta = torch.randn(d, device="cuda", generator=gen), torch.randn(d, device="cuda", generator=gen)
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



for k in ks:
    sd, si = mk.maxk_forward_cbsr(x, k)
    if k in (8, 16):
        bd, bs = mk.cbsr_bank_packed(sd, si, d), None
        fwd_plain = lambda: mk.spgemm_forward_packed(g.indptr, g.indices, val, bd, n, e, k, d)
    else:
        bd, _, bs = mk.cbsr_bank(sd, si, d, with_index=False)
        fwd_plain = lambda: mk.spgemm_forward_banked(g.indptr, g.indices, val, bd, bs, n, e, k, d)


    def separate():
        agg = fwd_plain()
        return mk.add_layernorm_forward(hs, agg, None, gamma, beta, 1e-5)


    f = t(fwd_plain)
    s = t(separate)
    ft = t(lambda: mk.spgemm_forward_ln(g.indptr, g.indices, val, bd, bs, n, e, k, d, hs, None, gamma, beta, 1e-5))
    fi = t(lambda: mk.spgemm_forward_ln(g.indptr, g.indices, val, bd, bs, n, e, k, d, hs, None, gamma, beta, 1e-5,
                                        keep_stats=False))
    print(f"[{tag}] {wl} k={k} d={d}: forward alone {f:.3f} ms | forward + add_layernorm kernels {s:.3f} ms | "
          f"epilogue inside the forward: training form {ft:.3f} ms, inference form {fi:.3f} ms")
