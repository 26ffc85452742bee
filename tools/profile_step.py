#!/usr/bin/env python
"""Small driver for ncu / timing: W warm-up steps then S steps of (forward SpGEMM, backward SSpMM)
on one of the BASELINE shapes, plain or banked.  Times each with CUDA events."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import maxk_kernels as mk
from spgemm_gnn_b200.graph import shaped_graph

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="reddit")
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--k", type=int, default=32)
ap.add_argument("--dim", type=int, default=256)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--max-nz", type=int, default=None)
ap.add_argument("--plain", action="store_true", help="unbanked kernels")
ap.add_argument("--bwd-block-mb", type=int, default=None, help="column-block size of the backward (0 = off)")
ap.add_argument("--banked-bwd", action="store_true", help="banked backward too (default: plain backward)")
a = ap.parse_args()
if a.max_nz:
    mk.set_max_nz(a.max_nz)
if a.bwd_block_mb is not None:
    mk.set_backward_block_mb(a.bwd_block_mb)
g = shaped_graph(a.workload, scale=a.scale, device="cuda")
n, e = g.num_nodes(), g.num_edges()
val = g.edge_weights("mean")
gen = torch.Generator(device="cuda").manual_seed(97)
x = torch.randn(n, a.dim, device="cuda", generator=gen)
dy = torch.randn(n, a.dim, device="cuda", generator=gen)
sd, si = mk.maxk_forward_cbsr(x, a.k)
banked = not a.plain and mk.banked_supported(a.k, a.dim)
packed = banked and a.k in (8, 16) and mk.packed_supported(a.k, a.dim)   # what the product runs at k = 8, 16
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for it in range(a.warmup + a.steps):
    ev[0].record()
    if packed:
        bp = mk.cbsr_bank_packed(sd, si, a.dim)
    elif banked:
        bd, bi, bs = mk.cbsr_bank(sd, si, a.dim)
    ev[1].record()
    if packed:
        out = mk.spgemm_forward_packed(g.indptr, g.indices, val, bp, n, e, a.k, a.dim)
    elif banked:
        out = mk.spgemm_forward_banked(g.indptr, g.indices, val, bd, bs, n, e, a.k, a.dim)
    else:
        out, _ = mk.spgemm_forward(g.indptr, g.indices, val, sd, si, n, e, a.k, a.dim)
    ev[2].record()
    if banked and a.banked_bwd and not packed:
        dxs = mk.spgemm_backward_banked(g.indptr, g.indices, val, dy, bs, n, e, a.k, a.dim)
    else:
        dxs = mk.spgemm_backward(g.indptr, g.indices, val, dy, si, n, e, a.k, a.dim)
    ev[3].record()
torch.cuda.synchronize()
print(f"{a.workload} N={n} E={e} k={a.k} D={a.dim} max_nz={mk.get_max_nz()} bwd_blocks={mk.backward_blocks(n, a.k, n, e)} "
      f"{'packed' if packed else 'banked' if banked else 'plain'}: bank {ev[0].elapsed_time(ev[1]):.3f} ms  "
      f"fwd {ev[1].elapsed_time(ev[2]):.3f} ms  bwd {ev[2].elapsed_time(ev[3]):.3f} ms")
