#!/usr/bin/env python
"""Forward / backward timings of ONE library build over a list of k (and work-record sizes), for
comparing experimental builds: `MAXK_LIB=<variant .so> python tools/variant_bench.py ...` -- one
process per variant (spgemm_gnn_b200/build.py --out=... -D...).  `--shard P` times rank 0's row
block of a P-way partition (global columns, no exchange): the per-rank kernel time of a P-GPU run."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import maxk_kernels as mk
from spgemm_gnn_b200 import dist as mdist
from spgemm_gnn_b200.graph import shaped_graph

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="reddit")
ap.add_argument("--ks", default="32")
ap.add_argument("--dim", type=int, default=256)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--shard", type=int, default=1)
ap.add_argument("--max-nz", default="")
ap.add_argument("--tag", default=os.path.basename(os.environ.get("MAXK_LIB", "product")))
ap.add_argument("--topk", action="store_true", help="time the MaxK top-k kernel too")
a = ap.parse_args()

g = shaped_graph(a.workload, device="cuda")
val = g.edge_weights("mean")
if a.shard > 1:
    local, r0, r1 = mdist.shard_graph(g, 0, a.shard)
    val = mdist.shard_edge_weights(g, local, r0, r1, "mean")
else:
    local = g
n_rows, n_src, e = local.num_nodes(), local.num_src, local.num_edges()
gen = torch.Generator(device="cuda").manual_seed(97)
x = torch.randn(n_src, a.dim, device="cuda", generator=gen)
dy = torch.randn(n_rows, a.dim, device="cuda", generator=gen)


def t(fn, reps=a.reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    s, f = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    f.record()
    torch.cuda.synchronize()
    return s.elapsed_time(f) / reps


for mz in ([int(v) for v in a.max_nz.split(",")] if a.max_nz else [mk.get_max_nz()]):
    mk.set_max_nz(mz)
    mk.clear_partition_cache()
    for k in (int(v) for v in a.ks.split(",")):
        sd, si = mk.maxk_forward_cbsr(x, k)
        part = mk.partition(local.indptr, n_rows)
        f = t(lambda: mk.spgemm_forward(local.indptr, local.indices, val, sd, si, n_rows, e, k, a.dim))
        b = t(lambda: mk.spgemm_backward(local.indptr, local.indices, val, dy, si, n_rows, e, k, a.dim))
        extra = ""
        if mk.use_banked(part.num_parts, e, k, a.dim):
            bd, _, bs = mk.cbsr_bank(sd, si, a.dim, with_index=False)
            fb = t(lambda: mk.spgemm_forward_banked(local.indptr, local.indices, val, bd, bs, n_rows, e, k, a.dim))
            extra = f" (banked kernel alone {fb:.3f})"
        if a.topk:
            extra += f" topk {t(lambda: mk.maxk_forward_cbsr(x, k)):.4f}"
            if mk.banked_supported(k, a.dim):
                extra += f" bank {t(lambda: mk.cbsr_bank(sd, si, a.dim, with_index=False)):.4f}"
                extra += f" topk+bank fused {t(lambda: mk.maxk_forward_cbsr_banked(x, k)):.4f}"
        print(f"[{a.tag}] {a.workload} shard 1/{a.shard} rows {n_rows} E {e} k {k} max_nz {mz} records {part.num_parts}: "
              f"fwd {f:.3f} ms{extra}  bwd {b:.3f} ms", flush=True)
