"""Why does the k = 16 forward measure 2.14-2.22 ms in tools/variant_bench.py and 2.39 ms inside bench.py's
ksweep?  Same call path; this probe varies what is alive in the allocator around the measurement."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import maxk_kernels as mk
from spgemm_gnn_b200.graph import shaped_graph

g = shaped_graph("reddit", device="cuda")
val = g.edge_weights("mean")
n, e, d = g.num_nodes(), g.num_edges(), 256


def t(fn, reps=10):
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


def fwd(x, k, label):
    sd, si = mk.maxk_forward_cbsr(x, k)
    f = t(lambda: mk.spgemm_forward(g.indptr, g.indices, val, sd, si, n, e, k, d))
    pk = mk.cbsr_bank_packed(sd, si, d) if k in (8, 16) else None
    extra = ""
    if pk is not None:
        out_ms = t(lambda: mk.spgemm_forward_packed(g.indptr, g.indices, val, pk, n, e, k, d))
        extra = f" (kernel on a resident packed table {out_ms:.3f})"
    print(f"{label}: k={k} fwd {f:.3f} ms{extra}  [allocated {torch.cuda.memory_allocated() >> 20} MB, reserved {torch.cuda.memory_reserved() >> 20} MB]", flush=True)


for seed in (97, 98):
    gen = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(n, d, device="cuda", generator=gen)
    fwd(x, 16, f"seed {seed} fresh")
    fwd(x, 8, f"seed {seed}")
    fwd(x, 16, f"seed {seed} after k=8")
    fwd(x, 32, f"seed {seed}")
    fwd(x, 16, f"seed {seed} after k=32")
    torch.cuda.empty_cache()
    fwd(x, 16, f"seed {seed} after empty_cache")
    junk = [torch.empty(200 << 20, dtype=torch.uint8, device="cuda") for _ in range(6)]
    fwd(x, 16, f"seed {seed} with 1.2 GB of other buffers alive")
    del junk
    dy = torch.randn(n, d, device="cuda", generator=gen)
    x2 = torch.randn(n, d, device="cuda", generator=gen)
    fwd(x2, 16, f"seed {seed} second x")
    del dy, x2
