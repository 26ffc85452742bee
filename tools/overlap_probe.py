#!/usr/bin/env python
"""Probe: forward (L1TEX-data-pipe bound) and backward (SM->L2 request-path bound) on two streams."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import maxk_kernels as mk
from spgemm_gnn_b200.graph import shaped_graph
g = shaped_graph("reddit", device="cuda"); n, e = g.num_nodes(), g.num_edges(); val = g.edge_weights("mean")
gen = torch.Generator(device="cuda").manual_seed(97)
x = torch.randn(n, 256, device="cuda", generator=gen); dy = torch.randn(n, 256, device="cuda", generator=gen)
sd, si = mk.maxk_forward_cbsr(x, 32)
def fwd(): return mk.spgemm_forward(g.indptr, g.indices, val, sd, si, n, e, 32, 256)
def bwd(): return mk.spgemm_backward(g.indptr, g.indices, val, dy, si, n, e, 32, 256)
for _ in range(3): fwd(); bwd()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): fwd(); bwd()
b.record(); torch.cuda.synchronize(); serial = a.elapsed_time(b) / 10
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()
a.record(); s1.wait_event(a); s2.wait_event(a)
for _ in range(10):
    with torch.cuda.stream(s1): o = fwd()
    with torch.cuda.stream(s2): d = bwd()
torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
b.record(); torch.cuda.synchronize()
print(f"serial fwd+bwd {serial:.3f} ms; concurrent on two streams {a.elapsed_time(b) / 10:.3f} ms")
