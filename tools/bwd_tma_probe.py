#!/usr/bin/env python
"""Probe of the experimental backward (mk_sspmm_bwd_tma): parity against the shipped kernel and
time per launch for 1, 2, 4 of the 4 neighbours of a warp step going through bulk reductions.
    python tools/bwd_tma_probe.py [reddit|ogbn-products|...] [k]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import maxk_kernels as mk
from spgemm_gnn_b200.graph import shaped_graph, synthetic_graph

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
shape = sys.argv[1] if len(sys.argv) > 1 else "reddit"
k = int(sys.argv[2]) if len(sys.argv) > 2 else 32
d = 256


def run(g, name, reps):
    n, e = g.num_nodes(), g.num_edges()
    val = g.edge_weights("mean")
    gen = torch.Generator(device=dev).manual_seed(97)
    x = torch.randn(n, d, device=dev, generator=gen)
    dy = torch.randn(n, d, device=dev, generator=gen)
    _, si = mk.maxk_forward_cbsr(x, k)
    res = {}
    for nt in (0, 1, 2, 4):
        if nt > 128 // k:
            continue
        mk.set_backward_tma(nt)
        out = mk.spgemm_backward(g.indptr, g.indices, val, dy, si, n, e, k, d)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            mk.spgemm_backward(g.indptr, g.indices, val, dy, si, n, e, k, d)
        b.record()
        torch.cuda.synchronize()
        res[nt] = out
        diff = float(((out - res[0]).abs().max() / res[0].abs().max()).item())
        print(f"{name} k={k} tma_neighbours={nt}: {a.elapsed_time(b) / reps:.3f} ms per launch "
              f"(memset included), max rel diff vs shipped kernel {diff:.2e}", flush=True)
    mk.set_backward_tma(0)


run(synthetic_graph(20001, 20001 * 150, seed=97, device=dev), "small", 5)
run(shaped_graph(shape, device=dev), shape, 10)
