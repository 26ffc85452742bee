#!/usr/bin/env python
"""Generates tests/golden/layers_reference.npz by running the REFERENCE'S OWN layer code
(`utils/maxk_layers.py::MaxKSAGEConv`, `MaxKGCNConv`; `utils/integrated_models.py::MaxKGINConv` at the
end of main) in the build container, the first two both ways:

  (A) the branch that trains in the reference's logs: `_aggregate_with_dgl`
      (utils/maxk_layers.py:208-222, 392-405) -> `graph.update_all(fn.copy_u, fn.mean|sum)`;
  (B) the reference's own CALL SITE of the native kernel: `_aggregate_with_custom_kernel`
      (utils/maxk_layers.py:100-184, 326-390) -- its CSR extraction, its per-edge weights, its
      `_extract_sparse_format`, and its positional call
      `maxk_kernels.spgemm_forward(ptr, idx, edge_weights, sp_data, sp_index, N, E, k, D)`.

DGL is not installed and the reference's compiled `maxk_kernels` cannot be imported (cp39, sm_80
SASS), so two stand-ins are injected, both marked below and both as small as the reference's use
of them allows:

  * `FakeDGLGraph`: the handful of DGLGraph members these two layers touch.  Its `update_all`
    IS the definition of DGL's built-ins `copy_u` + `sum` / `mean` over in-edges (a float64
    sparse product, differentiable): out[v] = sum (or mean) over edges u->v of h[u]; zero for
    nodes without in-edges.
  * `RecordingKernels.spgemm_forward`: records the arguments the reference passes and answers
    with the numpy oracle (oracle/maxk_oracle.py::spgemm_fwd).

What this pins: for the same weights and inputs, branch (B) == branch (A) only if `spgemm_forward`
means what the oracle says it means, called the way the reference calls it.  The SAGE cases also
record the BACKWARD of branch (A): the gradient that reaches the input of the MaxK for a seeded dY
(autograd through `update_all` and the fallback `MaxKFunction.backward`, utils/maxk_layers.py:37-45)
-- what `spgemm_backward` followed by the CBSR scatter has to equal.  The vectors (recorded
arguments + the output of branch A) are what the CUDA path is then held to
(tests/test_oracle.py, tests/test_gpu_parity.py).  What it does not pin: DGL's float32 summation
order -- nothing in /root/reference can.

Run once in the build container (needs /root/reference):  python tests/golden/make_golden_layers.py
"""
import contextlib
import os
import sys
import types

import numpy as np
import scipy.sparse as sp
import torch
import torch.nn as nn

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


# ------------------------------------------------------------------------- stand-in: dgl
def _stub_dgl():
    dgl = types.ModuleType("dgl")
    dgl_nn = types.ModuleType("dgl.nn")
    dgl_fn = types.ModuleType("dgl.function")
    dgl_nn_pt = types.ModuleType("dgl.nn.pytorch")
    dgl_nn_conv = types.ModuleType("dgl.nn.pytorch.conv")
    for name in ("SAGEConv", "GraphConv", "GINConv"):
        cls = type(name, (nn.Module,), {})
        setattr(dgl_nn, name, cls)
        setattr(dgl_nn_conv, name, cls)
    # dgl.function built-ins, reduced to descriptions FakeDGLGraph.update_all understands
    dgl_fn.copy_u = lambda u, out: ("copy_u", u, out)
    dgl_fn.sum = lambda msg, out: ("sum", msg, out)
    dgl_fn.mean = lambda msg, out: ("mean", msg, out)
    dgl_nn.pytorch = dgl_nn_pt
    dgl_nn_pt.conv = dgl_nn_conv
    dgl.nn = dgl_nn
    dgl.function = dgl_fn
    sys.modules.update({"dgl": dgl, "dgl.nn": dgl_nn, "dgl.function": dgl_fn,
                        "dgl.nn.pytorch": dgl_nn_pt, "dgl.nn.pytorch.conv": dgl_nn_conv})
    sys.modules["maxk_kernels"] = None      # the import in utils/maxk_layers.py:10 must fail


class FakeDGLGraph:
    """The DGLGraph surface utils/maxk_layers.py touches.  Edges u -> v; `in_csr` is indexed by
    destination v (what update_all reduces over), `out_csr` by source u (what DGL's
    adj_tensors('csr') returns)."""

    def __init__(self, src, dst, n):
        self.n, self.e = int(n), int(len(src))
        self.device = torch.device("cpu")
        ones = np.ones(self.e)
        self.in_csr = sp.csr_matrix((ones, (dst, src)), shape=(n, n))
        self.in_csr.sort_indices()
        self.out_csr = sp.csr_matrix((ones, (src, dst)), shape=(n, n))
        self.out_csr.sort_indices()
        self.ndata = {}
        self._sparse_format = True          # maxk_gnn_integrated.py:77-135 attaches this attribute

    @contextlib.contextmanager
    def local_scope(self):
        saved = dict(self.ndata)
        try:
            yield
        finally:
            self.ndata = saved

    def num_nodes(self):
        return self.n

    def num_edges(self):
        return self.e

    def in_degrees(self):
        return torch.from_numpy(np.diff(self.in_csr.indptr).astype(np.int64))

    def out_degrees(self):
        return torch.from_numpy(np.diff(self.out_csr.indptr).astype(np.int64))

    def adj_tensors(self, fmt):
        assert fmt == "csr"
        return (torch.from_numpy(self.out_csr.indptr.astype(np.int64)),
                torch.from_numpy(self.out_csr.indices.astype(np.int64)),
                torch.arange(self.e))

    def update_all(self, message, reduce):
        kind_m, u_field, msg_field = message
        kind_r, msg_field_r, out_field = reduce
        assert kind_m == "copy_u" and msg_field == msg_field_r and kind_r in ("sum", "mean")
        # float64 sparse product, differentiable: the backward of the reference's DGL branch is
        # whatever autograd makes of "sum of the in-neighbours' rows"
        adj = torch.sparse_csr_tensor(torch.from_numpy(self.in_csr.indptr.astype(np.int64)),
                                      torch.from_numpy(self.in_csr.indices.astype(np.int64)),
                                      torch.ones(self.e, dtype=torch.float64), size=(self.n, self.n))
        out = torch.sparse.mm(adj, self.ndata[u_field].double())
        if kind_r == "mean":
            out = out / self.in_degrees().clamp(min=1).double()[:, None]
        self.ndata[out_field] = out.float()


# ------------------------------------------------------------------ stand-in: maxk_kernels
class RecordingKernels:
    def __init__(self, oracle):
        self.oracle, self.calls = oracle, []

    def spgemm_forward(self, ptr, idx, val, sp_data, sp_index, num_nodes, num_edges, dim_sparse,
                       dim_origin):
        rec = dict(ptr=ptr.numpy().astype(np.int32), idx=idx.numpy().astype(np.int32),
                   val=val.numpy().astype(np.float32), sp_data=sp_data.detach().numpy().astype(np.float32),
                   sp_index=sp_index.numpy().astype(np.uint8), num_nodes=int(num_nodes),
                   num_edges=int(num_edges), dim_sparse=int(dim_sparse), dim_origin=int(dim_origin))
        self.calls.append(rec)
        out = self.oracle.spgemm_fwd(rec["ptr"], rec["idx"], rec["val"], rec["sp_data"],
                                     rec["sp_index"], rec["dim_origin"])
        return torch.from_numpy(out.astype(np.float32)), sp_index


def symmetric_graph(n, avg_deg, rng):
    """Random undirected graph with one self-loop per node (the reference adds them:
    maxk_gnn_dgl.py:221-223), as directed edges in both directions."""
    m = n * avg_deg // 2
    a, b = rng.integers(0, n, m), rng.integers(0, n, m)
    keep = a != b
    pairs = np.unique(np.stack([np.minimum(a, b)[keep], np.maximum(a, b)[keep]], 1), axis=0)
    src = np.concatenate([pairs[:, 0], pairs[:, 1], np.arange(n)])
    dst = np.concatenate([pairs[:, 1], pairs[:, 0], np.arange(n)])
    return src, dst


def main():
    if not os.path.isdir(REF):
        sys.exit("needs /root/reference (build container only)")
    sys.path.insert(0, ROOT)
    from oracle import maxk_oracle as mo        # the checker, used as the kernel stand-in only
    _stub_dgl()
    sys.path = [p for p in sys.path if os.path.abspath(p or ".") != ROOT]
    sys.path.insert(0, REF)
    from utils import maxk_layers as ref        # noqa: E402

    assert ref.KERNELS_AVAILABLE is False
    out = {}
    rng = np.random.default_rng(97)
    torch.manual_seed(97)
    cases = [("sage_mean", 200, 12, 64, 64, 16), ("sage_sum", 120, 8, 96, 128, 32),
             ("gcn_both", 150, 10, 48, 64, 8), ("sage_mean_wide", 90, 20, 128, 256, 32)]
    for name, n, deg, d_in, d_out, k in cases:
        src, dst = symmetric_graph(n, deg, rng)
        g = FakeDGLGraph(src, dst, n)
        feat = torch.randn(n, d_in)
        rec = RecordingKernels(mo)
        if name.startswith("sage"):
            conv = ref.MaxKSAGEConv(d_in, d_out, aggregator_type=name.split("_")[1], maxk=k)
            with torch.no_grad():
                y_dgl = conv(g, feat)                                   # branch (A)
                h_self, h_neigh = conv.fc_self(feat), conv.fc_neigh(feat)
                h_sparse = ref.MaxKFunction.apply(h_neigh, k)           # fallback MaxK (pinned already)
                ref.maxk_kernels, ref.KERNELS_AVAILABLE = rec, True
                try:
                    y_custom = conv._aggregate_with_custom_kernel(g, h_self, h_sparse)   # branch (B)
                finally:
                    ref.KERNELS_AVAILABLE = False
                    del ref.maxk_kernels
            # backward of branch (A): gradient reaching the input of the MaxK, for a seeded dY
            grabbed = []

            def grab(_module, _inputs, output):
                output.retain_grad()
                grabbed.append(output)

            hook = conv.fc_neigh.register_forward_hook(grab)
            y_again = conv(g, feat)
            hook.remove()
            dy = torch.randn(n, d_out)
            y_again.backward(dy)
            assert torch.equal(y_again.detach(), y_dgl)
            out[f"{name}_dy"] = dy.numpy()
            out[f"{name}_grad_maxk_in"] = grabbed[0].grad.numpy()
            out[f"{name}_h_self"] = h_self.numpy()
        else:
            conv = ref.MaxKGCNConv(d_in, d_out, norm="both", maxk=k)
            nn.init.normal_(conv.bias, std=0.1)
            with torch.no_grad():
                y_dgl = conv(g, feat)                                   # branch (A)
                h = torch.mm(feat, conv.weight)
                h_sparse = ref.MaxKFunction.apply(h, k)
                h_sparse = h_sparse * torch.pow(g.out_degrees().float().clamp(min=1), -0.5).unsqueeze(1)
                ref.maxk_kernels, ref.KERNELS_AVAILABLE = rec, True
                try:
                    y_custom = conv._aggregate_with_custom_kernel(g, h_sparse)           # branch (B)
                finally:
                    ref.KERNELS_AVAILABLE = False
                    del ref.maxk_kernels
            out[f"{name}_bias"] = conv.bias.detach().numpy()
        assert len(rec.calls) == 1, "the reference layer must reach its spgemm_forward call"
        call = rec.calls[0]
        err = float((y_custom - y_dgl).abs().max() / y_dgl.abs().max())
        assert err < 2e-6, f"{name}: custom-kernel branch differs from the DGL branch by {err:.2e}"
        # the graph is symmetric, so the source-indexed CSR the reference extracts equals the
        # destination-indexed one the aggregation needs
        assert np.array_equal(call["ptr"], g.in_csr.indptr) and np.array_equal(call["idx"], g.in_csr.indices)
        out[f"{name}_k"] = np.int64(k)
        out[f"{name}_y"] = y_dgl.numpy()
        for key in ("ptr", "idx", "val", "sp_data", "sp_index"):
            out[f"{name}_call_{key}"] = call[key]
        out[f"{name}_call_dims"] = np.array([call["num_nodes"], call["num_edges"], call["dim_sparse"],
                                             call["dim_origin"]], dtype=np.int64)
        print(f"{name}: N={n} E={g.e} {d_in}->{d_out} k={k}: branch (B) vs (A) max rel diff {err:.2e}")
    # utils/integrated_models.py::MaxKGINConv (221-270): (1 + eps) * feat + sum_j MaxK(feat)[j] -> mlp.
    # The module uses `fn.copy_u` / `fn.sum` without importing `dgl.function as fn` (NameError as
    # shipped); the name is injected here, nothing else is touched.  No GEMM precedes the MaxK, so
    # the layer is reproducible bit for bit up to the aggregation and can be pinned END TO END.
    from utils import integrated_models as ref_int   # noqa: E402
    ref_int.fn = sys.modules["dgl.function"]
    n, deg, d_in, d_out, k = 160, 14, 128, 64, 32
    src, dst = symmetric_graph(n, deg, rng)
    g = FakeDGLGraph(src, dst, n)
    feat = torch.randn(n, d_in)
    conv = ref_int.MaxKGINConv(d_in, d_out, learn_eps=True, maxk=k)
    with torch.no_grad():
        conv.eps.fill_(0.25)
        for layer in conv.mlp:
            if isinstance(layer, nn.Linear):
                nn.init.normal_(layer.bias, std=0.1)
    pre = []
    conv.mlp.register_forward_pre_hook(lambda _m, inp: pre.append(inp[0].detach().clone()))
    with torch.no_grad():
        y = conv(g, feat)
    out["gin_ptr"] = g.in_csr.indptr.astype(np.int32)
    out["gin_idx"] = g.in_csr.indices.astype(np.int32)
    out["gin_feat"] = feat.numpy()
    out["gin_dims"] = np.array([n, d_in, d_out, k], dtype=np.int64)
    out["gin_pre_mlp"] = pre[0].numpy()
    out["gin_y"] = y.numpy()
    for key, val in conv.state_dict().items():
        out["gin_sd_" + key] = val.numpy()
    print(f"gin: N={n} E={g.e} {d_in}->{d_out} k={k}: state dict {sorted(conv.state_dict())}")
    # MaxKSAGEConv END TO END: with scaled permutation matrices as weights both Linear layers are
    # exact on any hardware (one non-zero product per output), so the MaxK selects the same entries
    # everywhere and the whole layer (utils/maxk_layers.py:82-99 + 208-222) can be compared as one.
    n, deg, d, k = 140, 16, 64, 16
    src, dst = symmetric_graph(n, deg, rng)
    g = FakeDGLGraph(src, dst, n)
    feat = torch.randn(n, d)
    conv = ref.MaxKSAGEConv(d, d, aggregator_type="mean", maxk=k)
    with torch.no_grad():
        conv.fc_self.weight.copy_(0.5 * torch.eye(d)[torch.randperm(d)])
        conv.fc_neigh.weight.copy_(2.0 * torch.eye(d)[torch.randperm(d)])
        y = conv(g, feat)
    out["sagex_ptr"] = g.in_csr.indptr.astype(np.int32)
    out["sagex_idx"] = g.in_csr.indices.astype(np.int32)
    out["sagex_feat"] = feat.numpy()
    out["sagex_dims"] = np.array([n, d, k], dtype=np.int64)
    out["sagex_y"] = y.numpy()
    for key, val in conv.state_dict().items():
        out["sagex_sd_" + key] = val.numpy()
    print(f"sage end to end: N={n} E={g.e} d={d} k={k}: state dict {sorted(conv.state_dict())}")
    out["names"] = np.array([c[0] for c in cases])
    path = os.path.join(HERE, "layers_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", "torch", torch.__version__)


if __name__ == "__main__":
    main()
