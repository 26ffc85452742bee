"""Autograd Functions and conv layers of the reference's `utils/maxk_layers.py`, bodies
rewritten over the B200 kernels; class names, constructor arguments and call signatures kept.

What changed underneath:
  * CBSR (values + column ids) flows from the MaxK kernel straight into the SpGEMM kernel --
    no dense masked intermediate, no per-row Python `_extract_sparse_format` loop
    (utils/maxk_layers.py:224-265), no per-node `.item()` loop for the mean weights (:150-157);
  * the aggregation is an autograd Function whose backward IS `spgemm_backward` (the reference
    never calls it: its SpGEMM output is detached from autograd, utils/maxk_layers.py:166-171);
  * there is no DGL fallback and no CPU fallback: without the CUDA library every call raises.

`graph` is a `spgemm_gnn_b200.graph.CSRGraph` (destination-indexed CSR) or a DGLGraph, which
`graph.from_dgl` converts once and caches on the object; DGL itself is not needed.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn
from torch.autograd import Function

from . import maxk_kernels
from .graph import CSRGraph, from_dgl

KERNELS_AVAILABLE = True  # kept for callers that test it (maxk_gnn_integrated.py:24-31)

# f-3: top-k and banking as ONE kernel (mk_topk_cbsr_bank) on the single-GPU hot path.  Built, bit-identical
# to the two kernels, and measured slower on a B200 (profiles/r2/topk_tile.log: 0.272 ms against
# 0.164 + 0.064 ms on the Reddit shape -- both halves are bound by integer instruction issue, not by the
# 37 MB round trip the fusion removes), so it is opt-in.
FUSED_TOPK_BANK = os.environ.get("MAXK_FUSED_TOPK_BANK", "0") != "0"
# f-3: `h_self + aggregated -> LayerNorm` applied by the forward SpGEMM to the row it has just finished
FUSED_LN_EPILOGUE = os.environ.get("MAXK_FUSED_LN", "1") != "0"


# ---------------------------------------------------------------------------------------
# Linear layers: torch GEMMs (the contract of this tier), with ONE change in the backward.  The weight
# gradient dW = dY^T X reduces over all N nodes into an out x in matrix of a few tiles (256 x 256: four
# 128 x 128 tiles for 148 SMs), and cuBLAS runs it as such: 0.209 ms at N = 232,965 against 0.083 ms for the
# forward GEMM of the same size.  Cut into SPLITK_CHUNKS row blocks -- one batched GEMM, then a sum of the
# partial matrices -- it takes 0.094 ms (608 inputs: 0.462 -> 0.171 ms; products shape: 1.88 -> 0.71 ms;
# profiles/r2/dw_gemm_probe_call53.log), ~1.1 ms of a 32 ms MaxK-SAGE epoch on the Reddit shape.  Same
# module, same parameters and state dict as nn.Linear; MAXK_SPLITK_DW=0 restores autograd's own form.
# ---------------------------------------------------------------------------------------
SPLITK_DW = os.environ.get("MAXK_SPLITK_DW", "1") != "0"
SPLITK_MIN_ROWS = 8192


def weight_grad(grad_out: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """dW [out, in] = grad_out^T x for 2-D row-major operands, row-blocked when the reduction is long."""
    n = grad_out.shape[0]
    if not (SPLITK_DW and grad_out.is_cuda and n >= SPLITK_MIN_ROWS):
        return grad_out.t().mm(x)
    chunks = 32 if n >= (1 << 20) else 16
    rows = n // chunks
    main = rows * chunks
    g = grad_out.contiguous()
    xc = x.contiguous()
    dw = torch.bmm(g[:main].view(chunks, rows, g.shape[1]).transpose(1, 2),
                   xc[:main].view(chunks, rows, xc.shape[1])).sum(0)
    if main < n:
        dw = dw + g[main:].t().mm(xc[main:])
    return dw


class _LinearFunction(Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return torch.nn.functional.linear(x, weight, bias)

    @staticmethod
    def backward(ctx, grad_out):
        x, weight = ctx.saved_tensors
        gx = grad_out.mm(weight) if ctx.needs_input_grad[0] else None
        gw = weight_grad(grad_out, x) if ctx.needs_input_grad[1] else None
        gb = grad_out.sum(0) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        return gx, gw, gb


def linear(x: torch.Tensor, weight: torch.Tensor, bias=None) -> torch.Tensor:
    """`F.linear(x, weight, bias)` whose backward forms dW with `weight_grad`."""
    if x.dim() != 2 or not x.is_cuda or not SPLITK_DW:
        return torch.nn.functional.linear(x, weight, bias)
    return _LinearFunction.apply(x, weight, bias)


class Linear(nn.Linear):
    """nn.Linear (same parameters, initialisation and state dict) on `linear` above."""

    def forward(self, x):
        return linear(x, self.weight, self.bias)


# ---------------------------------------------------------------------------------------
# autograd Functions
# ---------------------------------------------------------------------------------------
class MaxKFunction(Function):
    """`MaxKFunction.apply(input, k)` -> dense [N,D] with everything but the k largest entries
    of each row zeroed (utils/maxk_layers.py:16-45, same contract as utils/models.py::MaxK).
    forward = top-k -> CBSR -> scatter; backward = gather at the kept columns -> scatter
    (== grad * mask, without ever storing the N x D mask)."""

    @staticmethod
    def forward(ctx, input, k=32):
        x = input.contiguous()
        sp_data, sp_index = maxk_kernels.maxk_forward_cbsr(x, k)
        ctx.save_for_backward(sp_index)
        ctx.dim_origin = x.shape[1]
        return maxk_kernels.cbsr_scatter(sp_data, sp_index, x.shape[1])

    @staticmethod
    def backward(ctx, grad_output):
        (sp_index,) = ctx.saved_tensors
        g = maxk_kernels.cbsr_gather(grad_output.contiguous(), sp_index)
        return maxk_kernels.cbsr_scatter(g, sp_index, ctx.dim_origin), None


class MaxKCBSRFunction(Function):
    """x [N,D] -> (sp_data [N,k], sp_index [N,k]).  backward: CBSR gradient -> dense."""

    @staticmethod
    def forward(ctx, input, k):
        x = input.contiguous()
        sp_data, sp_index = maxk_kernels.maxk_forward_cbsr(x, k)
        ctx.save_for_backward(sp_index)
        ctx.dim_origin = x.shape[1]
        ctx.mark_non_differentiable(sp_index)
        return sp_data, sp_index

    @staticmethod
    def backward(ctx, grad_data, _grad_index):
        (sp_index,) = ctx.saved_tensors
        return maxk_kernels.cbsr_scatter(grad_data.contiguous(), sp_index, ctx.dim_origin), None


class CBSRToDenseFunction(Function):
    """(sp_data, sp_index) -> dense masked [N,D]; backward gathers at the kept columns."""

    @staticmethod
    def forward(ctx, sp_data, sp_index, dim_origin):
        ctx.save_for_backward(sp_index)
        return maxk_kernels.cbsr_scatter(sp_data.contiguous(), sp_index, dim_origin)

    @staticmethod
    def backward(ctx, grad_dense):
        (sp_index,) = ctx.saved_tensors
        return maxk_kernels.cbsr_gather(grad_dense.contiguous(), sp_index), None, None


class SpGEMMFunction(Function):
    """out = A x Xs with A = CSR(ptr, idx, val) and Xs = CBSR(sp_data, sp_index).
    forward = `spgemm_forward`, backward (w.r.t. sp_data) = `spgemm_backward`."""

    @staticmethod
    def forward(ctx, sp_data, sp_index, ptr, idx, val, num_nodes, dim_origin):
        sp_data = sp_data.contiguous()
        k = sp_data.shape[1]
        out, _ = maxk_kernels.spgemm_forward(ptr, idx, val, sp_data, sp_index, num_nodes,
                                             idx.numel(), k, dim_origin)
        ctx.save_for_backward(sp_index, ptr, idx, val)
        ctx.meta = (num_nodes, k, dim_origin)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        sp_index, ptr, idx, val = ctx.saved_tensors
        num_nodes, k, dim_origin = ctx.meta
        dxs = maxk_kernels.spgemm_backward(ptr, idx, val, grad_out.contiguous(), sp_index,
                                           num_nodes, idx.numel(), k, dim_origin)
        return dxs, None, None, None, None, None, None


class MaxKAggregateFunction(Function):
    """feat [N,D] -> A x MaxK(feat) in two launches (f-3): `mk_topk_cbsr_bank` reads the dense row once
    and emits the sorted column ids + the banked table, the banked / packed forward SpGEMM consumes it.
    backward: `spgemm_backward` at the kept positions, then the CBSR gradient scattered to dense.
    Same values as MaxKCBSRFunction + SpGEMMFunction, bit for bit (the two kernels it fuses are)."""

    @staticmethod
    def forward(ctx, feat, k, ptr, idx, val, num_nodes, packed):
        x = feat.contiguous()
        d = x.shape[1]
        e = idx.numel()
        _, sp_index, table, bk_slot = maxk_kernels.maxk_forward_cbsr_banked(x, k, packed=packed)
        if packed:
            out = maxk_kernels.spgemm_forward_packed(ptr, idx, val, table, num_nodes, e, k, d)
        else:
            out = maxk_kernels.spgemm_forward_banked(ptr, idx, val, table, bk_slot, num_nodes, e, k, d)
        ctx.save_for_backward(sp_index, ptr, idx, val)
        ctx.meta = (num_nodes, k, d)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        sp_index, ptr, idx, val = ctx.saved_tensors
        num_nodes, k, d = ctx.meta
        dxs = maxk_kernels.spgemm_backward(ptr, idx, val, grad_out.contiguous(), sp_index, num_nodes,
                                           idx.numel(), k, d)
        return maxk_kernels.cbsr_scatter(dxs, sp_index, d), None, None, None, None, None, None


class MaxKAggregateLNFunction(Function):
    """y = LayerNorm(h_self + A x MaxK(h_neigh) + bias) * gamma + beta with the epilogue INSIDE the
    forward SpGEMM (mk_spgemm_fwd_banked_ln, f-3): the aggregated row never makes the round trip
    through memory (utils/maxk_layers.py:161-184 is top-k, SpGEMM, add, LayerNorm as separate passes).
    backward: fused LayerNorm backward -> gz (the gradient of h_self), SSpMM of gz at the kept
    positions -> gradient of h_neigh.  Same values as the unfused chain, bit for bit in the forward."""

    @staticmethod
    def forward(ctx, h_neigh, h_self, bias, gamma, beta, eps, k, ptr, idx, val, num_nodes):
        x = h_neigh.contiguous()
        d = x.shape[1]
        e = idx.numel()
        part = maxk_kernels.partition(ptr, num_nodes)
        packed = not maxk_kernels.use_banked(part.num_parts, e, k, d)
        sp_data, sp_index = maxk_kernels.maxk_forward_cbsr(x, k)
        if packed:
            table, slot = maxk_kernels.cbsr_bank_packed(sp_data, sp_index, d), None
        else:
            table, _, slot = maxk_kernels.cbsr_bank(sp_data, sp_index, d, with_index=False)
        keep = any(ctx.needs_input_grad[:5])
        y, z, mean, rstd = maxk_kernels.spgemm_forward_ln(
            ptr, idx, val, table, slot, num_nodes, e, k, d, None if h_self is None else h_self.contiguous(),
            None if bias is None else bias.contiguous(), gamma.contiguous(), beta.contiguous(), eps,
            keep_stats=keep)
        if keep:
            ctx.save_for_backward(z, gamma, mean, rstd, sp_index, ptr, idx, val)
        ctx.meta = (num_nodes, k, d, h_self is not None, bias is not None)
        return y

    @staticmethod
    def backward(ctx, grad_y):
        z, gamma, mean, rstd, sp_index, ptr, idx, val = ctx.saved_tensors
        num_nodes, k, d, has_self, has_bias = ctx.meta
        gz, dgamma, dbeta, dbias = maxk_kernels.layernorm_backward(grad_y.contiguous(), z, gamma, mean, rstd,
                                                                   want_dbias=has_bias)
        d_neigh = None
        if ctx.needs_input_grad[0]:
            dxs = maxk_kernels.spgemm_backward(ptr, idx, val, gz, sp_index, num_nodes, idx.numel(), k, d)
            d_neigh = maxk_kernels.cbsr_scatter(dxs, sp_index, d)
        return (d_neigh, gz if has_self else None, dbias, dgamma, dbeta, None, None, None, None, None, None)


def maxk_aggregate_add_norm(graph: CSRGraph, h_neigh: torch.Tensor, h_self, bias, norm, k: int,
                            weight_kind: str) -> torch.Tensor:
    """`norm(h_self + aggregate(MaxK(h_neigh)) + bias)` -- with the epilogue inside the forward SpGEMM
    where that exists and pays (single GPU, banked or packed forward, affine LayerNorm over <= 512 columns,
    k <= 32: at k = 64 the separate kernels measured faster, 5.765 against 5.916 ms on the Reddit shape,
    profiles/r2/fwd_per_width_call38.log), the separate kernels otherwise."""
    d = h_neigh.shape[1]
    if (FUSED_LN_EPILOGUE and getattr(graph, "world", 1) == 1 and h_neigh.is_cuda and h_neigh.dim() == 2
            and h_neigh.dtype == torch.float32 and isinstance(norm, nn.LayerNorm) and norm.elementwise_affine
            and norm.bias is not None and len(norm.normalized_shape) == 1 and d % 4 == 0 and d <= 512 and k <= 32
            and h_neigh.shape[0] == graph.num_src == graph.num_nodes() and maxk_kernels.banked_supported(k, d)):
        n, e = graph.num_nodes(), graph.num_edges()
        part = maxk_kernels.partition(graph.indptr, n)
        if maxk_kernels.use_banked(part.num_parts, e, k, d) or maxk_kernels.use_packed(part.num_parts, e, k, d):
            return MaxKAggregateLNFunction.apply(h_neigh, h_self, bias, norm.weight, norm.bias, norm.eps, k,
                                                 graph.indptr, graph.indices, graph.edge_weights(weight_kind), n)
    agg = maxk_aggregate(graph, h_neigh, k, weight_kind)
    if h_self is None:
        return add_layer_norm(agg, None, bias, norm)
    return add_layer_norm(h_self, agg, bias, norm)


class AddLayerNormFunction(Function):
    """y = LayerNorm(a + b + bias) * gamma + beta in one pass (f-3: the epilogue of the
    aggregation, utils/maxk_layers.py:174-182).  b and bias may be None."""

    @staticmethod
    def forward(ctx, a, b, bias, gamma, beta, eps):
        y, z, mean, rstd = maxk_kernels.add_layernorm_forward(
            a.contiguous(), None if b is None else b.contiguous(),
            None if bias is None else bias.contiguous(), gamma.contiguous(), beta.contiguous(), eps)
        ctx.save_for_backward(z, gamma, mean, rstd)
        ctx.has = (b is not None, bias is not None)
        return y

    @staticmethod
    def backward(ctx, grad_y):
        z, gamma, mean, rstd = ctx.saved_tensors
        has_b, has_bias = ctx.has
        gz, dgamma, dbeta, dbias = maxk_kernels.layernorm_backward(grad_y.contiguous(), z, gamma, mean,
                                                                   rstd, want_dbias=has_bias)
        return gz, (gz if has_b else None), dbias, dgamma, dbeta, None


def add_layer_norm(a, b, bias, norm):
    """`norm(a + b + bias)` -- fused when `norm` is an affine nn.LayerNorm over the last dim of CUDA
    float32 rows (what every model of the reference uses, utils/models.py:122), plain torch ops
    for any other `norm` module (BatchNorm in GNN_res, None)."""
    if (isinstance(norm, nn.LayerNorm) and norm.elementwise_affine and norm.bias is not None
            and a.dim() == 2 and len(norm.normalized_shape) == 1
            and maxk_kernels.add_layernorm_supported(a, a.shape[1])):
        return AddLayerNormFunction.apply(a, b, bias, norm.weight, norm.bias, norm.eps)
    out = a if b is None else a + b
    if bias is not None:
        out = out + bias
    return out if norm is None else norm(out)


def aggregate_cbsr(graph: CSRGraph, sp_data, sp_index, weight_kind: str, dim_origin: int):
    """A x Xs for a CBSR table; on a `dist.ShardedGraph` the table is row-sharded and the call
    includes the all-gather (forward) and the reduce-scatter (backward)."""
    val = graph.edge_weights(weight_kind)
    if getattr(graph, "world", 1) > 1:
        from .dist import DistSpGEMMFunction
        return DistSpGEMMFunction.apply(sp_data, sp_index, graph.indptr, graph.indices, val,
                                        graph.num_nodes(), dim_origin, graph.group)
    return SpGEMMFunction.apply(sp_data, sp_index, graph.indptr, graph.indices, val,
                                graph.num_nodes(), dim_origin)


def maxk_aggregate(graph: CSRGraph, feat: torch.Tensor, k: int, weight_kind: str) -> torch.Tensor:
    """MaxK -> CBSR -> SpGEMM in one go: sum_j w(i<-j) * maxk(feat)[j].  The hot path."""
    if (FUSED_TOPK_BANK and getattr(graph, "world", 1) == 1 and feat.is_cuda and feat.dim() == 2
            and feat.dtype == torch.float32):
        n, e, d = graph.num_nodes(), graph.num_edges(), feat.shape[1]
        if feat.shape[0] == graph.num_src and maxk_kernels.banked_supported(k, d):
            part = maxk_kernels.partition(graph.indptr, n)
            banked = maxk_kernels.use_banked(part.num_parts, e, k, d)
            if banked or maxk_kernels.use_packed(part.num_parts, e, k, d):   # top-k + banking fused
                return MaxKAggregateFunction.apply(feat, k, graph.indptr, graph.indices,
                                                   graph.edge_weights(weight_kind), n, not banked)
    sp_data, sp_index = MaxKCBSRFunction.apply(feat, k)
    return aggregate_cbsr(graph, sp_data, sp_index, weight_kind, feat.shape[1])


def extract_sparse_format(sparse_tensor: torch.Tensor, maxk: int):
    """`MaxKSAGEConv._extract_sparse_format` (utils/maxk_layers.py:224-265, a per-row Python loop
    there): CBSR of a dense masked matrix -- the first `maxk` non-zeros of every row in ascending
    column order, padded with (0.0, index 0).  Kept for callers of the reference method; the hot
    path never needs it because the MaxK kernel emits CBSR directly.  Plain torch ops."""
    n, d = sparse_tensor.shape
    nz = sparse_tensor != 0
    order = torch.argsort((~nz).to(torch.uint8), dim=1, stable=True)[:, :maxk]
    vals = torch.gather(sparse_tensor, 1, order)
    keep = torch.arange(maxk, device=sparse_tensor.device)[None, :] < nz.sum(1, keepdim=True)
    idx_dtype = torch.uint8 if d <= 256 else torch.int16
    sp_index = torch.where(keep, order, torch.zeros_like(order)).to(idx_dtype)
    if d > 256:
        sp_index = sp_index.view(torch.uint16)
    return torch.where(keep, vals, torch.zeros_like(vals)), sp_index


def _dense_aggregate(graph: CSRGraph, feat: torch.Tensor, weight_kind: str) -> torch.Tensor:
    """Non-MaxK (`--nonlinear relu`) branch: dense SpMM through cuSPARSE, the comparator the
    reference reports its speed-ups against (README.md:136).  Not the hot path."""
    if getattr(graph, "world", 1) > 1:  # comparator path only: gather the dense rows
        import torch.distributed as dist
        full = torch.empty((graph.num_src, feat.shape[1]), dtype=feat.dtype, device=feat.device)
        dist.all_gather_into_tensor(full, feat.contiguous(), group=graph.group)
        feat = full
    key = ("adj", weight_kind)
    adj = graph._cache.get(key)
    if adj is None:
        adj = torch.sparse_csr_tensor(graph.indptr.to(torch.int64), graph.indices.to(torch.int64),
                                      graph.edge_weights(weight_kind),
                                      size=(graph.num_nodes(), graph.num_src))
        graph._cache[key] = adj
    return torch.sparse.mm(adj, feat)


# ---------------------------------------------------------------------------------------
# layers (utils/maxk_layers.py:47-265, 267-447; utils/integrated_models.py:221-270)
# ---------------------------------------------------------------------------------------
class MaxKSAGEConv(nn.Module):
    """h_self + aggregate(MaxK(fc_neigh(feat))), then norm and dropout
    (utils/maxk_layers.py:82-99, 161-184)."""

    def __init__(self, in_feats, out_feats, aggregator_type="mean", feat_drop=0.0, norm=None, maxk=32):
        super().__init__()
        if aggregator_type not in ("mean", "sum"):
            raise ValueError(f"Unsupported aggregator type: {aggregator_type}")
        self.in_feats = in_feats
        self.out_feats = out_feats
        self.aggregator_type = aggregator_type
        self.maxk = maxk
        self.fc_self = Linear(in_feats, out_feats, bias=False)
        self.fc_neigh = Linear(in_feats, out_feats, bias=False)
        self.norm = norm
        if norm is not None and isinstance(norm, type):
            self.norm = norm(out_feats)
        self.feat_drop = nn.Dropout(feat_drop)
        self.maxk_fn = MaxKFunction.apply
        self.reset_parameters()

    def reset_parameters(self):
        gain = nn.init.calculate_gain("relu")
        nn.init.xavier_uniform_(self.fc_self.weight, gain=gain)
        nn.init.xavier_uniform_(self.fc_neigh.weight, gain=gain)

    def _extract_sparse_format(self, sparse_tensor):
        return extract_sparse_format(sparse_tensor, self.maxk)

    def forward(self, graph, feat):
        graph = from_dgl(graph)       # CSRGraph passes through; a DGLGraph is converted once and cached
        h_self = self.fc_self(feat)
        h_neigh = self.fc_neigh(feat)
        output = maxk_aggregate_add_norm(graph, h_neigh, h_self, None, self.norm, self.maxk,
                                         self.aggregator_type)
        return self.feat_drop(output)


class MaxKGCNConv(nn.Module):
    """D_in^-1/2 A D_out^-1/2 MaxK(feat W) + b (utils/maxk_layers.py:300-324, 370-390); both
    normalisations are folded into the per-edge weights once per graph.

    `norm='both'` follows `dglnn.GraphConv(norm='both')` -- the layer the reference TRAINS with
    (utils/models.py:252).  The reference's own MaxKGCNConv class puts both degree factors on the
    source node instead (A D_out^-1/2 D_in^-1/2, utils/maxk_layers.py:315-318, 372-376);
    `reference_norm=True` reproduces that class exactly (edge weights `'reference_gcn'`), for users who
    need its numbers rather than GraphConv's."""

    def __init__(self, in_feats, out_feats, norm="both", weight=True, bias=True,
                 allow_zero_in_degree=False, maxk=32, reference_norm=False):
        super().__init__()
        self.reference_norm = bool(reference_norm)
        if norm not in ("none", "both", "right"):
            raise ValueError(f"Unsupported norm: {norm}")
        self.in_feats = in_feats
        self.out_feats = out_feats
        self.norm = norm
        self.maxk = maxk
        self.allow_zero_in_degree = allow_zero_in_degree
        if weight:
            self.weight = nn.Parameter(torch.empty(in_feats, out_feats))
        else:
            self.register_parameter("weight", None)
        if bias:
            self.bias = nn.Parameter(torch.empty(out_feats))
        else:
            self.register_parameter("bias", None)
        self.maxk_fn = MaxKFunction.apply
        self.reset_parameters()

    def reset_parameters(self):
        if self.weight is not None:
            nn.init.xavier_uniform_(self.weight)
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def _extract_sparse_format(self, sparse_tensor):
        return extract_sparse_format(sparse_tensor, self.maxk)

    def forward(self, graph, feat):
        graph = from_dgl(graph)
        if not self.allow_zero_in_degree:
            zero = graph._cache.get("has_zero_in")
            if zero is None:
                zero = bool((graph.in_degrees() == 0).any())
                graph._cache["has_zero_in"] = zero
            if zero:
                raise ValueError("Graph has nodes with zero in-degree")
        if self.weight is not None:
            feat = torch.mm(feat, self.weight)
        kind = "reference_gcn" if (self.reference_norm and self.norm == "both") else self.norm
        output = maxk_aggregate(graph, feat, self.maxk, kind)
        if self.bias is not None:
            output = output + self.bias
        return output


class MaxKGINConv(nn.Module):
    """mlp((1 + eps) * feat + sum_j MaxK(feat)[j]) (utils/integrated_models.py:221-270, where
    the sum is a DGL `update_all(copy_u, sum)` on the dense masked features)."""

    def __init__(self, in_feats, out_feats, learn_eps=True, maxk=32):
        super().__init__()
        self.in_feats = in_feats
        self.out_feats = out_feats
        self.maxk = maxk
        if learn_eps:
            self.eps = nn.Parameter(torch.zeros(1))
        else:
            self.register_buffer("eps", torch.zeros(1))
        self.mlp = nn.Sequential(Linear(in_feats, out_feats), nn.ReLU(),
                                 Linear(out_feats, out_feats))
        self.maxk_fn = MaxKFunction.apply
        self.reset_parameters()

    def reset_parameters(self):
        for layer in self.mlp:
            if isinstance(layer, nn.Linear):
                nn.init.xavier_uniform_(layer.weight)

    def forward(self, graph, feat):
        graph = from_dgl(graph)
        neigh = maxk_aggregate(graph, feat, self.maxk, "sum")
        output = (1 + self.eps) * feat + neigh
        return self.mlp(output)
