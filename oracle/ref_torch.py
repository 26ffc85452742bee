"""CPU ORACLE (test infrastructure, NOT product code) -- the reference's TRAINING path
restated in plain torch: `torch.topk` -> 0/1 mask -> multiply for MaxK
(utils/models.py:12-26) and a CSR SpMM in the place of DGL's
`graph.update_all(copy_u, mean|sum)` (utils/models.py:72,163,284,407), forward and
backward through autograd.

This is what `bench.py --impl reference` and the `cpu_baseline` object time on the host
cores ("port": DGL itself is not in the image, `torch.sparse.mm` on a sparse_csr tensor
stands in for its CPU SpMM, as BASELINE.md section 4 lays down), and what the loss-curve
parity test trains against.  It is also the second opinion for the numpy/C oracle on
tie-free inputs.
"""
from __future__ import annotations

import warnings

import torch
import torch.nn as nn
import torch.nn.functional as F

warnings.filterwarnings("ignore", message=".*Sparse CSR tensor support is in beta.*")
warnings.filterwarnings("ignore", message=".*Sparse invariant checks are implicitly disabled.*")


class RefMaxK(torch.autograd.Function):
    """utils/models.py:12-26."""

    @staticmethod
    def forward(ctx, inp, k=1):
        _, sel = inp.topk(k, dim=1)
        mask = torch.zeros_like(inp)
        mask.scatter_(1, sel, 1)
        ctx.save_for_backward(mask)
        return inp * mask

    @staticmethod
    def backward(ctx, grad_output):
        (mask,) = ctx.saved_tensors
        return grad_output * mask, None


def csr_matrix(indptr: torch.Tensor, indices: torch.Tensor, val: torch.Tensor, num_src: int):
    n = indptr.numel() - 1
    return torch.sparse_csr_tensor(
        indptr.to(torch.int64), indices.to(torch.int64), val, size=(n, num_src)
    )


def aggregate(adj, x):
    """DGL `update_all(copy_u('h','m'), sum('m','neigh'))` with the edge weight folded
    into `adj` (mean = 1/deg_in rows, GCN = both-side normalisation, GIN = ones)."""
    return torch.sparse.mm(adj, x)


def layer_forward_backward(adj, x, grad_y, k):
    """One MaxK + aggregation layer, forward and backward.  Returns (Y, dX)."""
    x = x.detach().requires_grad_(True)
    y = aggregate(adj, RefMaxK.apply(x, k))
    y.backward(grad_y)
    return y.detach(), x.grad


# -- the three model families of utils/models.py with the DGL conv restated ---------------
class RefSAGEConv(nn.Module):
    """dgl.nn.SAGEConv(in, out, 'mean', feat_drop, norm=...) as the reference uses it
    (utils/models.py:140): in == out so DGL aggregates first, then fc_neigh; bias on the
    sum; `rst = fc_self(h) + fc_neigh(mean_neigh(h)) + bias`, then norm."""

    def __init__(self, in_feats, out_feats, feat_drop=0.0, norm=None):
        super().__init__()
        self.feat_drop = nn.Dropout(feat_drop)
        self.fc_self = nn.Linear(in_feats, out_feats, bias=False)
        self.fc_neigh = nn.Linear(in_feats, out_feats, bias=False)
        self.bias = nn.Parameter(torch.zeros(out_feats))
        self.norm = norm
        gain = nn.init.calculate_gain("relu")
        nn.init.xavier_uniform_(self.fc_self.weight, gain=gain)
        nn.init.xavier_uniform_(self.fc_neigh.weight, gain=gain)

    def forward(self, adj_mean, feat):
        h = self.feat_drop(feat)
        rst = self.fc_self(h) + self.fc_neigh(aggregate(adj_mean, h)) + self.bias
        if self.norm is not None:
            rst = self.norm(rst)
        return rst


class RefSAGE(nn.Module):
    """utils/models.py:109-166."""

    def __init__(self, in_size, hid_size, num_hid_layers, out_size, maxk=32, feat_drop=0.5,
                 norm=False, nonlinear="maxk"):
        super().__init__()
        self.layers = nn.ModuleList()
        self.num_layers = num_hid_layers
        for _ in range(num_hid_layers):
            nl = nn.LayerNorm(hid_size, elementwise_affine=True) if norm else None
            self.layers.append(RefSAGEConv(hid_size, hid_size, feat_drop=feat_drop, norm=nl))
        self.lin_in = nn.Linear(in_size, hid_size)
        self.lin_out = nn.Linear(hid_size, out_size)
        nn.init.xavier_uniform_(self.lin_in.weight)
        nn.init.xavier_uniform_(self.lin_out.weight)
        self.k = maxk
        self.nonlinear = nonlinear

    def forward(self, adj_mean, x):
        x = self.lin_in(x)
        for i in range(self.num_layers):
            if self.nonlinear == "maxk":
                x = RefMaxK.apply(x, self.k)
            elif self.nonlinear == "relu":
                x = F.relu(x)
            x = self.layers[i](adj_mean, x)
        return self.lin_out(x)


class RefGCN(nn.Module):
    """utils/models.py:240-288: lin -> MaxK -> dropout -> GraphConv(norm='both',
    weight=None, bias) -> LayerNorm."""

    def __init__(self, in_size, hid_size, num_hid_layers, out_size, maxk=32, feat_drop=0.5,
                 norm=False, nonlinear="maxk"):
        super().__init__()
        self.num_layers = num_hid_layers
        self.norm = norm
        self.dropoutlayers = nn.ModuleList(nn.Dropout(feat_drop) for _ in range(num_hid_layers))
        self.gcn_bias = nn.ParameterList(nn.Parameter(torch.zeros(hid_size)) for _ in range(num_hid_layers))
        self.normlayers = nn.ModuleList(
            nn.LayerNorm(hid_size, elementwise_affine=True) for _ in range(num_hid_layers if norm else 0)
        )
        self.linlayers = nn.ModuleList(nn.Linear(hid_size, hid_size) for _ in range(num_hid_layers))
        for lin in self.linlayers:
            nn.init.xavier_uniform_(lin.weight)
        self.lin_in = nn.Linear(in_size, hid_size)
        self.lin_out = nn.Linear(hid_size, out_size)
        nn.init.xavier_uniform_(self.lin_in.weight)
        nn.init.xavier_uniform_(self.lin_out.weight)
        self.k = maxk
        self.nonlinear = nonlinear

    def forward(self, adj_both, x):
        x = self.lin_in(x).relu()
        for i in range(self.num_layers):
            x = self.linlayers[i](x)
            if self.nonlinear == "maxk":
                x = RefMaxK.apply(x, self.k)
            elif self.nonlinear == "relu":
                x = F.relu(x)
            x = self.dropoutlayers[i](x)
            x = aggregate(adj_both, x) + self.gcn_bias[i]
            if self.norm:
                x = self.normlayers[i](x)
        return self.lin_out(x)


class RefGIN(nn.Module):
    """utils/models.py:363-411: GINConv(apply_func=None, 'sum', learn_eps=True):
    rst = (1 + eps) * h + sum_neigh(h)."""

    def __init__(self, in_size, hid_size, num_hid_layers, out_size, maxk=32, feat_drop=0.5,
                 norm=False, nonlinear="maxk"):
        super().__init__()
        self.num_layers = num_hid_layers
        self.norm = norm
        self.dropoutlayers = nn.ModuleList(nn.Dropout(feat_drop) for _ in range(num_hid_layers))
        self.eps = nn.ParameterList(nn.Parameter(torch.zeros(1)) for _ in range(num_hid_layers))
        self.normlayers = nn.ModuleList(
            nn.LayerNorm(hid_size, elementwise_affine=True) for _ in range(num_hid_layers if norm else 0)
        )
        self.linlayers = nn.ModuleList(nn.Linear(hid_size, hid_size) for _ in range(num_hid_layers))
        for lin in self.linlayers:
            nn.init.xavier_uniform_(lin.weight)
        self.lin_in = nn.Linear(in_size, hid_size)
        self.lin_out = nn.Linear(hid_size, out_size)
        nn.init.xavier_uniform_(self.lin_in.weight)
        nn.init.xavier_uniform_(self.lin_out.weight)
        self.k = maxk
        self.nonlinear = nonlinear

    def forward(self, adj_sum, x):
        x = self.lin_in(x).relu()
        for i in range(self.num_layers):
            x = self.linlayers[i](x)
            if self.nonlinear == "maxk":
                x = RefMaxK.apply(x, self.k)
            elif self.nonlinear == "relu":
                x = F.relu(x)
            x = self.dropoutlayers[i](x)
            x = (1.0 + self.eps[i]) * x + aggregate(adj_sum, x)
            if self.norm:
                x = self.normlayers[i](x)
        return self.lin_out(x)
