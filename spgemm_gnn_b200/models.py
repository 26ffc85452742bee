"""Model families of the reference, running on the B200 hot path.

Two sets, both with the reference constructor
`(in_size, hid_size, num_hid_layers, out_size, maxk=32, feat_drop=0.5, norm=False, nonlinear="maxk")`:

* `SAGE`, `GCN`, `GIN` -- drop-ins of `utils/models.py:109-166, 240-288, 363-411`, the models
  every logged result of the reference was trained with.  There MaxK is `topk -> mask -> mul`
  and the conv is a DGL layer (`SAGEConv(mean)`, `GraphConv(norm='both', weight=None)`,
  `GINConv(None, 'sum', learn_eps=True)`); here MaxK emits CBSR once per layer and the conv's
  `update_all` is the SpGEMM kernel on it, with the SSpMM kernel as its backward.  Parameter
  names follow DGL's (`layers.i.fc_self.weight`, `layers.i.fc_neigh.weight`, `layers.i.bias`,
  `gcnlayers.i.bias`, `gcnlayers.i.eps`) so that state dicts line up.
* `MaxKSAGE`, `MaxKGCN`, `MaxKGIN` -- drop-ins of `utils/integrated_models.py:8-219`, stacking the
  layer classes of `maxk_layers.py` (MaxK applied inside the conv).

`g` is a `CSRGraph`.  With `nonlinear="relu"` the aggregation is the dense cuSPARSE SpMM (the
comparator of the reference's speed-up figures), not the hot path.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn
import torch.nn.functional as F
import torch.nn.init as init
from .maxk_layers import (CBSRToDenseFunction, Linear, linear, MaxKCBSRFunction, MaxKFunction, MaxKGCNConv,
                          MaxKGINConv, MaxKSAGEConv, _dense_aggregate, add_layer_norm, aggregate_cbsr)


def _aggregate_cbsr(g, sp_data, sp_index, kind, dim):
    return aggregate_cbsr(g, sp_data, sp_index, kind, dim)


# The first and last Linear of every model have extents that are not multiples of 4 on the real
# datasets (Reddit: 602 inputs, 41 classes), which sends cuBLAS to its unaligned kernels
# (`..._align1`: 0.66 ms for the 602 -> 256 GEMM on the Reddit shape against ~0.08 ms for a
# 256 -> 256 one, gpurun_out/epoch_profile.txt).  With MAXK_ALIGN_GEMM=1 the node features are
# padded ONCE with zero columns (`pad_features`) and the two weights are padded on the fly inside
# the call -- same parameters, same state dict, same result up to summation order.  Measured on a B200
# (profiles/r2/epoch_ln_align_call15.log): MaxK-SAGE epoch on the Reddit shape 34.67 -> 33.4 ms; on by
# default, MAXK_ALIGN_GEMM=0 switches it off.
_ALIGN_GEMM = os.environ.get("MAXK_ALIGN_GEMM", "1") != "0"
_ALIGN = 8


def set_align_gemm(on: bool) -> None:
    global _ALIGN_GEMM
    _ALIGN_GEMM = bool(on)


def align_gemm() -> bool:
    return _ALIGN_GEMM


def pad_features(x: torch.Tensor, multiple: int = _ALIGN) -> torch.Tensor:
    """Node features with zero columns appended up to a multiple of `multiple` (done once per
    dataset); the models accept them in the place of the unpadded matrix."""
    pad = (-x.shape[1]) % multiple
    return x if pad == 0 else F.pad(x, (0, pad))


def aligned_linear(lin: nn.Linear, x: torch.Tensor) -> torch.Tensor:
    """`lin(x)`.  `x` may carry zero columns beyond `lin.in_features` (pad_features); with
    MAXK_ALIGN_GEMM the output extent is padded to a multiple of 8 inside the GEMM as well and the
    result sliced back."""
    extra_in = x.shape[1] - lin.in_features
    if extra_in < 0:
        raise RuntimeError(f"input has {x.shape[1]} columns, the layer expects {lin.in_features}")
    extra_out = (-lin.out_features) % _ALIGN if _ALIGN_GEMM else 0
    if extra_in == 0 and extra_out == 0:
        return lin(x)
    w = F.pad(lin.weight, (0, extra_in, 0, extra_out))
    b = None if lin.bias is None else F.pad(lin.bias, (0, extra_out))
    y = linear(x, w, b)
    return y if extra_out == 0 else y[:, :lin.out_features]


# ---------------------------------------------------------------------------------------
# utils/models.py family
# ---------------------------------------------------------------------------------------
class _SAGEConvMean(nn.Module):
    """dgl.nn.SAGEConv(in, out, 'mean', feat_drop, norm) for in == out (utils/models.py:140):
    rst = fc_self(h) + fc_neigh(mean_j h_j) + bias, then norm.  `cbsr` carries MaxK(h) when the
    layer input is the MaxK output, so the mean runs on the sparse form."""

    def __init__(self, in_feats, out_feats, feat_drop=0.0, norm=None):
        super().__init__()
        self.feat_drop = nn.Dropout(feat_drop)
        self.fc_self = Linear(in_feats, out_feats, bias=False)
        self.fc_neigh = Linear(in_feats, out_feats, bias=False)
        self.bias = nn.Parameter(torch.zeros(out_feats))
        self.norm = norm
        gain = nn.init.calculate_gain("relu")
        nn.init.xavier_uniform_(self.fc_self.weight, gain=gain)
        nn.init.xavier_uniform_(self.fc_neigh.weight, gain=gain)

    def forward(self, g, feat, cbsr=None, dim=None):
        if cbsr is not None:
            # the conv input is MaxK's output: dropout on its k kept entries per row is dropout
            # on the masked matrix (zeros stay zeros); the dense form is only needed by fc_self
            sp_data, sp_index = cbsr
            sp_data = self.feat_drop(sp_data)
            h = CBSRToDenseFunction.apply(sp_data, sp_index, dim)
            neigh = _aggregate_cbsr(g, sp_data, sp_index, "mean", dim)
        else:
            h = self.feat_drop(feat)
            neigh = _dense_aggregate(g, h, "mean")
        return add_layer_norm(self.fc_self(h), self.fc_neigh(neigh), self.bias, self.norm)


class SAGE(nn.Module):
    def __init__(self, in_size, hid_size, num_hid_layers, out_size, maxk=32, feat_drop=0.5,
                 norm=False, nonlinear="maxk", cache_strategy=None):
        super().__init__()
        self.layers = nn.ModuleList()
        self.num_layers = num_hid_layers
        for _ in range(num_hid_layers):
            nl = nn.LayerNorm(hid_size, elementwise_affine=True) if norm else None
            self.layers.append(_SAGEConvMean(hid_size, hid_size, feat_drop=feat_drop, norm=nl))
        self.lin_in = Linear(in_size, hid_size)
        self.lin_out = Linear(hid_size, out_size)
        init.xavier_uniform_(self.lin_in.weight)
        init.xavier_uniform_(self.lin_out.weight)
        self.k = maxk
        self.nonlinear = nonlinear

    def forward(self, g, x):
        x = aligned_linear(self.lin_in, x)
        for i in range(self.num_layers):
            if self.nonlinear == "maxk":
                cbsr = MaxKCBSRFunction.apply(x, self.k)
                x = self.layers[i](g, None, cbsr, x.shape[1])
            else:
                if self.nonlinear == "relu":
                    x = F.relu(x)
                x = self.layers[i](g, x)
        return aligned_linear(self.lin_out, x)


class _GraphConvBoth(nn.Module):
    """dgl.nn.GraphConv(hid, hid, norm='both', weight=None -> no weight, bias=True)."""

    needs_dense = False  # the masked dense matrix is never formed for GCN

    def __init__(self, feats):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(feats))
        self.feats = feats

    def pieces(self, g, feat, cbsr=None):
        """(a, b, bias) with conv(feat) == a + b + bias, left unsummed so the caller can fuse the
        sum into the LayerNorm that follows."""
        if cbsr is not None:
            out = _aggregate_cbsr(g, cbsr[0], cbsr[1], "both", self.feats)
        else:
            out = _dense_aggregate(g, feat, "both")
        return out, None, self.bias

    def forward(self, g, feat, cbsr=None):
        a, _, bias = self.pieces(g, feat, cbsr)
        return a + bias


class _GINConvSum(nn.Module):
    """dgl GINConv(apply_func=None, 'sum', learn_eps=True): (1 + eps) * h + sum_j h_j."""

    needs_dense = True   # (1 + eps) * h uses the masked dense matrix

    def __init__(self):
        super().__init__()
        self.eps = nn.Parameter(torch.zeros(1))

    def pieces(self, g, feat, cbsr=None):
        if cbsr is not None:
            neigh = _aggregate_cbsr(g, cbsr[0], cbsr[1], "sum", feat.shape[1])
        else:
            neigh = _dense_aggregate(g, feat, "sum")
        return (1 + self.eps) * feat, neigh, None

    def forward(self, g, feat, cbsr=None):
        a, b, _ = self.pieces(g, feat, cbsr)
        return a + b


class _LinMaxKConvStack(nn.Module):
    """Shared body of GCN and GIN (utils/models.py:274-288, 397-411): lin -> MaxK -> dropout ->
    conv -> LayerNorm."""

    def __init__(self, conv_factory, in_size, hid_size, num_hid_layers, out_size, maxk, feat_drop,
                 norm, nonlinear):
        super().__init__()
        self.num_layers = num_hid_layers
        self.norm = norm
        self.dropoutlayers = nn.ModuleList(nn.Dropout(feat_drop) for _ in range(num_hid_layers))
        self.gcnlayers = nn.ModuleList(conv_factory() for _ in range(num_hid_layers))
        self.normlayers = nn.ModuleList(
            nn.LayerNorm(hid_size, elementwise_affine=True) for _ in range(num_hid_layers if norm else 0))
        self.linlayers = nn.ModuleList(Linear(hid_size, hid_size) for _ in range(num_hid_layers))
        for lin in self.linlayers:
            init.xavier_uniform_(lin.weight)
        self.lin_in = Linear(in_size, hid_size)
        self.lin_out = Linear(hid_size, out_size)
        init.xavier_uniform_(self.lin_in.weight)
        init.xavier_uniform_(self.lin_out.weight)
        self.k = maxk
        self.nonlinear = nonlinear

    def forward(self, g, x):
        x = aligned_linear(self.lin_in, x).relu()
        for i in range(self.num_layers):
            x = self.linlayers[i](x)
            cbsr = None
            drop = self.dropoutlayers[i]
            if self.nonlinear == "maxk":
                sp_data, sp_index = MaxKCBSRFunction.apply(x, self.k)
                if self.training and drop.p > 0:
                    sp_data = drop(sp_data)          # dropout on the kept entries == on x*mask
                if self.gcnlayers[i].needs_dense:
                    x = CBSRToDenseFunction.apply(sp_data, sp_index, x.shape[1])
                cbsr = (sp_data, sp_index)
            else:
                if self.nonlinear == "relu":
                    x = F.relu(x)
                x = drop(x)
            a, b, bias = self.gcnlayers[i].pieces(g, x, cbsr)
            x = add_layer_norm(a, b, bias, self.normlayers[i] if self.norm else None)
        return aligned_linear(self.lin_out, x)


class GCN(_LinMaxKConvStack):
    def __init__(self, in_size, hid_size, num_hid_layers, out_size, maxk=32, feat_drop=0.5,
                 norm=False, nonlinear="maxk"):
        super().__init__(lambda: _GraphConvBoth(hid_size), in_size, hid_size, num_hid_layers,
                         out_size, maxk, feat_drop, norm, nonlinear)


class GIN(_LinMaxKConvStack):
    def __init__(self, in_size, hid_size, num_hid_layers, out_size, maxk=32, feat_drop=0.5,
                 norm=False, nonlinear="maxk"):
        super().__init__(_GINConvSum, in_size, hid_size, num_hid_layers, out_size, maxk, feat_drop,
                         norm, nonlinear)


# ---------------------------------------------------------------------------------------
# utils/integrated_models.py family
# ---------------------------------------------------------------------------------------
class MaxKSAGE(nn.Module):
    """utils/integrated_models.py:8-66."""

    def __init__(self, in_size, hid_size, num_hid_layers, out_size, maxk=32, feat_drop=0.5,
                 norm=False, nonlinear="maxk"):
        super().__init__()
        self.num_layers = num_hid_layers
        self.nonlinear = nonlinear
        self.lin_in = Linear(in_size, hid_size)
        self.lin_out = Linear(hid_size, out_size)
        self.layers = nn.ModuleList()
        for _ in range(num_hid_layers):
            nl = nn.LayerNorm(hid_size, elementwise_affine=True) if norm else None
            self.layers.append(MaxKSAGEConv(hid_size, hid_size, aggregator_type="mean",
                                            feat_drop=feat_drop, norm=nl, maxk=maxk))
        if nonlinear == "maxk":
            self.maxk_fn = MaxKFunction.apply
            self.k_values = [maxk] * num_hid_layers
        self.reset_parameters()

    def reset_parameters(self):
        init.xavier_uniform_(self.lin_in.weight)
        init.xavier_uniform_(self.lin_out.weight)

    def forward(self, g, x):
        x = aligned_linear(self.lin_in, x)
        for layer in self.layers:
            if self.nonlinear == "relu":
                x = F.relu(x)
            x = layer(g, x)
        return aligned_linear(self.lin_out, x)


class _IntegratedStack(nn.Module):
    def __init__(self, conv_attr, conv_factory, in_size, hid_size, num_hid_layers, out_size,
                 maxk, feat_drop, norm, nonlinear):
        super().__init__()
        self.num_layers = num_hid_layers
        self.nonlinear = nonlinear
        self.norm = norm
        self.conv_attr = conv_attr
        self.dropoutlayers = nn.ModuleList(nn.Dropout(feat_drop) for _ in range(num_hid_layers))
        setattr(self, conv_attr, nn.ModuleList(conv_factory() for _ in range(num_hid_layers)))
        self.normlayers = nn.ModuleList(
            nn.LayerNorm(hid_size, elementwise_affine=True) for _ in range(num_hid_layers if norm else 0))
        self.linlayers = nn.ModuleList(Linear(hid_size, hid_size) for _ in range(num_hid_layers))
        self.lin_in = Linear(in_size, hid_size)
        self.lin_out = Linear(hid_size, out_size)
        if nonlinear == "maxk":
            self.maxk_fn = MaxKFunction.apply
            self.k_values = [maxk] * num_hid_layers
        self.reset_parameters()

    def reset_parameters(self):
        init.xavier_uniform_(self.lin_in.weight)
        init.xavier_uniform_(self.lin_out.weight)
        for linear in self.linlayers:
            init.xavier_uniform_(linear.weight)

    def forward(self, g, x):
        x = aligned_linear(self.lin_in, x).relu()
        convs = getattr(self, self.conv_attr)
        for i in range(self.num_layers):
            x = self.linlayers[i](x)
            if self.nonlinear == "relu":
                x = F.relu(x)
            x = self.dropoutlayers[i](x)
            x = convs[i](g, x)
            if self.norm:
                x = self.normlayers[i](x)
        return aligned_linear(self.lin_out, x)


class MaxKGCN(_IntegratedStack):
    """utils/integrated_models.py:68-142."""

    def __init__(self, in_size, hid_size, num_hid_layers, out_size, maxk=32, feat_drop=0.5,
                 norm=False, nonlinear="maxk"):
        super().__init__("gcnlayers", lambda: MaxKGCNConv(hid_size, hid_size, maxk=maxk), in_size,
                         hid_size, num_hid_layers, out_size, maxk, feat_drop, norm, nonlinear)


class MaxKGIN(_IntegratedStack):
    """utils/integrated_models.py:144-219."""

    def __init__(self, in_size, hid_size, num_hid_layers, out_size, maxk=32, feat_drop=0.5,
                 norm=False, nonlinear="maxk"):
        super().__init__("ginlayers", lambda: MaxKGINConv(hid_size, hid_size, maxk=maxk), in_size,
                         hid_size, num_hid_layers, out_size, maxk, feat_drop, norm, nonlinear)


MODELS = {"sage": SAGE, "gcn": GCN, "gin": GIN,
          "maxk-sage": MaxKSAGE, "maxk-gcn": MaxKGCN, "maxk-gin": MaxKGIN}
