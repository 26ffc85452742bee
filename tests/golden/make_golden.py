#!/usr/bin/env python
"""Generates tests/golden/*.npz by running the REFERENCE'S OWN Python for the parts of the
hot path that are importable in the build container:

  * `utils/models.py::MaxK`               (forward mask semantics, backward grad*mask)
  * `utils/maxk_layers.py::MaxKFunction`  (fallback branch, the only one that ever ran)
  * `utils/maxk_layers.py::MaxKSAGEConv._extract_sparse_format` (CBSR layout + padding)

Both modules `import dgl` at the top; DGL is not installed, so an empty stand-in module
is put into `sys.modules` first.  Nothing of DGL is executed by the three items above.
The reference's compiled `maxk_kernels` (cp39, sm_80) cannot be imported, which is what
selects the fallback branch, exactly like the reference's own logged runs
(run/reddit.log:26 "Custom kernels: Not available").

Run once in the build container (needs /root/reference):  python tests/golden/make_golden.py
The GPU box has no /root/reference; tests only read the committed .npz files.
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


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
    dgl_nn.pytorch = dgl_nn_pt
    dgl_nn_pt.conv = dgl_nn_conv
    dgl.nn = dgl_nn
    dgl.function = dgl_fn
    sys.modules.update({
        "dgl": dgl, "dgl.nn": dgl_nn, "dgl.function": dgl_fn,
        "dgl.nn.pytorch": dgl_nn_pt, "dgl.nn.pytorch.conv": dgl_nn_conv,
    })
    # make sure the product's shim of the same name can never be picked up here
    sys.modules["maxk_kernels"] = None


def main():
    if not os.path.isdir(REF):
        sys.exit("needs /root/reference (build container only)")
    _stub_dgl()
    sys.path = [p for p in sys.path if os.path.abspath(p or ".") != os.path.abspath(os.path.join(HERE, "..", ".."))]
    sys.path.insert(0, REF)
    from utils import models as ref_models          # noqa: E402
    from utils import maxk_layers as ref_layers      # noqa: E402

    assert ref_layers.KERNELS_AVAILABLE is False
    out = {}
    cases = [(48, 64, 8), (48, 64, 16), (40, 256, 32), (24, 256, 64), (16, 384, 16), (33, 32, 32), (20, 96, 1)]
    g = torch.Generator().manual_seed(97)       # utils/config.py:54 default seed
    for ci, (n, d, k) in enumerate(cases):
        x = torch.randn(n, d, generator=g)        # tie-free almost surely (asserted below)
        assert all(torch.unique(r).numel() == d for r in x)
        gy = torch.randn(n, d, generator=g)
        # utils/models.py::MaxK
        xi = x.clone().requires_grad_(True)
        y = ref_models.MaxK.apply(xi, k)
        y.backward(gy)
        # utils/maxk_layers.py::MaxKFunction (fallback branch)
        xj = x.clone().requires_grad_(True)
        y2 = ref_layers.MaxKFunction.apply(xj, k)
        y2.backward(gy)
        assert torch.equal(y, y2) and torch.equal(xi.grad, xj.grad)
        pre = f"c{ci}_"
        out[pre + "x"] = x.numpy()
        out[pre + "k"] = np.int64(k)
        out[pre + "grad_out"] = gy.numpy()
        out[pre + "maxk_out"] = y.detach().numpy()
        out[pre + "maxk_grad_in"] = xi.grad.numpy()
        if d <= 256:
            # _extract_sparse_format casts columns to uint8: only meaningful for D <= 256
            conv = ref_layers.MaxKSAGEConv(d, d, maxk=k)
            sp_data, sp_index = conv._extract_sparse_format(y.detach())
            out[pre + "sp_data"] = sp_data.numpy()
            out[pre + "sp_index"] = sp_index.numpy()
    # rows with fewer than k non-zeros: the padding convention (0.0, idx 0)
    x = torch.zeros(6, 64)
    x[0, [3, 9]] = torch.tensor([1.5, -2.0])
    x[1, 0] = 4.0
    x[3, torch.arange(0, 64, 4)] = torch.arange(1, 17, dtype=torch.float32)
    conv = ref_layers.MaxKSAGEConv(64, 64, maxk=8)
    sp_data, sp_index = conv._extract_sparse_format(x)
    out["pad_x"] = x.numpy()
    out["pad_k"] = np.int64(8)
    out["pad_sp_data"] = sp_data.numpy()
    out["pad_sp_index"] = sp_index.numpy()
    out["num_cases"] = np.int64(len(cases))
    path = os.path.join(HERE, "maxk_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", "torch", torch.__version__)


if __name__ == "__main__":
    main()
