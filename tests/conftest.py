import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "maxk_reference.npz"))


@pytest.fixture(scope="session")
def golden_layers():
    """Outputs and kernel-call arguments of the reference's own MaxKSAGEConv / MaxKGCNConv code
    (tests/golden/make_golden_layers.py)."""
    return np.load(os.path.join(ROOT, "tests", "golden", "layers_reference.npz"))


def layer_cases(gl):
    """(name, call-argument dict, expected layer output, the part added to the aggregation)."""
    for name in [str(s) for s in gl["names"]]:
        call = {k: gl[f"{name}_call_{k}"] for k in ("ptr", "idx", "val", "sp_data", "sp_index")}
        n, e, k, d = (int(v) for v in gl[f"{name}_call_dims"])
        call.update(n=n, e=e, k=k, d=d)
        add = gl[f"{name}_h_self"] if name.startswith("sage") else gl[f"{name}_bias"][None, :]
        yield name, call, gl[f"{name}_y"], add


@pytest.fixture(scope="session")
def built_lib():
    """The product library, built in-tree if stale (nvcc cross-compiles without a GPU)."""
    from spgemm_gnn_b200 import build as _b
    return _b.build()


def small_graph(n, avg_deg, seed=97, device="cpu"):
    from spgemm_gnn_b200.graph import synthetic_graph
    return synthetic_graph(n, int(n * avg_deg), seed=seed, device=device)
