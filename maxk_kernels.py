"""Top-level alias so that the reference's `import maxk_kernels` (utils/maxk_layers.py:10,
maxk_gnn_integrated.py:24-31) resolves to the B200 library when the repo root is on
sys.path.

Two bindings of the same C ABI (include/maxk_b200.h) stand behind the name: the ctypes shim
(spgemm_gnn_b200/maxk_kernels.py, the default; it also carries the additions the layers use) and the
compiled pybind11 / ATen extension the reference's own build produces (spgemm_gnn_b200/binding/,
`MAXK_BINDING=pybind`): its five entry points then replace the shim's."""
import os as _os

from spgemm_gnn_b200.maxk_kernels import *  # noqa: F401,F403
from spgemm_gnn_b200.maxk_kernels import __all__  # noqa: F401

if _os.environ.get("MAXK_BINDING", "ctypes") == "pybind":
    from spgemm_gnn_b200.maxk_kernels_ext import (  # noqa: F401
        maxk_backward, maxk_forward, maxk_forward_cbsr, spgemm_backward, spgemm_forward)
