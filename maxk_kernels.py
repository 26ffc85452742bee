"""Top-level alias so that the reference's `import maxk_kernels` (utils/maxk_layers.py:10,
maxk_gnn_integrated.py:24-31) resolves to the B200 library when the repo root is on
sys.path."""
from spgemm_gnn_b200.maxk_kernels import *  # noqa: F401,F403
from spgemm_gnn_b200.maxk_kernels import __all__  # noqa: F401
