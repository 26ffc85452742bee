"""spgemm_gnn_b200 -- B200-native (sm_100a) MaxK-GNN aggregation hot path.

Host side of the drop-in boundary: `maxk_kernels` (the reference's extension entry points
over the C ABI of libmaxk_b200.so), `maxk_layers` (the reference's autograd Functions and
conv layers), `graph` (CSR container + synthetic shapes), `dist` (1-D row partition over
torch.distributed).  Importing the package never needs a GPU; calling a kernel does.
"""
from . import graph  # noqa: F401

__version__ = "0.1.0"
