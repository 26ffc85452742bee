"""Peer-memory exchange kernels (csrc/peer.cu, bank.cu PUSH form) on the GPU.

The checks live in tools/peer_check.py and run in a subprocess: a kernel that gives up waiting
for a peer traps, which would poison this process's CUDA context for every later test."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port() -> str:
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return str(s.getsockname()[1])


def _run(cmd, timeout=600, **extra_env):
    env = dict(os.environ, MAXK_PEER_TIMEOUT_MS="20000", **extra_env)
    p = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       timeout=timeout)
    out = p.stdout.decode()
    assert p.returncode == 0, out[-4000:]
    return out


@pytest.mark.parametrize("world", [2, 4])
def test_virtual_ranks_on_one_gpu(built_lib, world):
    """Copy-engine push + the forward that waits per source block, and the reduce-scatter by loads (all
    virtual ranks in ONE launch): WORLD virtual ranks on one device, several rounds through the same
    windows (epoch / ready / done flags, both table buffers re-used)."""
    out = _run([sys.executable, "tools/peer_check.py", "virtual", str(world)])
    assert "virtual peer check: OK" in out


def test_sharded_path_through_peer_windows_world1(built_lib):
    """dist.sharded_forward / sharded_backward with MAXK_PEER_EXCHANGE on an NCCL group of one rank:
    IPC export, window views, index copy-out, `out=` of spgemm_backward -- against the NCCL path."""
    out = _run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "1",
                "--master-addr", "127.0.0.1", "--master-port", _free_port(), "tools/peer_check.py", "dist"])
    assert "dist peer check: OK" in out


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_path_through_peer_windows_two_gpus(built_lib):
    out = _run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                "--master-addr", "127.0.0.1", "--master-port", _free_port(), "tools/peer_check.py", "dist", "--stress"])
    assert "dist peer check: OK" in out


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_multicast_forms_two_gpus(built_lib):
    """The NVLink multicast forms (symmetric-memory windows, mk_peer_push_mc, mk_peer_reduce_scatter_mc) forced
    at two ranks: forward bit-equal to the NCCL form, backward to summation order.  On a box without
    multicast the windows fall back to CUDA IPC and the run checks the unicast forms again."""
    out = _run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                "--master-addr", "127.0.0.1", "--master-port", _free_port(), "tools/peer_check.py", "dist", "--stress"],
               MAXK_PEER_PUSH="mc")
    assert "dist peer check: OK" in out
