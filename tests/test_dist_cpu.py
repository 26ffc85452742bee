"""N > 1 host logic on CPU: two gloo ranks, 1-D row partition, CBSR all-gather forward and
CBSR-gradient reduce(-scatter) backward.  The per-rank SpGEMM / SSpMM are computed by the oracle
here (no GPU in this tier); the test pins the sharding arithmetic: padded equal-row blocks,
global column ids, gathered-table indexing, and that the folded gradient equals the single-
process result."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import c_oracle
        from spgemm_gnn_b200 import dist as mdist
        from spgemm_gnn_b200.graph import synthetic_graph

        n, d, k = 1001, 64, 16           # n not divisible by world: padding path
        g = synthetic_graph(n, n * 14, seed=5)
        rng = np.random.default_rng(5)
        x = rng.standard_normal((n, d)).astype(np.float32)
        dy = rng.standard_normal((n, d)).astype(np.float32)
        val_full = g.edge_weights("mean")
        local, r0, r1 = mdist.shard_graph(g, rank, world)
        val = mdist.shard_edge_weights(g, local, r0, r1, "mean")
        r = mdist.rows_per_rank(n, world)
        assert local.num_nodes() == r and local.num_src == world * r
        # local MaxK (oracle stands in for the kernel), padded to r rows
        xl = np.zeros((r, d), np.float32)
        xl[: r1 - r0] = x[r0:r1]
        sd, si = c_oracle.maxk_cbsr(xl, k)
        fd, fi = mdist.allgather_cbsr(torch.from_numpy(sd), torch.from_numpy(si))
        assert fd.shape == (world * r, k) and fi.dtype == torch.uint8
        wd, wi = c_oracle.maxk_cbsr(x, k)
        assert np.array_equal(fd.numpy()[:n], wd) and np.array_equal(fi.numpy()[:n], wi)
        # forward on the shard against the gathered table == rows of the single-process result
        out = c_oracle.spgemm_fwd(local.indptr.numpy(), local.indices.numpy(), val.numpy(),
                                  fd.numpy(), fi.numpy(), d)
        want = c_oracle.spgemm_fwd(g.indptr.numpy(), g.indices.numpy(), val_full.numpy(), wd, wi, d)
        np.testing.assert_allclose(out[: r1 - r0], want[r0:r1], rtol=1e-12, atol=1e-12)
        assert not out[r1 - r0:].any()                     # padded rows stay empty
        # backward: per-rank contributions to every node, folded by reduce-scatter
        dyl = np.zeros((r, d), np.float32)
        dyl[: r1 - r0] = dy[r0:r1]
        part = c_oracle.sspmm_bwd(local.indptr.numpy(), local.indices.numpy(), val.numpy(), dyl,
                                  fi.numpy())
        mine = mdist.reduce_scatter_rows(torch.from_numpy(part))
        want_b = c_oracle.sspmm_bwd(g.indptr.numpy(), g.indices.numpy(), val_full.numpy(), dy, wi)
        np.testing.assert_allclose(mine.numpy()[: r1 - r0], want_b[r0:r1], rtol=1e-10, atol=1e-12)
        # replicated weights: gradient all-reduce in one bucket
        p1, p2 = torch.nn.Parameter(torch.zeros(3, 2)), torch.nn.Parameter(torch.zeros(5))
        p1.grad, p2.grad = torch.full((3, 2), float(rank + 1)), torch.arange(5.0) * (rank + 1)
        mdist.allreduce_grads([p1, p2])
        tot = sum(range(1, world + 1))
        assert torch.equal(p1.grad, torch.full((3, 2), float(tot)))
        assert torch.equal(p2.grad, torch.arange(5.0) * tot)
        ret[rank] = "ok"
    except Exception as e:  # surfaced by the parent
        import traceback
        ret[rank] = traceback.format_exc()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", [2, 3])
def test_row_partition_gloo(world):
    """world 2: 1001 rows -> 501 + 500 (+1 padded); world 3: 334 + 334 + 333 (+1 padded)."""
    port = 29500 + (os.getpid() % 2000) + world
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {r: "ok" for r in range(world)}, dict(ret)


def test_random_relabel_keeps_the_graph_and_balances_shards():
    sys.path.insert(0, ROOT)
    from spgemm_gnn_b200 import dist as mdist
    from spgemm_gnn_b200.graph import from_edges
    # a graph whose heavy rows are all at the front: equal-row shards are unbalanced
    n = 4000
    deg = np.where(np.arange(n) < 400, 60, 2)
    rows = np.repeat(np.arange(n), deg)
    rng = np.random.default_rng(0)
    cols = rng.integers(0, n, rows.size)
    g = from_edges(torch.from_numpy(rows), torch.from_numpy(cols), n)
    g2, perm = mdist.random_relabel(g, seed=3)
    assert g2.num_edges() == g.num_edges()
    a = set(zip(perm[g.row_ids()].tolist(), perm[g.indices.long()].tolist()))
    b = set(zip(g2.row_ids().tolist(), g2.indices.tolist()))
    assert a == b
    def imbalance(gr):
        r = mdist.rows_per_rank(n, 8)
        nnz = [int(gr.indptr[min((p + 1) * r, n)] - gr.indptr[min(p * r, n)]) for p in range(8)]
        return max(nnz) / (sum(nnz) / 8)
    assert imbalance(g) > 3.0 and imbalance(g2) < 1.3


def _worker_nnz(rank, world, port, ret, tmp):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import c_oracle
        from spgemm_gnn_b200 import dist as mdist
        from spgemm_gnn_b200.graph import from_edges

        # power-law graph with the heavy rows at the FRONT: equal-row shards would be badly unbalanced
        n, d, k = 3000, 64, 16
        rng = np.random.default_rng(9)
        deg = np.maximum((600.0 / (1.0 + np.arange(n)) ** 0.7).astype(np.int64), 2)
        rows = np.repeat(np.arange(n), deg)
        cols = rng.integers(0, n, rows.size)
        g = from_edges(torch.from_numpy(rows), torch.from_numpy(cols), n)
        sg = mdist.ShardedGraph(g, rank, world, balance="nnz")
        # nnz imbalance below 2 % WITHOUT relabelling; the row counts differ a lot
        nnz = torch.tensor([float(sg.num_edges())])
        alln = [torch.zeros(1) for _ in range(world)]
        dist.all_gather(alln, nnz)
        alln = torch.cat(alln)
        assert float(alln.max() / alln.mean()) < 1.02, alln
        b = sg.bounds
        assert b[0] == 0 and b[-1] == n and (b[1] - b[0]) * 3 < (b[-1] - b[-2])
        r = sg.rows_per_rank
        assert r % 16 == 0 and r >= max(b[p + 1] - b[p] for p in range(world)) and sg.num_src == world * r
        # column ids are table rows: block q starts at q * r
        x = rng.standard_normal((n, d)).astype(np.float32)
        dy = rng.standard_normal((n, d)).astype(np.float32)
        xl = sg.local_rows(torch.from_numpy(x)).numpy()
        sd, si = c_oracle.maxk_cbsr(xl, k)
        fd, fi = mdist.allgather_cbsr(torch.from_numpy(sd), torch.from_numpy(si))
        wd, wi = c_oracle.maxk_cbsr(x, k)
        for q in range(world):                               # gathered table == global table, block by block
            assert np.array_equal(fd.numpy()[q * r: q * r + b[q + 1] - b[q]], wd[b[q]:b[q + 1]])
        val = sg.edge_weights("mean")
        out = c_oracle.spgemm_fwd(sg.indptr.numpy(), sg.indices.numpy(), val.numpy(), fd.numpy(), fi.numpy(), d)
        want = c_oracle.spgemm_fwd(g.indptr.numpy(), g.indices.numpy(), g.edge_weights("mean").numpy(), wd, wi, d)
        nloc = sg.row_end - sg.row_begin
        np.testing.assert_allclose(out[:nloc], want[sg.row_begin:sg.row_end], rtol=1e-12, atol=1e-12)
        assert not out[nloc:].any()
        dyl = sg.local_rows(torch.from_numpy(dy)).numpy()
        part = c_oracle.sspmm_bwd(sg.indptr.numpy(), sg.indices.numpy(), val.numpy(), dyl, fi.numpy())
        mine = mdist.reduce_scatter_rows(torch.from_numpy(part))
        want_b = c_oracle.sspmm_bwd(g.indptr.numpy(), g.indices.numpy(), g.edge_weights("mean").numpy(), dy, wi)
        np.testing.assert_allclose(mine.numpy()[:nloc], want_b[sg.row_begin:sg.row_end], rtol=1e-10, atol=1e-12)
        # shard file round trip: a rank can start from its file alone
        path = os.path.join(tmp, f"shard{rank}")
        mdist.save_shard(sg, path)
        sg2 = mdist.load_shard(path)
        assert torch.equal(sg2.indptr, sg.indptr) and torch.equal(sg2.indices, sg.indices)
        assert sg2.bounds == sg.bounds and sg2.rows_per_rank == r and sg2.rank == rank and sg2.world == world
        assert all(torch.equal(sg2.edge_weights(kd), sg.edge_weights(kd)) for kd in ("mean", "both", "sum"))
        with pytest.raises(RuntimeError):
            sg.row_slice(0, 1)
        assert sg.to("cpu") is sg
        ret[rank] = "ok"
    except Exception:
        import traceback
        ret[rank] = traceback.format_exc()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_nnz_balanced_shards_gloo_world3(tmp_path):
    """f-4: nnz-balanced contiguous row ranges (no relabel), remapped table-row column ids, forward /
    backward against the single-process oracle, shard files."""
    world = 3
    port = 29500 + (os.getpid() % 2000) + 17
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_nnz, args=(world, port, ret, str(tmp_path)), nprocs=world, join=True)
    assert dict(ret) == {r: "ok" for r in range(world)}, dict(ret)
