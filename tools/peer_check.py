#!/usr/bin/env python
"""Checks of the peer-memory exchange kernels (csrc/peer.cu, bank.cu PUSH form, peer.py).

    python tools/peer_check.py virtual [WORLD] [--all-k]    one GPU: WORLD virtual ranks in one process
    torchrun --nproc-per-node N tools/peer_check.py dist [--bench] [--products] [--stress]    N GPUs

`virtual` exercises the kernels and their flag protocol inside one process (every "peer" window
is a local buffer), `dist` runs the sharded forward/backward of dist.py with the peer path and
with NCCL on the same inputs and compares them (forward bit-exact, backward to summation order),
then optionally times both.  Runs in its own process (tests/test_gpu_peer.py calls this file as a
subprocess): a peer time-out leaves garbage behind that no later test should inherit.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch


def virtual(world: int) -> int:
    """One GPU, WORLD virtual ranks in one process (every "peer" window is a local buffer).  The
    overlapped all-gather needs no kernel that waits for another launch -- the copy engines carry rows
    and flags -- so the ranks simply take turns on one stream; the reduce-scatter, whose blocks do wait
    for peers' flags, runs all ranks in ONE launch (mk_peer_reduce_scatter_virtual)."""
    import maxk_kernels as mk
    from spgemm_gnn_b200 import dist as mdist, peer
    from spgemm_gnn_b200.graph import synthetic_graph

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    peer._TIMEOUT_MS = 8000
    d = 256
    gen = torch.Generator(device=dev).manual_seed(11)
    ok = True
    g = synthetic_graph(4 * 1003, 4 * 1003 * 90, seed=5, device=dev)
    n = g.num_nodes()
    wins_all = []

    for kk in ((8, 16, 32, 64) if "--all-k" in sys.argv else (32,)):
        r = mdist.rows_per_rank(n, world)
        rows = world * r
        per_rank = [r * kk * 4, r * kk * 2, r * kk]
        offs, total = peer.layout([world * b for b in per_rank] * 2)
        wins = peer.PeerWindow.create_virtual(total, world, dev)
        wins_all += wins
        shards = [mdist.shard_graph(g, q, world) for q in range(world)]
        vals = [mdist.shard_edge_weights(g, sh[0], sh[1], sh[2], "mean") for sh in shards]
        for rnd in range(4):   # rounds 2, 3 re-use the two table buffers: release / begin_push handshake
            mode = ("dma", "sm", "sm", "dma")[rnd]     # who moves the rows: copy engines / NVLink-store CTAs
            xs = [torch.randn(r, d, device=dev, generator=gen) for _ in range(world)]
            cb = [mk.maxk_forward_cbsr(x, kk) for x in xs]
            ref = [mk.cbsr_bank(sd, si, d, with_index=False) for sd, si in cb]
            want = (torch.cat([a for a, _, _ in ref]), torch.cat([c for _, _, c in ref]),
                    torch.cat([si for _, si in cb]))
            outs = []
            for q in range(world):
                w = wins[q]
                buf = w.next_buffer()
                o = offs[3 * buf: 3 * buf + 3]
                mine = slice(q * r, (q + 1) * r)
                peer.begin_push(w, buf)
                fd = w.view(o[0], (rows, kk), torch.float32)
                fs = w.view(o[1], (rows, kk), torch.int16)
                fi = w.view(o[2], (rows, kk), torch.uint8)
                fd.zero_(); fs.zero_(); fi.zero_()          # nothing of the previous round survives
            for q in range(world):
                w = wins[q]
                buf = w._buf
                o = offs[3 * buf: 3 * buf + 3]
                mine = slice(q * r, (q + 1) * r)
                fd = w.view(o[0], (rows, kk), torch.float32)
                fs = w.view(o[1], (rows, kk), torch.int16)
                fi = w.view(o[2], (rows, kk), torch.uint8)
                fi[mine].copy_(cb[q][1])
                mk.cbsr_bank(cb[q][0], cb[q][1], d, with_index=False, out=(fd[mine], fs[mine]))
                if mode == "dma":
                    peer.publish_and_push(w, buf, o, per_rank)
                else:                                       # the pushers wait for nobody: ranks can take turns
                    peer.publish(w, buf)
                    peer.push_sm(w, o, per_rank)
            for q in range(world):
                w = wins[q]
                buf = w._buf
                o = offs[3 * buf: 3 * buf + 3]
                fd = w.view(o[0], (rows, kk), torch.float32)
                fs = w.view(o[1], (rows, kk), torch.int16)
                fi = w.view(o[2], (rows, kk), torch.uint8)
                local = shards[q][0]
                split = mk.block_split(local.indptr, local.indices, r, world, q, r)
                # round 2: the forward kernel carries pusher CTAs as well (they re-send rows that are
                # already there -- on one device the real overlap cannot be staged, the code path can)
                x = peer.exchange(w, r, o, per_rank) if rnd == 2 else peer.exchange(w, r)
                if rnd % 2 == 1:   # rounds 1, 3: the forward in source-block phases (one launch each)
                    blk = mk.block_pointers(local.indptr, local.indices, r, world, r)
                    out = mk.spgemm_forward_banked(local.indptr, local.indices, vals[q], fd, fs, r, local.num_edges(),
                                                   kk, d, phases=mk.forward_phases(world, q), blk=blk,
                                                   n_blocks=world, wait=x)
                else:
                    out = mk.spgemm_forward_banked(local.indptr, local.indices, vals[q], fd, fs, r, local.num_edges(),
                                                   kk, d, split=split, wait=x)
                peer.join_push(w)
                outs.append((out, fd.clone(), fs.clone(), fi.clone()))
                peer.release(w)
            torch.cuda.synchronize()
            for q in range(world):
                out, fd, fs, fi = outs[q]
                local = shards[q][0]
                split = mk.block_split(local.indptr, local.indices, r, world, q, r)
                if rnd % 2 == 1:
                    blk = mk.block_pointers(local.indptr, local.indices, r, world, r)
                    out_ref = mk.spgemm_forward_banked(local.indptr, local.indices, vals[q], want[0], want[1], r,
                                                       local.num_edges(), kk, d, phases=mk.forward_phases(world, q),
                                                       blk=blk, n_blocks=world)
                else:
                    out_ref = mk.spgemm_forward_banked(local.indptr, local.indices, vals[q], want[0], want[1], r,
                                                       local.num_edges(), kk, d, split=split)
                good = (torch.equal(fd, want[0]) and torch.equal(fs, want[1]) and torch.equal(fi, want[2])
                        and torch.equal(out, out_ref))
                ep, err = wins[q].epoch()
                good &= (ep == rnd + 1 and err == 0)
                ok &= bool(good)
            print(f"push ({mode}) + waiting forward k={kk} round {rnd}: {'OK' if ok else 'FAIL'}")

    # ---- reduce-scatter by loads: fixed rank order, so bit-equal to the same fold in torch
    k, r = 32, 1000
    rows = world * r
    offs1, total1 = peer.layout([rows * k * 4])
    wins1 = peer.PeerWindow.create_virtual(total1, world, dev)
    for rnd in range(3):
        parts = []
        for q in range(world):
            v = wins1[q].view(offs1[0], (rows, k), torch.float32)
            v.copy_(torch.randn(rows, k, device=dev, generator=gen))
            parts.append(v.clone())
        torch.cuda.synchronize()
        outs = peer.reduce_scatter_virtual(wins1, offs1[0], r, k, grid=5)
        torch.cuda.synchronize()
        for q in range(world):
            acc = parts[0][q * r:(q + 1) * r].clone()
            for p in parts[1:]:
                acc += p[q * r:(q + 1) * r]
            ok &= torch.equal(outs[q], acc)
            ep, err = wins1[q].epoch()
            ok &= (ep == rnd + 1 and err == 0)
        print(f"reduce_scatter round {rnd}: {'OK' if ok else 'FAIL'}")
    for w in wins_all + wins1:
        w.close()
    print("virtual peer check:", "OK" if ok else "FAIL")
    return 0 if ok else 1


def distributed(bench: bool) -> int:
    import torch.distributed as dist

    import maxk_kernels as mk
    from spgemm_gnn_b200 import dist as mdist, peer
    from spgemm_gnn_b200.graph import shaped_graph, synthetic_graph

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    cases = [("small k32", synthetic_graph(20001, 20001 * 150, seed=97, device=dev), 32, 256),
             ("small k16 plain", synthetic_graph(9001, 9001 * 20, seed=3, device=dev), 16, 128)]
    if bench:
        cases.append(("reddit k32", "reddit", 32, 256))
    if "--products" in sys.argv:
        cases.append(("ogbn-products k32", "ogbn-products", 32, 256))
    for name, g, k, d in cases:
        if isinstance(g, str):
            g = shaped_graph(g, device=dev)
        local, r0, r1 = mdist.shard_graph(g, rank, world)
        val = mdist.shard_edge_weights(g, local, r0, r1, "mean")
        n_rows = local.num_nodes()
        gen = torch.Generator(device=dev).manual_seed(97 + rank)
        x = torch.randn(n_rows, d, device=dev, generator=gen)
        dy = torch.randn(n_rows, d, device=dev, generator=gen)
        sd, si = mk.maxk_forward_cbsr(x, k)
        ptr, idx = local.indptr, local.indices

        def step():
            out, fi = mdist.sharded_forward(sd, si, ptr, idx, val, n_rows, d)
            return out, fi, mdist.sharded_backward(dy, fi, ptr, idx, val, n_rows, d)

        def timed(fn, reps=20):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record()
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b) / reps], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        res, times = {}, {}
        for mode in (False, True, True):   # NCCL, peer, peer again (window re-use: epochs 2..)
            peer.set_enabled(mode)
            res[mode] = step()
            if bench:
                times[mode] = timed(step)
        if bench and "--sweep" in sys.argv:
            # who moves the rows / how many pusher CTAs, forward and backward timed on their own
            fi0 = res[True][1]
            lim = peer.set_max_mb(0)          # the peer forms at every size
            for push, pushers, ph in (("nccl", 0, 0), ("mc", 0, 0), ("sm_seq", 592, 0), ("sm", 592, 0), ("dma", 0, 0),
                                      ("auto", 0, 0), ("nccl", 0, 0)):
                peer.set_enabled(push != "nccl")
                peer._PHASES = bool(ph)
                if push != "nccl":
                    peer._PUSH, peer._PUSHERS = push, pushers or peer._PUSHERS
                tf = timed(lambda: mdist.sharded_forward(sd, si, ptr, idx, val, n_rows, d))
                tb = timed(lambda: mdist.sharded_backward(dy, fi0, ptr, idx, val, n_rows, d))
                ts = timed(step)
                if rank == 0:
                    print(f"  {name} [{push} {pushers} phases={ph}]: fwd {tf:.3f}  bwd {tb:.3f}  fwd+bwd {ts:.3f} ms", flush=True)
            peer._PUSH, peer._PUSHERS, peer._PHASES = "auto", 592, False
            peer.set_max_mb(lim)
            peer.set_enabled(True)
        o0, f0, b0 = res[False]
        o1, f1, b1 = res[True]
        fwd_equal = torch.equal(o0, o1) and torch.equal(f0, f1)
        if "--stress" in sys.argv:   # back-to-back collectives through the same windows: the forward
            for _ in range(200):    # is deterministic, so every repetition must reproduce it bit for bit
                o, f, _b = step()
                fwd_equal &= torch.equal(o, o1) and torch.equal(f, f1)
        scale = b0.abs().max().clamp_min(1e-20)
        bwd_err = float(((b0 - b1).abs().max() / scale).item())
        good = torch.tensor([1 if (fwd_equal and bwd_err < 1e-5) else 0], device=dev)
        dist.all_reduce(good, op=dist.ReduceOp.MIN)
        ok &= bool(good.item())
        if rank == 0:
            msg = f"{name}: forward bit-equal {fwd_equal}, backward max rel diff {bwd_err:.2e}"
            if bench:
                msg += f" | ms/layer NCCL {times[False]:.3f}  peer {times[True]:.3f}"
            print(msg, "OK" if good.item() else "FAIL", flush=True)
        del g, local, val, x, dy, sd, si, res
        torch.cuda.empty_cache()
    peer.set_enabled(False)
    torch.cuda.synchronize()
    try:
        peer.check_errors()
    except peer.PeerTimeoutError as exc:
        ok = False
        print(f"rank {rank}: {exc}", flush=True)
    dist.barrier()
    peer.close_all()
    dist.destroy_process_group()
    if rank == 0:
        print("dist peer check:", "OK" if ok else "FAIL")
    return 0 if ok else 1


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "virtual"
    if mode == "virtual":
        sys.exit(virtual(int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 4))
    sys.exit(distributed("--bench" in sys.argv))
