#!/usr/bin/env python
"""Checks of the peer-memory exchange kernels (csrc/peer.cu, bank.cu PUSH form, peer.py).

    python tools/peer_check.py virtual [WORLD] [--all-k]    one GPU: WORLD virtual ranks, one stream each
    torchrun --nproc-per-node N tools/peer_check.py dist [--bench] [--products] [--stress]    N GPUs
    (either form: --push-mode 2 selects the experimental vector-copy form of the fused bank + push)

`virtual` exercises the kernels and their flag protocol inside one process (every "peer" window
is a local buffer), `dist` runs the sharded forward/backward of dist.py with the peer path and
with NCCL on the same inputs and compares them (forward bit-exact, backward to summation order),
then optionally times both.  Runs in its own process because a kernel that gives up on a peer
traps, which poisons the CUDA context (tests/test_gpu_peer.py calls this file as a subprocess).
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch


def virtual(world: int) -> int:
    import maxk_kernels as mk
    from spgemm_gnn_b200 import peer

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    peer._TIMEOUT_MS = 8000
    r, k, d = 1000, 32, 256           # r*k*1 is a multiple of 16; r not a multiple of 128
    gen = torch.Generator(device=dev).manual_seed(11)
    streams = [torch.cuda.Stream() for _ in range(world)]
    ok = True

    # ---- all-gather by stores, three rounds through the same windows (epochs 1..3)
    rows = world * r
    offs, total = peer.layout([rows * k * 4, rows * k])
    wins = peer.PeerWindow.create_virtual(total, world, dev)
    for rnd in range(3):
        data = [torch.randn(r, k, device=dev, generator=gen) for _ in range(world)]
        index = [torch.randint(0, 256, (r, k), device=dev, generator=gen, dtype=torch.int32).to(torch.uint8)
                 for _ in range(world)]
        torch.cuda.synchronize()
        outs = []
        for q in range(world):
            with torch.cuda.stream(streams[q]):
                outs.append(peer.allgather(wins[q], [data[q], index[q]], offs, grid=6))
        torch.cuda.synchronize()
        want_d, want_i = torch.cat(data), torch.cat(index)
        for q in range(world):
            good = torch.equal(outs[q][0], want_d) and torch.equal(outs[q][1], want_i)
            ep, err = wins[q].epoch()
            good &= (ep == rnd + 1 and err == 0)
            ok &= good
        print(f"allgather round {rnd}: {'OK' if ok else 'FAIL'}")

    # ---- fused bank + push against cbsr_bank + concatenation (--all-k: every banked width)
    wins3 = []
    for kk in ((8, 16, 32, 64) if "--all-k" in sys.argv else (k,)):
        offs3, total3 = peer.layout([rows * kk * 4, rows * kk * 2, rows * kk])
        wk = peer.PeerWindow.create_virtual(total3, world, dev)
        wins3 += wk
        for rnd in range(2):
            xs = [torch.randn(r, d, device=dev, generator=gen) for _ in range(world)]
            cb = [mk.maxk_forward_cbsr(x, kk) for x in xs]
            ref = [mk.cbsr_bank(sd, si, d, with_index=False) for sd, si in cb]
            torch.cuda.synchronize()
            outs = []
            for q in range(world):
                with torch.cuda.stream(streams[q]):
                    outs.append(peer.bank_push(wk[q], cb[q][0], cb[q][1], d, offs3))
            torch.cuda.synchronize()
            want = (torch.cat([a for a, _, _ in ref]), torch.cat([c for _, _, c in ref]),
                    torch.cat([si for _, si in cb]))
            for q in range(world):
                good = all(torch.equal(a, b) for a, b in zip(outs[q], want))
                ok &= good
            print(f"bank_push k={kk} round {rnd}: {'OK' if ok else 'FAIL'}")

    # ---- reduce-scatter by loads: fixed rank order, so bit-equal to the same fold in torch
    offs1, total1 = peer.layout([rows * k * 4])
    wins1 = peer.PeerWindow.create_virtual(total1, world, dev)
    for rnd in range(3):
        parts = []
        for q in range(world):
            v = wins1[q].view(offs1[0], (rows, k), torch.float32)
            v.copy_(torch.randn(rows, k, device=dev, generator=gen))
            parts.append(v.clone())
        torch.cuda.synchronize()
        outs = []
        for q in range(world):
            with torch.cuda.stream(streams[q]):
                outs.append(peer.reduce_scatter(wins1[q], offs1[0], r, k, grid=5))
        torch.cuda.synchronize()
        for q in range(world):
            acc = parts[0][q * r:(q + 1) * r].clone()
            for p in parts[1:]:
                acc += p[q * r:(q + 1) * r]
            ok &= torch.equal(outs[q], acc)
        print(f"reduce_scatter round {rnd}: {'OK' if ok else 'FAIL'}")
    for w in wins + wins3 + wins1:
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

        res, times = {}, {}
        for mode in (False, True, True):   # NCCL, peer, peer again (window re-use: epochs 2..)
            peer.set_enabled(mode)
            res[mode] = step()
            if bench:
                for _ in range(3):
                    step()
                torch.cuda.synchronize()
                dist.barrier()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(20):
                    step()
                b.record()
                torch.cuda.synchronize()
                t = torch.tensor([a.elapsed_time(b) / 20], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                times[mode] = float(t.item())
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
    dist.barrier()
    peer.close_all()
    dist.destroy_process_group()
    if rank == 0:
        print("dist peer check:", "OK" if ok else "FAIL")
    return 0 if ok else 1


if __name__ == "__main__":
    if "--push-mode" in sys.argv:   # 2 = experimental "own table first, then block copies" bank_push
        from spgemm_gnn_b200 import peer as _peer
        _peer._PUSH_MODE = int(sys.argv[sys.argv.index("--push-mode") + 1])
    mode = sys.argv[1] if len(sys.argv) > 1 else "virtual"
    if mode == "virtual":
        sys.exit(virtual(int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 4))
    sys.exit(distributed("--bench" in sys.argv))
