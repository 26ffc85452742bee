#!/usr/bin/env python
"""How long do the copy-engine pushes of the overlapped all-gather take?  torchrun, 2 ranks.
Emulates the push of an N-rank run towards ONE real peer: (N-1) x (segments + flag) back to back on a
side stream, for the segment sizes of the Reddit shape at N ranks; prints the time per peer and the
time of a single isolated copy of each size, with 1 and 2 side streams."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from spgemm_gnn_b200 import _lib, peer

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
L = _lib.lib()
k = 32
for nodes, label in ((232965, "reddit"), (2449029, "products")):
    for nranks in (2, 8):
        r = -(-nodes // nranks)
        per_rank = [r * k * 4, r * k * 2, r * k]
        offs, total = peer.layout([nranks * b for b in per_rank])
        win = peer.PeerWindow.create(total, None)
        assert win is not None
        torch.cuda.synchronize(); dist.barrier()
        n = len(per_rank)
        o = (ctypes.c_int64 * n)(*offs); b = (ctypes.c_int64 * n)(*per_rank)
        streams = [torch.cuda.Stream() for _ in range(4)]
        for nstreams in (1, 2, 4):
            ts = []
            for it in range(4):
                torch.cuda.synchronize(); dist.barrier()
                t0 = torch.cuda.Event(enable_timing=True); t0.record()
                ends = []
                for si in range(nstreams):
                    streams[si].wait_event(t0)
                for p in range(nranks - 1):      # every "peer" is the one real peer
                    st = streams[p % nstreams]
                    _lib.check(L.mk_peer_push(win.ptrs, 2, rank, n, o, b, st.cuda_stream), "push")
                for si in range(nstreams):
                    e = torch.cuda.Event(enable_timing=True); e.record(streams[si]); ends.append(e)
                torch.cuda.synchronize()
                ts.append(max(t0.elapsed_time(e) for e in ends))
            if rank == 0:
                mb = sum(per_rank) / 1e6
                print(f"{label} N={nranks}: {nranks-1} peers x {mb:.1f} MB ({n} copies + flag each), {nstreams} stream(s): "
                      f"{min(ts[1:])*1e3:.0f} us total = {min(ts[1:])*1e3/(nranks-1):.1f} us per peer "
                      f"({(nranks-1)*mb/1e3/(min(ts[1:])*1e-3):.0f} GB/s)", flush=True)
        torch.cuda.synchronize(); dist.barrier()
        win.close()
dist.barrier()
dist.destroy_process_group()
