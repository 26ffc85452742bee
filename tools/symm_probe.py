#!/usr/bin/env python
"""torchrun probe: does this box give CUDA multicast (NVLS) memory through torch's symmetric-memory
allocator?  Prints the multicast pointer, the peers' buffer pointers, checks peer visibility."""
import os
import sys

import torch
import torch.distributed as dist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
try:
    import torch.distributed._symmetric_memory as symm
    from torch._C._distributed_c10d import _SymmetricMemory as S
    try:
        print(f"rank {rank}: has_multicast_support =", S.has_multicast_support(torch._C._autograd.DeviceType.CUDA, dev.index), flush=True)
    except Exception as exc:  # noqa: BLE001
        print(f"rank {rank}: has_multicast_support query failed: {exc}", flush=True)
except Exception as exc:  # noqa: BLE001
    print(f"rank {rank}: symmetric memory import failed: {exc}", flush=True)
    sys.exit(0)
try:
    t = symm.empty(64 << 20, dtype=torch.uint8, device=dev)
    h = symm.rendezvous(t, dist.group.WORLD)
    print(f"rank {rank}: multicast_ptr {h.multicast_ptr:#x} buffer_ptrs {[hex(p) for p in h.buffer_ptrs]} "
          f"signal_pad_size {h.signal_pad_size} buffer_size {h.buffer_size}", flush=True)
    t.fill_(rank + 1)
    h.barrier(channel=0)
    seen = [int(h.get_buffer(q, (16,), torch.uint8)[0].item()) for q in range(world)]
    print(f"rank {rank}: first byte of every peer's buffer {seen}", flush=True)
    h.barrier(channel=0)
except Exception as exc:  # noqa: BLE001
    print(f"rank {rank}: symmetric memory failed: {type(exc).__name__}: {exc}", flush=True)
dist.barrier()
dist.destroy_process_group()
