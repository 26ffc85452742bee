"""Host <-> device copy rates of the bench's e2e buffers (pinned, GPU-local CPUs): H2D alone, D2H alone, both at
once, two H2D streams -- what bounds bench.py's end-to-end step (477 MB in, 268 MB out)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

n, d, k = 232965, 256, 32
with bench.gpu_local_cpus(0):
    hx = torch.empty((n, d), dtype=torch.float32, pin_memory=True)
    hdy = torch.empty((n, d), dtype=torch.float32, pin_memory=True)
    hout = torch.empty((n, d), dtype=torch.float32, pin_memory=True)
    hdxs = torch.empty((n, k), dtype=torch.float32, pin_memory=True)
hx.fill_(1.0); hdy.fill_(2.0)
dx, dy, out = (torch.empty((n, d), device="cuda") for _ in range(3))
dxs = torch.empty((n, k), device="cuda")
s1, s2, s3 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()


def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for s in (s1, s2, s3):
        s.wait_event(a)
    for _ in range(reps):
        fn()
    for s in (s1, s2, s3):
        torch.cuda.current_stream().wait_stream(s)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def h2d_one():
    with torch.cuda.stream(s1):
        dx.copy_(hx, non_blocking=True); dy.copy_(hdy, non_blocking=True)


def h2d_two():
    with torch.cuda.stream(s1):
        dx.copy_(hx, non_blocking=True)
    with torch.cuda.stream(s2):
        dy.copy_(hdy, non_blocking=True)


def d2h():
    with torch.cuda.stream(s3):
        hout.copy_(out, non_blocking=True); hdxs.copy_(dxs, non_blocking=True)


def both():
    h2d_one(); d2h()


inb, outb = 2 * n * d * 4, n * d * 4 + n * k * 4
for name, fn, nb in (("H2D one stream", h2d_one, inb), ("H2D two streams", h2d_two, inb), ("D2H", d2h, outb),
                     ("H2D + D2H at once", both, inb)):
    ms = t(fn)
    print(f"{name}: {ms:.3f} ms per step, {nb / ms / 1e6:.1f} GB/s ({'in' if nb == inb else 'out'} bytes)", flush=True)
