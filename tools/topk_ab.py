"""MaxK top-k kernel timing on the Reddit row count (A/B: MAXK_TOPK_STRIDED=1 selects the round-1 mapping).
usage: python tools/topk_ab.py [label [D:k]]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spgemm_gnn_b200 import maxk_kernels as mk  # noqa: E402


def main():
    label = sys.argv[1] if len(sys.argv) > 1 else os.environ.get("MAXK_TOPK_STRIDED", "0")
    n = 232965
    gen = torch.Generator(device="cuda").manual_seed(97)
    shapes = ((256, (8, 16, 32, 64)), (128, (32,)), (384, (32,)), (512, (32,)), (1024, (64,)), (250, (32,)))
    if len(sys.argv) > 2:                    # "256:32" -- one shape only (the ncu target)
        d_, k_ = sys.argv[2].split(":")
        shapes = ((int(d_), (int(k_),)),)
    for d, ks in shapes:
        x = torch.randn(n, d, device="cuda", generator=gen)
        for k in ks:
            for _ in range(3):
                mk.maxk_forward_cbsr(x, k)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(20):
                mk.maxk_forward_cbsr(x, k)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 20
            w = 1 if d <= 256 else 2
            gbs = (n * d * 4 + n * k * (4 + w)) / ms / 1e6
            print(f"topk[{label}] N={n} D={d} k={k}: {ms:.4f} ms  {gbs:.0f} GB/s", flush=True)
        del x


if __name__ == "__main__":
    main()
