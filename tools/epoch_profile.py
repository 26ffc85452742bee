#!/usr/bin/env python
"""Kernel-time breakdown of one training epoch (torch profiler), to see what is left around the
aggregation kernels."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from spgemm_gnn_b200.models import MODELS
from spgemm_gnn_b200.train import synthetic_task, train_epochs

ap = argparse.ArgumentParser()
ap.add_argument("--dataset", default="reddit"); ap.add_argument("--model", default="sage")
ap.add_argument("--tf32", action="store_true"); ap.add_argument("--scale", type=float, default=1.0)
a = ap.parse_args()
dev = torch.device("cuda")
torch.backends.cuda.matmul.allow_tf32 = a.tf32
g, x, y, m, _, _, fin, ncls = synthetic_task(a.dataset, a.scale, dev)
model = MODELS[a.model](fin, 256, 3, ncls, maxk=32, feat_drop=0.5, norm=True).to(dev)
train_epochs(model, g, x, y, m, 3, eval_every=1)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    _, times = train_epochs(model, g, x, y, m, 3, eval_every=1)
print("epoch ms", [round(t * 1e3, 2) for t in times])
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=32, max_name_column_width=160))
