"""Full-graph training on a synthetic graph of one of the reference's dataset shapes -- the role
of `maxk_gnn_integrated.py` (and of `maxk_gnn_dgl.py`, the driver that produced every logged
number) with the DGL/OGB dataset loaders replaced by a seeded generator (no network, no DGL).

    python -m spgemm_gnn_b200.train --dataset reddit --model sage --maxk 32 --epochs 50

Flags keep the names of `utils/config.py:30-70`.  With torchrun the graph is row-partitioned
over the ranks (see `dist.py`); weights are replicated and their gradients all-reduced.
"""
from __future__ import annotations

import argparse
import json
import os
import time

import torch
import torch.nn as nn
import torch.nn.functional as F

from .graph import FEATS, SHAPES, CSRGraph, shaped_graph
from .models import MODELS


def synthetic_task(name: str, scale: float, device, seed: int = 97):
    """Graph + random features / labels / 66-10-24 split with the logged in_feats and class
    counts of the real dataset.  Accuracy is meaningless; losses are not."""
    g = shaped_graph(name, scale=scale, seed=seed, device=device)
    n = g.num_nodes()
    in_feats, classes = FEATS[name]
    gen = torch.Generator(device=device).manual_seed(seed + 1)
    feats = torch.randn(n, in_feats, device=device, generator=gen)
    labels = torch.randint(0, classes, (n,), device=device, generator=gen)
    r = torch.rand(n, device=device, generator=gen)
    train_mask, val_mask = r < 0.66, (r >= 0.66) & (r < 0.76)
    return g, feats, labels, train_mask, val_mask, ~(train_mask | val_mask), in_feats, classes


def train_epochs(model: nn.Module, g: CSRGraph, feats, labels, train_mask, epochs: int,
                 lr: float = 0.01, weight_decay: float = 0.0, eval_every: int = 0, log=None):
    """The loop of maxk_gnn_dgl.py:98-134: one full-graph forward + backward per epoch (+ an
    eval forward every `eval_every` epochs).  Returns (losses, seconds per epoch).

    With a `dist.ShardedGraph` every rank passes ITS rows of feats / labels / mask: the loss is
    the sum over the local train nodes divided by the global count, the weight gradients are
    all-reduced, and every rank takes the same optimizer step."""
    opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=lr,
                           weight_decay=weight_decay)
    losses, times = [], []
    from .models import align_gemm, pad_features
    if align_gemm():
        feats = pad_features(feats)      # zero columns up to a multiple of 8, once (MAXK_ALIGN_GEMM)
    cuda = feats.is_cuda
    sharded = getattr(g, "world", 1) > 1
    if sharded:
        import torch.distributed as dist
        from .dist import allreduce_grads
        count = train_mask.sum().to(torch.float32)
        dist.all_reduce(count, group=g.group)
    for ep in range(epochs):
        model.train()
        if cuda:
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        logits = model(g, feats)
        if sharded:
            loss = F.cross_entropy(logits[train_mask], labels[train_mask], reduction="sum") / count
        else:
            loss = F.cross_entropy(logits[train_mask], labels[train_mask])
        opt.zero_grad(set_to_none=True)
        loss.backward()
        if sharded:
            allreduce_grads(model.parameters(), g.group)
            loss = loss.detach().clone()
            dist.all_reduce(loss, group=g.group)
        opt.step()
        if eval_every and (ep + 1) % eval_every == 0:
            model.eval()
            with torch.no_grad():
                model(g, feats)
        if cuda:
            torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
        losses.append(float(loss.detach()))
        if log:
            log(f"epoch {ep:4d} loss {losses[-1]:.6f} time {times[-1] * 1e3:.2f} ms")
    return losses, times


def train_epochs_graphed(model: nn.Module, g: CSRGraph, feats, labels, train_mask, epochs: int,
                         lr: float = 0.01, weight_decay: float = 0.0, warmup: int = 3,
                         eval_forward: bool = False):
    """Same training step, captured ONCE in a CUDA graph and replayed: on small graphs (Flickr
    shape) an epoch is ~60 short kernels and launch overhead, not the GPU, sets the pace.  The
    hot-path kernels are plain stream launches through the C ABI, so they capture like any other
    kernel; the work records are built (and synchronised on) before the capture.  With a
    `dist.ShardedGraph` the NCCL collectives of the step are captured too.
    The first `warmup` epochs run eagerly on a side stream (they train too).  With `eval_forward`
    the per-epoch evaluation forward of the reference loop (maxk_gnn_dgl.py:134) is a second graph."""
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.Adam(params, lr=lr, weight_decay=weight_decay, capturable=True)
    idx = train_mask.nonzero(as_tuple=True)[0]
    target = labels[idx]
    losses, times = [], []
    model.train()
    from .models import align_gemm, pad_features
    if align_gemm():
        feats = pad_features(feats)      # zero columns up to a multiple of 8, once (MAXK_ALIGN_GEMM)
    static_loss = torch.zeros((), device=feats.device)
    sharded = getattr(g, "world", 1) > 1
    if sharded:
        import torch.distributed as dist
        from .dist import allreduce_grads
        count = train_mask.sum().to(torch.float32)
        dist.all_reduce(count, group=g.group)

    def step():
        logits = model(g, feats)
        if sharded:
            loss = F.cross_entropy(logits.index_select(0, idx), target, reduction="sum") / count
        else:
            loss = F.cross_entropy(logits.index_select(0, idx), target)
        opt.zero_grad(set_to_none=False)
        loss.backward()
        if sharded:
            allreduce_grads(params, g.group)
        static_loss.copy_(loss.detach())
        if sharded:
            dist.all_reduce(static_loss, group=g.group)
        opt.step()

    def eval_step():
        model.eval()
        with torch.no_grad():
            model(g, feats)
        model.train()

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(min(warmup, epochs)):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            step()
            if eval_forward:
                eval_step()
            torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
            losses.append(float(static_loss))
    torch.cuda.current_stream().wait_stream(side)
    if epochs <= warmup:
        return losses, times
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        step()
    eval_graph = None
    if eval_forward:
        eval_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(eval_graph):
            eval_step()
    # the captures themselves did not execute anything
    for _ in range(epochs - warmup):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        graph.replay()
        if eval_graph is not None:
            eval_graph.replay()
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
        losses.append(float(static_loss))
    return losses, times


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--dataset", default="flickr", choices=sorted(SHAPES))
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--model", default="sage", choices=sorted(MODELS))
    ap.add_argument("--epochs", type=int, default=50)
    ap.add_argument("--w_lr", type=float, default=0.01)
    ap.add_argument("--w_weight_decay", type=float, default=0.0)
    ap.add_argument("--hidden_dim", type=int, default=256)
    ap.add_argument("--hidden_layers", type=int, default=3)
    ap.add_argument("--nonlinear", default="maxk", choices=["maxk", "relu"])
    ap.add_argument("--maxk", type=int, default=32)
    ap.add_argument("--dropout", type=float, default=0.5)
    ap.add_argument("--norm", action="store_true")
    ap.add_argument("--gpu", type=int, default=0)
    ap.add_argument("--seed", type=int, default=97)
    ap.add_argument("--shard_balance", default="rows", choices=["rows", "nnz"],
                    help="multi-GPU row partition: equal rows (random node orders) or equal stored entries")
    ap.add_argument("--no_tf32", dest="tf32", action="store_false",
                    help="fp32 GEMMs (default: TF32, as the reference sets at maxk_gnn_dgl.py:30-33)")
    ap.add_argument("--eval_every", type=int, default=1, help="eval forward every n epochs (reference: 1)")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--cuda_graph", action="store_true",
                    help="capture the train step (and the eval forward) in CUDA graphs")
    a = ap.parse_args(argv)
    if not torch.cuda.is_available():
        raise SystemExit("training needs a CUDA device: the aggregation has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1:
        a.gpu = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(a.gpu)
    dev = torch.device("cuda", a.gpu)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(a.seed)
    torch.backends.cuda.matmul.allow_tf32 = a.tf32
    torch.backends.cudnn.allow_tf32 = a.tf32
    g, feats, labels, train_mask, _, _, in_feats, classes = synthetic_task(a.dataset, a.scale, dev, a.seed)
    model = MODELS[a.model](in_feats, a.hidden_dim, a.hidden_layers, classes, maxk=a.maxk,
                            feat_drop=a.dropout, norm=a.norm, nonlinear=a.nonlinear).to(dev)
    n_nodes, n_edges = g.num_nodes(), g.num_edges()
    if world > 1:  # 1-D row partition: every rank keeps its rows of the graph and of the node data
        from .dist import ShardedGraph
        sg = ShardedGraph(g, rank, world, balance=a.shard_balance)
        feats, labels, train_mask = sg.local_rows(feats), sg.local_rows(labels), sg.local_rows(train_mask)
        g = sg
        # weights were initialised from the common seed above; from here on every rank draws its own
        # random numbers, so that the dropout masks of the shards are independent
        torch.manual_seed(a.seed + 1000003 * (rank + 1))
    say = print if rank == 0 else (lambda *_: None)
    say(f"{a.dataset}: {n_nodes} nodes, {n_edges} edges; model {a.model} "
        f"{sum(p.numel() for p in model.parameters())} params; {world} GPU(s)")
    if a.cuda_graph:
        losses, times = train_epochs_graphed(model, g, feats, labels, train_mask, a.epochs, a.w_lr,
                                             a.w_weight_decay, eval_forward=bool(a.eval_every))
    else:
        losses, times = train_epochs(model, g, feats, labels, train_mask, a.epochs, a.w_lr,
                                     a.w_weight_decay, eval_every=a.eval_every,
                                     log=say if a.verbose else None)
    steady = sorted(times[len(times) // 5:])
    say(json.dumps({"dataset": a.dataset, "model": a.model, "nonlinear": a.nonlinear, "maxk": a.maxk,
                    "gpus": world, "epochs": a.epochs, "first_loss": losses[0], "final_loss": losses[-1],
                    "epoch_ms_median": steady[len(steady) // 2] * 1e3,
                    "eval_forward_per_epoch": bool(a.eval_every), "cuda_graph": bool(a.cuda_graph)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
