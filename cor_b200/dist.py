"""Multi-GPU plumbing for the region path: one process per GPU over torch.distributed (NCCL on the
GPU box, gloo in the CPU tests).  The path is data-parallel by triplet: pooling, the
segmentation loss and the fg/bg losses are rank-local; the only exchange is the all-gather of the
(bf16) region embeddings so that contrastive negatives span every rank, and -- for retrieval -- the
all-gather of per-rank top-k lists followed by a k-way merge.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

__all__ = ["world", "all_gather_rows", "merge_topk", "shard_range"]


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


class _AllGatherRows(torch.autograd.Function):
    """all_gather along dim 0 with an autograd-correct backward.

    Every rank computes the SAME contrastive loss form on its own queries against ALL regions, so
    rank j's regions receive gradient from every rank: backward = reduce-scatter(sum) of the gathered
    gradient (implemented as all_reduce + slice when the backend lacks reduce_scatter_tensor)."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        ws = dist.get_world_size(group)
        ctx.rows = x.shape[0]
        x = x.contiguous()
        out = x.new_empty((ws * x.shape[0],) + tuple(x.shape[1:]))
        try:
            dist.all_gather_into_tensor(out, x, group=group)
        except (RuntimeError, NotImplementedError):   # gloo builds without the tensor variant
            parts = [torch.empty_like(x) for _ in range(ws)]
            dist.all_gather(parts, x, group=group)
            out = torch.cat(parts, 0)
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        rank = dist.get_rank(ctx.group)
        if g.is_cuda:
            out = g.new_empty((ctx.rows,) + tuple(g.shape[1:]))
            dist.reduce_scatter_tensor(out, g, op=dist.ReduceOp.SUM, group=ctx.group)
            return out, None
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=ctx.group)
        return g[rank * ctx.rows:(rank + 1) * ctx.rows].clone(), None


def all_gather_rows(x: torch.Tensor, group=None) -> torch.Tensor:
    """[n_local, ...] -> [world * n_local, ...] (rank-major); identity when not distributed.
    Equal n_local on every rank is required (the batch is evenly sharded)."""
    _, ws = world()
    if ws == 1:
        return x
    return _AllGatherRows.apply(x, group)


def shard_range(n: int, rank: int, ws: int):
    """Contiguous, balanced [lo, hi) shard of n items for `rank`."""
    base, rem = divmod(n, ws)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def merge_topk(idx_local: torch.Tensor, score_local: torch.Tensor, offset: int, k: int, group=None):
    """Gallery sharded by region: every rank holds its local top-k (indices local to its shard,
    scores canonical fp32).  Gathers the k*(score, global idx) lists and merges them under the same
    total order the single-GPU kernel uses (score desc, index asc).  Returns ([Nq,k] idx, score)."""
    _, ws = world()
    gidx = idx_local + offset
    if ws == 1:
        return gidx[:, :k], score_local[:, :k]
    kk = idx_local.shape[1]
    idx_all = [torch.empty_like(gidx) for _ in range(ws)]
    sc_all = [torch.empty_like(score_local) for _ in range(ws)]
    dist.all_gather(idx_all, gidx.contiguous(), group=group)
    dist.all_gather(sc_all, score_local.contiguous(), group=group)
    idx_cat, sc_cat = torch.cat(idx_all, 1), torch.cat(sc_all, 1)          # [Nq, ws*kk]
    # total order: score desc, then global index asc -> two stable sorts (secondary key first)
    o1 = torch.argsort(idx_cat, dim=1, stable=True)
    idx_s, sc_s = idx_cat.gather(1, o1), sc_cat.gather(1, o1)
    o2 = torch.argsort(sc_s, dim=1, descending=True, stable=True)
    k = min(k, ws * kk)
    return idx_s.gather(1, o2)[:, :k], sc_s.gather(1, o2)[:, :k]
