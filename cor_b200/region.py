"""Region-level retrieval and scoring (north-star extensions, "Class N" in SURVEY.md section 0):
multi-mask region pooling, the full region x query similarity matrix, InfoNCE over (all-gathered)
negatives, top-k retrieval, and the fused training step the benchmark times.

None of these has a reference implementation; where they overlap the reference they reduce to it:
row (b,m) of ``pool_regions`` equals ``loss_func.mask_pooling(emb[b], masks[b,m])`` and the target
column of the similarity matrix equals the cosine inside ``fg_feat_similarity_loss``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

from . import dist as cdist
from . import ops
from ._lib import CorError

__all__ = ["pool_regions", "region_query_similarity", "region_infonce_loss", "topk_regions", "region_step", "RegionStepOut",
           "StepBuffers"]


def pool_regions(emb: torch.Tensor, masks: torch.Tensor, background: bool = False, engine: str = "auto", want_bf16: bool = True):
    """emb [B,C,h,w] x candidate masks [B,M,H,W] -> ops.RegionPool with unit rows fg [B,M,C]
    (and bg when ``background``), plus per-mask stats (validity, denominators)."""
    return ops.region_pool(emb, masks, transform=ops.W_CLAMP, normalize=True, pair=background, engine=engine, want_bf16=want_bf16)


def region_query_similarity(regions: torch.Tensor, queries: torch.Tensor, engine: str = "auto") -> torch.Tensor:
    """[Nr,D] unit region rows x [Nq,D] unit queries -> cosine matrix S [Nq,Nr] (bf16 operands, fp32 accumulate)."""
    return ops.similarity(regions, queries, engine)


def region_infonce_loss(regions: torch.Tensor, queries: torch.Tensor, targets: torch.Tensor, tau: float = 0.07,
                        gather: bool = True, engine: str = "auto", regions_bf16: Optional[torch.Tensor] = None) -> torch.Tensor:
    """InfoNCE of each local query against ALL regions.  With ``gather`` (and an initialised process
    group) the region rows are all-gathered across ranks first and ``targets`` -- indices into the
    LOCAL region rows -- are offset by rank * n_local.  tau is a build choice (the reference defines none)."""
    regions = regions.reshape(-1, regions.shape[-1])
    rank, ws = cdist.world()
    if gather and ws > 1:
        n_local = regions.shape[0]
        regions = cdist.all_gather_rows(regions)
        targets = targets + rank * n_local
        regions_bf16 = None
    return ops.infonce_loss(regions, queries.reshape(-1, queries.shape[-1]), targets, tau, engine, regions_bf16)


def topk_regions(regions: torch.Tensor, queries: torch.Tensor, k: int, sharded: bool = False, engine: str = "auto"):
    """Top-k regions per query, total order (score desc, index asc).  With ``sharded`` each rank holds
    a contiguous, equal shard of the gallery: local top-k, all-gather of k*(score,idx), k-way merge."""
    idx, score = ops.topk_retrieve(regions, queries, min(k, regions.shape[0]), engine)
    rank, ws = cdist.world()
    if sharded and ws > 1:
        return cdist.merge_topk(idx, score, rank * regions.shape[0], k)
    return idx, score


@dataclass
class RegionStepOut:
    loss: torch.Tensor
    seg: torch.Tensor
    fg: torch.Tensor
    bg: torch.Tensor
    nce: torch.Tensor
    regions: torch.Tensor          # [B, M, C] unit rows (detached view for retrieval)


def region_step(pred: torch.Tensor, emb: torch.Tensor, comb: torch.Tensor, masks: torch.Tensor, *, tau: float = 0.07,
                nce_weight: float = 1.0, gather: bool = True, pool_engine: str = "auto", sim_engine: str = "auto",
                bg_mode: int = 0) -> RegionStepOut:
    """One forward of the whole region path for a batch of triplets with M candidate masks each
    (mask 0 of every image is the ground-truth ``query_mask``):

        seg  = wbce_with_wiou_loss(pred, resample(masks[:,0]))                 trainer_v3_g.py:67-68
        fg,bg = fg/bg_feat_similarity_loss(emb, comb, masks[:,0])              trainer_v3_g.py:69-71
        nce  = InfoNCE(comb_b vs all B*M (x world) pooled regions, target = region (b,0))   [Class N]
        loss = seg + 5 fg + 5 bg + nce_weight * nce

    ONE pass over the masks and ONE pass over the feature map serve fg, bg and all M regions.
    """
    B, M = masks.shape[:2]
    if comb.shape[0] != B or emb.shape[0] != B or pred.shape[0] != B:
        raise CorError("region_step: batch sizes differ")
    pool = ops.region_pool(emb, masks, transform=ops.W_CLAMP, normalize=True, pair=True, engine=pool_engine, want_bf16=True)
    Cc = pool.fg.shape[-1]
    q = comb.reshape(B, -1)
    gt_rows = torch.arange(B, device=emb.device) * M
    losses, _ = ops.fgbg_losses(pool.fg[:, 0, :], pool.bg[:, 0, :], q, pool.stats.view(B, M, 4)[:, 0, :], bg_mode)
    seg = ops.seg_loss(pred, masks[:, 0:1])
    nce = region_infonce_loss(pool.fg.reshape(B * M, Cc), q, gt_rows, tau, gather, sim_engine,
                              regions_bf16=pool.fg_bf16)
    loss = seg + 5 * losses[0] + 5 * losses[1] + nce_weight * nce
    return RegionStepOut(loss=loss, seg=seg, fg=losses[0], bg=losses[1], nce=nce, regions=pool.fg.detach())


class StepBuffers:
    """Static device buffers (+ a pinned loss slot) for the end-to-end call: the public entry a trainer
    uses when its batch arrives from a DataLoader on the host.  ``run`` copies this step's inputs
    host->device on the current stream, runs :func:`region_step` (+ backward) and returns the loss on
    the host, so the timed region of ``bench.py``'s e2e leg contains the copies.

    ``capture()`` records forward + backward of the step ONCE into a CUDA graph over these static
    buffers (launch-bound inner loop: ~30 small kernels); afterwards ``run`` / ``replay`` re-issue
    the whole step with a single graph launch.  Gradients land in ``self.grads``."""

    def __init__(self, B, M, C=256, h=64, w=64, H=1024, W=1024, hp=256, wp=256, D=None, device="cuda",
                 emb_dtype=torch.bfloat16, mask_dtype=torch.float32):
        D = C if D is None else D
        self.device = torch.device(device)
        mk = lambda shape, dt: torch.empty(shape, dtype=dt, device=self.device)
        self.d = {"pred": mk((B, 1, hp, wp), emb_dtype), "emb": mk((B, C, h, w), emb_dtype), "comb": mk((B, 1, D), torch.float32),
                  "masks": mk((B, M, H, W), mask_dtype)}
        self.loss_host = torch.empty((), dtype=torch.float32).pin_memory()
        self.graph = None
        self.loss = None
        self.grads = {}
        self.launches_per_step = 0

    @property
    def h2d_bytes(self):
        return sum(t.numel() * t.element_size() for t in self.d.values())

    def load(self, src: dict):
        """Copy one batch (host pinned or device tensors) into the static buffers on the current stream."""
        for k, t in self.d.items():
            t.copy_(src[k], non_blocking=True)

    def _step(self, backward, emb_grad, kw):
        pred = self.d["pred"].detach().requires_grad_(backward)
        comb = self.d["comb"].detach().requires_grad_(backward)
        emb = self.d["emb"].detach().requires_grad_(backward and emb_grad)
        out = region_step(pred, emb, comb, self.d["masks"], **kw)
        if backward:
            out.loss.backward()
        grads = {"pred": pred.grad, "comb": comb.grad, "emb": emb.grad}
        return out.loss.detach(), grads

    def capture(self, backward: bool = True, emb_grad: bool = True, warmup: int = 3, **kw):
        """Warm up on a side stream, then capture fwd+bwd of the step into a CUDA graph."""
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._step(backward, emb_grad, kw)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        n0 = ops.LAUNCHES["count"]
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.loss, self.grads = self._step(backward, emb_grad, kw)
        self.launches_per_step = ops.LAUNCHES["count"] - n0
        self.graph = g
        return self

    def replay(self):
        self.graph.replay()
        ops.LAUNCHES["count"] += self.launches_per_step
        return self.loss

    def run(self, host: dict, backward: bool = True, emb_grad: bool = True, **kw) -> float:
        self.load(host)
        if self.graph is not None:
            loss = self.replay()
        else:
            loss, self.grads = self._step(backward, emb_grad, kw)
        self.loss_host.copy_(loss, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(self.loss_host)
