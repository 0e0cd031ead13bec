"""CPU port of the reference region path on the same ATen op sequence (torch, differentiable).

TEST INFRASTRUCTURE ONLY (see oracle/np_oracle.py header): imported by tests/, smoke() and the
``cpu_baseline`` / ``torch_eager_gpu`` / ``--impl reference`` baseline legs of bench.py (as the thing the product is
compared WITH, never as the thing measured as ours), never by ``cor_b200/``.

Why a second oracle: the reference is Python and cannot travel to the GPU box, and the numpy
restatement (np_oracle.py) is single-threaded and has no autograd.  This port issues the same
ATen operators, in the same order, as the reference call sites it cites, so that
  (a) timing it on the box's host cores is a fair stand-in for the reference's own CPU cost
      (``cpu_baseline.kind == "port"``), and
  (b) ``torch.autograd`` through it yields the gradients the reference's backward produces,
      which is what the CUDA backward kernels are checked against.
It is pinned against the same golden vectors as np_oracle.py (values AND gradients,
tests/test_oracle_golden.py).  Multi-mask inputs use SURVEY 8c's expansion recipe
(``repeat_interleave`` of the feature map, masks flattened to [B*M,1,H,W]).
"""
from __future__ import annotations

import torch
import torch.nn.functional as tf


def _resize(x: torch.Tensor, hw) -> torch.Tensor:
    if tuple(x.shape[-2:]) == tuple(hw):
        return x
    return tf.interpolate(x, size=tuple(hw), mode="bilinear", align_corners=False)


def plain_masked_mean(feat: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """mask_adapter.py:13-25 (MaskedPooling.forward): [B,C,h,w] x [B,1,H,W] -> [B,C]."""
    m = _resize(mask, feat.shape[2:])
    return (feat * m).sum((2, 3)) / (m.sum((2, 3)) + 1e-8)


def softmax_map_pool(maps: torch.Tensor, feat: torch.Tensor, group: int) -> torch.Tensor:
    """mask_adapter.py:62-80 (MaskAdapterPooling tail): maps [B,N,h,w], feat [B,C,h,w] -> [B,N/group,C]."""
    B, C = feat.shape[:2]
    maps = _resize(maps, feat.shape[2:])
    N = maps.shape[1]
    w = tf.softmax(tf.logsigmoid(maps).view(B, N, -1), dim=-1)
    pooled = torch.bmm(w, feat.reshape(B, C, -1).permute(0, 2, 1))
    return pooled.reshape(B, N // group, group, C).mean(dim=-2).contiguous()


def unit_region_feature(emb: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """loss_func.py:35-56 (mask_pooling): -> [B,1,C] unit rows."""
    m = _resize(mask, emb.shape[2:]).clamp(min=0, max=1)
    pooled = (emb * m).sum((2, 3)) / (m.sum((2, 3)) + 1e-8)
    return tf.normalize(pooled, p=2, dim=-1).unsqueeze(1)


def fg_loss(emb, comb, mask) -> torch.Tensor:
    """loss_func.py:59-85."""
    keep = mask.sum(dim=(1, 2, 3)) > 0
    if not keep.any():
        return torch.tensor(0.0, device=emb.device)
    cos = tf.cosine_similarity(unit_region_feature(emb[keep], mask[keep]), comb[keep], dim=-1)
    return 1 - cos.mean()


def bg_loss(emb, comb, mask) -> torch.Tensor:
    """loss_func.py:88-126, broadcasting of :120-123 preserved ([V,1,C] vs [V,C], dim=1)."""
    inv = 1 - mask
    keep = inv.sum(dim=(1, 2, 3)) > 0
    if not keep.any():
        return torch.tensor(0.0, device=emb.device)
    region = unit_region_feature(emb[keep], inv[keep])
    cos = tf.cosine_similarity(region, comb[keep].squeeze(1), dim=1)
    return (cos + 1).mean()


def edge_weighted_seg_loss(pred, target, w1: float = 1.0, w2: float = 1.0) -> torch.Tensor:
    """loss_func.py:5-32 (wbce_with_wiou_loss), pred / target at the same size."""
    edge = 1 + 5 * (tf.avg_pool2d(target, kernel_size=31, stride=1, padding=15) - target).abs()
    bce = tf.binary_cross_entropy_with_logits(pred, target, reduction="none")
    wbce = (edge * bce).sum(dim=(2, 3)) / edge.sum(dim=(2, 3))
    prob = torch.sigmoid(pred)
    inter = (prob * target * edge).sum(dim=(2, 3))
    union = ((prob + target) * edge).sum(dim=(2, 3)) - inter
    wiou = 1 - (inter + 1e-6) / (union + 1e-6)
    return (w1 * wbce + w2 * wiou).mean()


# ---- Class N segmentation-loss variants (names from the reference's stale .pyc, textbook definitions) ----
def _soft_terms(pred, target, smooth: float = 1.0, alpha: float = 0.25, gamma: float = 2.0):
    """Per-sample terms [N*C] of the variants below: bce, iou, dice, wbce, wdice, focal (target at pred's size)."""
    edge = 1 + 5 * (tf.avg_pool2d(target, kernel_size=31, stride=1, padding=15) - target).abs()
    bce = tf.binary_cross_entropy_with_logits(pred, target, reduction="none")
    p = torch.sigmoid(pred)
    d = (2, 3)
    inter, sp, stt = (p * target).sum(d), p.sum(d), target.sum(d)
    iw, uw = (p * target * edge).sum(d), ((p + target) * edge).sum(d)
    pt = p * target + (1 - p) * (1 - target)
    at = alpha * target + (1 - alpha) * (1 - target)
    return {"bce": bce.mean(d), "iou": 1 - (inter + 1e-6) / (sp + stt - inter + 1e-6),
            "dice": 1 - (2 * inter + smooth) / (sp + stt + smooth), "wbce": (edge * bce).sum(d) / edge.sum(d),
            "wdice": 1 - (2 * iw + smooth) / (uw + smooth), "wiou": 1 - (iw + 1e-6) / (uw - iw + 1e-6),
            "focal": (at * (1 - pt).clamp(min=0) ** gamma * bce).mean(d)}


def bce_with_iou_loss(pred, target):
    t = _soft_terms(pred, target)
    return (t["bce"] + t["iou"]).mean()


def bce_with_dice_loss(pred, target, smooth: float = 1.0):
    t = _soft_terms(pred, target, smooth=smooth)
    return (t["bce"] + t["dice"]).mean()


def wbce_with_wdice_loss(pred, target, smooth: float = 1.0):
    t = _soft_terms(pred, target, smooth=smooth)
    return (t["wbce"] + t["wdice"]).mean()


def focal_loss_with_iou_loss(pred, target, alpha: float = 0.25, gamma: float = 2.0):
    t = _soft_terms(pred, target, alpha=alpha, gamma=gamma)
    return (t["focal"] + t["iou"]).mean()


def seg_loss_fullres(pred, query_mask) -> torch.Tensor:
    """trainer_v3_g.py:67-68."""
    return edge_weighted_seg_loss(pred, _resize(query_mask, pred.shape[2:]))


def trainer_loss(pred, emb, comb, query_mask) -> torch.Tensor:
    """trainer_v3_g.py:67-73: seg + 5*fg + 5*bg."""
    return seg_loss_fullres(pred, query_mask) + 5 * fg_loss(emb, comb, query_mask) + 5 * bg_loss(emb, comb, query_mask)


def val_post(pred: torch.Tensor, hw=None) -> torch.Tensor:
    """trainer_v3_g.py:226-231 / vailder.py:427-430."""
    if hw is not None:
        pred = _resize(pred, hw)
    p = torch.sigmoid(pred)
    lo = torch.amin(p, dim=(1, 2, 3), keepdim=True)
    hi = torch.amax(p, dim=(1, 2, 3), keepdim=True)
    return (p - lo) / (hi - lo + 1e-8)


# ---- Class N (no reference implementation; parity unpinned by the reference) ---------------
def multi_mask_regions(emb: torch.Tensor, masks: torch.Tensor, background: bool = False) -> torch.Tensor:
    """[B,C,h,w] x [B,M,H,W] -> unit rows [B,M,C] through the M=1 function above."""
    B, M = masks.shape[:2]
    flat = masks.reshape(B * M, 1, *masks.shape[2:])
    if background:
        flat = 1 - flat
    return unit_region_feature(emb.repeat_interleave(M, 0), flat).reshape(B, M, -1)


def similarity(regions: torch.Tensor, queries: torch.Tensor, bf16_operands: bool = True) -> torch.Tensor:
    """[N_r,D], [N_q,D] -> S [N_q,N_r]; operands rounded to bf16 (straight-through) first."""
    if bf16_operands:
        regions = regions + (regions.bfloat16().to(regions.dtype) - regions).detach()
        queries = queries + (queries.bfloat16().to(queries.dtype) - queries).detach()
    return queries @ regions.t()


def infonce(regions, queries, targets, tau: float = 0.07, bf16_operands: bool = True) -> torch.Tensor:
    return tf.cross_entropy(similarity(regions, queries, bf16_operands) / tau, targets)


def region_step_loss(pred, emb, comb, masks, tau: float = 0.07, nce_weight: float = 1.0,
                     regions_all=None, target_offset: int = 0):
    """The bench 'step' on CPU: reference trainer loss on the GT mask (mask 0 of each image) plus
    the Class-N InfoNCE of each composed query against all B*M candidate regions."""
    B, M = masks.shape[:2]
    base = trainer_loss(pred, emb, comb, masks[:, 0:1])
    regions = multi_mask_regions(emb, masks).reshape(B * M, -1)
    if regions_all is None:
        regions_all = regions
    targets = torch.arange(B, device=masks.device) * M + target_offset
    nce = infonce(regions_all, comb[:, 0, :].float(), targets, tau)
    return base + nce_weight * nce, regions
