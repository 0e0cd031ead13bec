"""Drop-in for the reference's ``utils/loss_func.py``: same names, arguments and return shapes,
computed by the sm_100a kernels in libcor_b200.so.  CUDA tensors only -- no CPU fallback.

    from cor_b200.loss_func import wbce_with_wiou_loss, fg_feat_similarity_loss, bg_feat_similarity_loss

Differences from the reference that callers can observe:
  * no host sync: validity of samples (``valid.any()`` at loss_func.py:76,109) is resolved on the
    device; when no sample is valid the returned zero still carries a (zero) grad_fn instead of being
    a grad-less ``torch.tensor(0.0)``;
  * ``fg_feat_similarity_loss`` and ``bg_feat_similarity_loss`` called back to back on the same
    tensors (utils/trainer_v3_g.py:69-71) share ONE pass over the feature map and the mask.  The half
    that was not asked for is parked (per thread) until the matching call consumes it; it is dropped as
    soon as the first result dies or a backward runs through it, so a caller that only ever uses one of
    the two keeps no graph alive beyond its own result.  The match is by tensor identity and
    ``_version``: inputs refilled behind autograd's back (CUDA-graph replay, raw memcpy) are not seen as
    changed -- such callers should use ``fg_bg_feat_similarity_loss`` / ``region_path_loss`` directly.
"""
from __future__ import annotations

import threading
import weakref

import torch

from . import ops
from ._lib import CorError

__all__ = ["wbce_with_wiou_loss", "mask_pooling", "fg_feat_similarity_loss", "bg_feat_similarity_loss",
           "fg_bg_feat_similarity_loss", "segmentation_loss", "region_path_loss", "BG_REFERENCE", "BG_PAIRED",
           "bce_with_iou_loss", "bce_with_dice_loss", "wbce_with_wdice_loss", "focal_loss_with_iou_loss"]

BG_REFERENCE = 0   # what loss_func.py:120-123 computes (cosine over the broadcast row axis)
BG_PAIRED = 1      # per-sample cos(bg_b, comb_b) + 1, the form its docstring describes


def wbce_with_wiou_loss(pred: torch.Tensor, mask: torch.Tensor, w1: float = 1.0, w2: float = 1.0) -> torch.Tensor:
    """utils/loss_func.py:5-32.  pred logits [N,C,H,W]; mask [N,C,H,W] in [0,1] (a mask at another
    resolution is bilinearly resampled inside the kernel, fusing trainer_v3_g.py:67)."""
    return ops.seg_loss(pred, mask, w1, w2)


# ---- Class N: the dice / focal / plain variants whose NAMES survive in the reference's stale bytecode
# (utils/__pycache__/loss_func.cpython-310.pyc: bce_with_iou_loss, bce_with_dice_loss, wbce_with_wdice_loss,
# focal_loss_with_iou_loss, ...) but whose source and constants do not.  Textbook definitions, restated in
# oracle/aten_port.py; every one is the same single pass as wbce_with_wiou_loss with another coefficient vector, forward
# and backward.  Parity unpinned by the reference.
def bce_with_iou_loss(pred: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """mean(bce) + soft IoU loss per sample, averaged over samples."""
    return ops.seg_loss(pred, mask, coef=ops.seg_coef(bce=1.0, iou=1.0))


def bce_with_dice_loss(pred: torch.Tensor, mask: torch.Tensor, smooth: float = 1.0) -> torch.Tensor:
    """mean(bce) + soft dice loss 1 - (2 sum(pt) + smooth)/(sum(p) + sum(t) + smooth)."""
    return ops.seg_loss(pred, mask, coef=ops.seg_coef(bce=1.0, dice=1.0), dice_smooth=smooth)


def wbce_with_wdice_loss(pred: torch.Tensor, mask: torch.Tensor, smooth: float = 1.0) -> torch.Tensor:
    """Edge-weighted BCE (loss_func.py:18-22) + edge-weighted dice 1 - (2 sum(ptw) + smooth)/(sum((p+t)w) + smooth)."""
    return ops.seg_loss(pred, mask, coef=ops.seg_coef(wbce=1.0, wdice=1.0), dice_smooth=smooth)


def focal_loss_with_iou_loss(pred: torch.Tensor, mask: torch.Tensor, alpha: float = 0.25, gamma: float = 2.0) -> torch.Tensor:
    """Sigmoid focal loss mean(a_t (1 - p_t)^gamma bce) + soft IoU loss."""
    return ops.seg_loss(pred, mask, coef=ops.seg_coef(focal=1.0, iou=1.0), focal_alpha=alpha, focal_gamma=gamma)


def segmentation_loss(pred: torch.Tensor, query_mask: torch.Tensor, w1: float = 1.0, w2: float = 1.0) -> torch.Tensor:
    """utils/trainer_v3_g.py:67-68 in one kernel: resample the full-resolution mask + wbce/wiou."""
    return ops.seg_loss(pred, query_mask, w1, w2)


def mask_pooling(embeddings: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """utils/loss_func.py:35-56: [B,C,H,W] x [B,1,Hm,Wm] -> L2-normalised [B,1,C]."""
    if mask.dim() != 4 or mask.shape[1] != 1:
        raise CorError(f"mask_pooling: mask must be [B,1,H,W], got {tuple(mask.shape)}")
    return ops.region_pool(embeddings, mask, transform=ops.W_CLAMP, normalize=True, engine="stream").fg


_memo = threading.local()          # .entry = (key, which, pair): one parked half per thread


def _same_call(entry, tensors):
    """True if ``entry`` was produced from these very tensor objects, unmodified since."""
    refs, versions, grad_mode = entry[0]
    return (grad_mode == torch.is_grad_enabled() and all(r() is t for r, t in zip(refs, tensors))
            and versions == tuple(t._version for t in tensors))


def fg_bg_feat_similarity_loss(query_image_embeddings: torch.Tensor, comb_support_feat: torch.Tensor, query_mask: torch.Tensor,
                               bg_mode: int = BG_REFERENCE):
    """Both feature losses of utils/loss_func.py:59-126 from one pass: returns (fg_loss, bg_loss)."""
    if query_mask.dim() != 4 or query_mask.shape[1] != 1:
        raise CorError(f"query_mask must be [B,1,H,W], got {tuple(query_mask.shape)}")
    B = query_image_embeddings.shape[0]
    pool = ops.region_pool(query_image_embeddings, query_mask, transform=ops.W_CLAMP, normalize=True, pair=True, engine="stream")
    comb = comb_support_feat.reshape(B, -1)
    losses, _ = ops.fgbg_losses(pool.fg.reshape(B, -1), pool.bg.reshape(B, -1), comb, pool.stats, bg_mode)
    return losses[0], losses[1]


def _shared(query_image_embeddings, comb_support_feat, query_mask, which: int):
    """The trainer calls fg then bg on the same three tensors (trainer_v3_g.py:69-71): the first call
    runs the fused pass and parks the other half for the second.  The parked entry is keyed on the
    identity (weakref) and version of the tensor objects and is consumed by the matching call."""
    tensors = (query_image_embeddings, comb_support_feat, query_mask)
    hit = getattr(_memo, "entry", None)
    _memo.entry = None
    if hit is not None and hit[1] != which and _same_call(hit, tensors):
        return hit[2][which]
    pair = fg_bg_feat_similarity_loss(*tensors)
    key = (tuple(weakref.ref(t) for t in tensors), tuple(t._version for t in tensors), torch.is_grad_enabled())
    entry = (key, which, pair)
    _memo.entry = entry
    out = pair[which]

    def drop(*_):
        if getattr(_memo, "entry", None) is entry:
            _memo.entry = None

    # the parked half shares the pooling node with ``out``: once a backward has gone through ``out`` (its buffers are
    # freed) or ``out`` itself is gone, the other half must be recomputed rather than served from the memo
    if out.requires_grad:
        out.register_hook(lambda g: drop())
    out = out.view_as(out)            # a fresh tensor object whose lifetime is the caller's: finalizer below
    weakref.finalize(out, drop)
    return out


def fg_feat_similarity_loss(query_image_embeddings, comb_support_feat, query_mask) -> torch.Tensor:
    """utils/loss_func.py:59-85: 1 - mean over non-empty samples of cos(pooled fg, comb)."""
    return _shared(query_image_embeddings, comb_support_feat, query_mask, 0)


def bg_feat_similarity_loss(query_image_embeddings, comb_support_feat, query_mask) -> torch.Tensor:
    """utils/loss_func.py:88-126, value-identical to the reference including its broadcast at :120-123."""
    return _shared(query_image_embeddings, comb_support_feat, query_mask, 1)


def region_path_loss(pred_mask, query_image_embeddings, comb_support_feat, query_mask) -> torch.Tensor:
    """Loss composition of utils/trainer_v3_g.py:67-73: seg + 5*fg + 5*bg."""
    seg = segmentation_loss(pred_mask, query_mask)
    fg, bg = fg_bg_feat_similarity_loss(query_image_embeddings, comb_support_feat, query_mask)
    return seg + 5 * fg + 5 * bg
