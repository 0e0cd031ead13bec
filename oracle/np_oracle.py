"""CPU oracle for the CORE region pooling / scoring / loss path (numpy restatement).

TEST INFRASTRUCTURE ONLY.  Nothing under ``cor_b200/`` may import this module; only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may.  The product path is the CUDA library and fails loudly without it.

What this restates
------------------
The reference (wangtong627/COR, mounted at /root/reference in the build container) is pure
Python on PyTorch: every hot-path function is a short chain of ATen ops.  The arithmetic
therefore lives in a third-party dependency that is NOT under /root/reference: **PyTorch ATen**
(pinned ``torch==2.6.0`` in ``requirements.txt:79``; 2.11.0 in this image).  This file restates
the published algorithms of the ATen ops on the path (``upsample_bilinear2d`` with
``align_corners=False``, ``avg_pool2d`` with zero padding / count_include_pad,
``binary_cross_entropy_with_logits``, ``normalize``, ``cosine_similarity``, ``softmax`` /
``logsigmoid`` / ``bmm``) in plain numpy, then composes them exactly the way the reference call
sites do.  Each function cites the reference ``file:line`` it follows.

Parity pinning
--------------
Class R functions (those with a live reference implementation) are pinned against golden
vectors produced by importing the unmodified reference in the build container
(``oracle/gen_golden.py`` -> ``tests/golden/*.npz``; checked by ``tests/test_oracle_golden.py``).
Class N functions (region x query similarity matrix, InfoNCE, top-k, dice / focal) have NO
reference implementation (finding 2): **parity unpinned by the reference**
for those; they are pinned only by reduction identities to Class R (diag of the similarity
matrix == the reference's paired cosine, M=1 pooling == ``mask_pooling``).

All arithmetic is float32 unless a function says otherwise; reductions accumulate in float64
and round once, which is the "exactly rounded" value any fp32 summation order approximates.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------------------
# ATen building blocks
# --------------------------------------------------------------------------------------
def _src_index(out_size: int, in_size: int):
    """ATen ``area_pixel_compute_source_index`` for align_corners=False, non-cubic.

    scale = in/out computed in float32; src = scale*(dst+0.5)-0.5 clamped at 0;
    i0 = floor(src), i1 = i0 + (i0 < in-1), lam1 = src - i0, lam0 = 1 - lam1.
    (torch/aten/src/ATen/native/UpSample.h; called from the reference at
    mask_adapter.py:20,58,62,158, loss_func.py:47, trainer_v3_g.py:67,226.)
    """
    scale = F32(in_size) / F32(out_size)
    dst = np.arange(out_size, dtype=F32)
    src = scale * (dst + F32(0.5)) - F32(0.5)
    src = np.maximum(src, F32(0.0)).astype(F32)
    i0 = np.minimum(src.astype(np.int64), in_size - 1)
    i1 = i0 + (i0 < in_size - 1)
    lam1 = (src - i0.astype(F32)).astype(F32)
    lam0 = (F32(1.0) - lam1).astype(F32)
    return i0, i1, lam0, lam1


def bilinear_resize(x: np.ndarray, out_hw) -> np.ndarray:
    """``F.interpolate(x, size=out_hw, mode="bilinear", align_corners=False)`` (no antialias).

    x: [..., H, W] float32.  Same-size resize is the identity (lam1 == 0).
    """
    x = np.asarray(x, dtype=F32)
    H, W = x.shape[-2:]
    oh, ow = int(out_hw[0]), int(out_hw[1])
    y0, y1, wy0, wy1 = _src_index(oh, H)
    x0, x1, wx0, wx1 = _src_index(ow, W)
    top = x[..., y0, :]
    bot = x[..., y1, :]
    t = top[..., :, x0] * wx0 + top[..., :, x1] * wx1
    b = bot[..., :, x0] * wx0 + bot[..., :, x1] * wx1
    out = wy0[:, None] * t + wy1[:, None] * b
    return out.astype(F32)


def box_mean_31(t: np.ndarray, k: int = 31) -> np.ndarray:
    """``F.avg_pool2d(t, kernel_size=31, stride=1, padding=15)``: zero padding,
    count_include_pad=True, i.e. always divide by 31*31 (loss_func.py:18)."""
    t = np.asarray(t, dtype=np.float64)
    r = k // 2
    pad = [(0, 0)] * (t.ndim - 2) + [(r + 1, r), (r + 1, r)]
    c = np.pad(t, pad).cumsum(-2).cumsum(-1)
    H, W = t.shape[-2:]
    s = c[..., k:k + H, k:k + W] - c[..., 0:H, k:k + W] - c[..., k:k + H, 0:W] + c[..., 0:H, 0:W]
    return (s / float(k * k)).astype(F32)


def log_sigmoid(x):
    x = np.asarray(x, dtype=np.float64)
    return np.minimum(x, 0.0) - np.log1p(np.exp(-np.abs(x)))


def sigmoid(x):
    x = np.asarray(x, dtype=np.float64)
    return 1.0 / (1.0 + np.exp(-x))


def l2_normalize(x: np.ndarray, eps: float = 1e-12) -> np.ndarray:
    """``F.normalize(x, p=2, dim=-1)``: x / max(||x||_2, eps)  (support_branch.py:85,
    cir_feature_fuse.py:59, loss_func.py:53)."""
    x = np.asarray(x, dtype=np.float64)
    n = np.sqrt((x * x).sum(-1, keepdims=True))
    return (x / np.maximum(n, eps)).astype(F32)


def cosine_similarity(x1: np.ndarray, x2: np.ndarray, axis: int, eps: float = 1e-8) -> np.ndarray:
    """ATen ``cosine_similarity``: broadcast, then sum((x1/max(||x1||,eps)) * (x2/max(||x2||,eps)))
    along ``axis`` (ATen/native/Distance.cpp).  Broadcasting happens BEFORE the norms, which is
    what makes the reference's bg loss (below) what it is."""
    a, b = np.broadcast_arrays(np.asarray(x1, dtype=np.float64), np.asarray(x2, dtype=np.float64))
    na = np.maximum(np.sqrt((a * a).sum(axis, keepdims=True)), eps)
    nb = np.maximum(np.sqrt((b * b).sum(axis, keepdims=True)), eps)
    return ((a / na) * (b / nb)).sum(axis).astype(F32)


# --------------------------------------------------------------------------------------
# Class R: reference-pinned functions
# --------------------------------------------------------------------------------------
def masked_pooling(feat: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """``MaskedPooling.forward`` lib/support_model/mask_adapter.py:13-25.

    feat [B,C,h,w], mask [B,1,H,W] -> [B,C].  No clamp, no L2-norm, eps 1e-8 on the denominator.
    """
    feat = np.asarray(feat, dtype=F32)
    mask = np.asarray(mask, dtype=F32)
    if mask.shape[2:] != feat.shape[2:]:
        mask = bilinear_resize(mask, feat.shape[2:])
    num = (feat.astype(np.float64) * mask.astype(np.float64)).sum((2, 3))
    den = mask.astype(np.float64).sum((2, 3)) + 1e-8
    return (num / den).astype(F32)


def mask_adapter_pool_tail(maps: np.ndarray, feat: np.ndarray, num_output_maps: int) -> np.ndarray:
    """Pooling tail of ``MaskAdapterPooling.forward`` lib/support_model/mask_adapter.py:62-80.

    maps [B,N,h,w] (N = Q*num_output_maps), feat [B,C,h,w] -> [B,Q,C].
    W = softmax_P(logsigmoid(maps)); pooled = W @ feat^T; mean over each group of maps.
    (The same-size ``interpolate`` at :62-67 is an identity.)
    """
    maps = np.asarray(maps, dtype=F32)
    feat = np.asarray(feat, dtype=F32)
    B, N = maps.shape[:2]
    C = feat.shape[1]
    if maps.shape[2:] != feat.shape[2:]:
        maps = bilinear_resize(maps, feat.shape[2:])
    ls = log_sigmoid(maps.reshape(B, N, -1))
    ls = ls - ls.max(-1, keepdims=True)
    w = np.exp(ls)
    w = w / w.sum(-1, keepdims=True)
    pooled = np.einsum("bnp,bcp->bnc", w, feat.reshape(B, C, -1).astype(np.float64))
    q = N // num_output_maps
    return pooled.reshape(B, q, num_output_maps, C).mean(2).astype(F32)


def mask_pooling(emb: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """``mask_pooling`` utils/loss_func.py:35-56: resize, clamp(0,1), masked mean (+1e-8),
    L2-normalise -> [B,1,C]."""
    emb = np.asarray(emb, dtype=F32)
    mask = np.asarray(mask, dtype=F32)
    if mask.shape[2:] != emb.shape[2:]:
        mask = bilinear_resize(mask, emb.shape[2:])
    mask = np.clip(mask, 0.0, 1.0).astype(np.float64)
    num = (emb.astype(np.float64) * mask).sum((2, 3))
    den = mask.sum((2, 3)) + 1e-8
    pooled = (num / den).astype(F32)
    return l2_normalize(pooled)[:, None, :]


def fg_feat_similarity_loss(emb, comb, mask) -> np.float32:
    """``fg_feat_similarity_loss`` utils/loss_func.py:59-85.

    valid_b = sum(mask_b) > 0 on the FULL-resolution mask; pooled fg on valid rows;
    1 - mean_valid cos(pool_b, comb_b); no valid row -> 0 (grad-less in the reference).
    """
    mask = np.asarray(mask, dtype=F32)
    comb = np.asarray(comb, dtype=F32)
    valid = mask.astype(np.float64).sum((1, 2, 3)) > 0
    if not valid.any():
        return F32(0.0)
    q = mask_pooling(np.asarray(emb)[valid], mask[valid])          # [V,1,C]
    cos = cosine_similarity(q, comb[valid], axis=-1)               # [V,1]
    return F32(1.0 - cos.astype(np.float64).mean())


def bg_feat_similarity_loss(emb, comb, mask) -> np.float32:
    """``bg_feat_similarity_loss`` utils/loss_func.py:88-126 -- INCLUDING its broadcasting.

    bg = 1 - mask; valid_b = sum(bg_b) > 0; bg_feat = mask_pooling(...) is [V,1,C] but the
    support feature is squeezed to [V,C] (:120) and ``cosine_similarity(..., dim=1)`` (:123)
    broadcasts the pair to [V,V,C] and reduces over the *row* axis.  The value the reference
    returns is therefore NOT mean_b(cos(bg_b, comb_b)+1); it is

        mean_{i,c} [ bg[i,c]/max(sqrt(V)|bg[i,c]|,eps) * sum_j comb[j,c] / max(||comb[:,c]||,eps) ] + 1

    This oracle reproduces what the reference computes (verified against it bit-for-tolerance in
    tests/test_oracle_golden.py).  ``bg_feat_similarity_loss_paired`` below is the evidently
    intended per-sample form, offered by the product as an explicit option.
    """
    mask = np.asarray(mask, dtype=F32)
    comb = np.asarray(comb, dtype=F32)
    bg = (F32(1.0) - mask).astype(F32)
    valid = bg.astype(np.float64).sum((1, 2, 3)) > 0
    if not valid.any():
        return F32(0.0)
    feat = mask_pooling(np.asarray(emb)[valid], bg[valid])          # [V,1,C]
    sup = comb[valid][:, 0, :]                                      # [V,C]
    sim = cosine_similarity(feat, sup, axis=1)                      # broadcast [V,V,C] -> [V,C]
    return F32((sim.astype(np.float64) + 1.0).mean())


def bg_feat_similarity_loss_paired(emb, comb, mask) -> np.float32:
    """Per-sample background loss mean_valid(cos(bg_b, comb_b) + 1) -- the form the docstring at
    loss_func.py:88-101 describes.  Class N (not what the reference computes for B>1 or C>1)."""
    mask = np.asarray(mask, dtype=F32)
    comb = np.asarray(comb, dtype=F32)
    bg = (F32(1.0) - mask).astype(F32)
    valid = bg.astype(np.float64).sum((1, 2, 3)) > 0
    if not valid.any():
        return F32(0.0)
    feat = mask_pooling(np.asarray(emb)[valid], bg[valid])[:, 0, :]
    cos = cosine_similarity(feat, comb[valid][:, 0, :], axis=-1)
    return F32((cos.astype(np.float64) + 1.0).mean())


def wbce_with_wiou_loss(pred, mask, w1: float = 1.0, w2: float = 1.0) -> np.float32:
    """``wbce_with_wiou_loss`` utils/loss_func.py:5-32 (pred logits and mask at the same size).

    weit = 1 + 5|boxmean31(mask) - mask|; wbce = sum(weit*bce)/sum(weit);
    inter = sum(sig*mask*weit); union = sum((sig+mask)*weit) - inter;
    wiou = 1 - (inter+1e-6)/(union+1e-6); mean over [N,C] of w1*wbce + w2*wiou.
    """
    x = np.asarray(pred, dtype=F32).astype(np.float64)
    t = np.asarray(mask, dtype=F32).astype(np.float64)
    weit = 1.0 + 5.0 * np.abs(box_mean_31(t).astype(np.float64) - t)
    bce = (1.0 - t) * x - log_sigmoid(x)
    wbce = (weit * bce).sum((2, 3)) / weit.sum((2, 3))
    p = sigmoid(x)
    inter = (p * t * weit).sum((2, 3))
    union = ((p + t) * weit).sum((2, 3)) - inter
    wiou = 1.0 - (inter + 1e-6) / (union + 1e-6)
    return F32((w1 * wbce + w2 * wiou).mean())


def segmentation_loss(pred, query_mask, w1: float = 1.0, w2: float = 1.0) -> np.float32:
    """Trainer call site utils/trainer_v3_g.py:67-68: bilinear-resample the full-resolution query
    mask to the logit size, then ``wbce_with_wiou_loss``."""
    pred = np.asarray(pred, dtype=F32)
    target = bilinear_resize(np.asarray(query_mask, dtype=F32), pred.shape[2:])
    return wbce_with_wiou_loss(pred, target, w1, w2)


def region_path_loss(pred, emb, comb, query_mask) -> np.float32:
    """Loss composition utils/trainer_v3_g.py:67-73: seg + 5*fg + 5*bg."""
    seg = np.float64(segmentation_loss(pred, query_mask))
    fg = np.float64(fg_feat_similarity_loss(emb, comb, query_mask))
    bg = np.float64(bg_feat_similarity_loss(emb, comb, query_mask))
    return F32(seg + 5.0 * fg + 5.0 * bg)


def val_postprocess(pred, out_hw=None) -> np.ndarray:
    """Validation post-process utils/trainer_v3_g.py:226-231 (with upsample) and
    utils/vailder.py:427-430 (without): optional bilinear resize, sigmoid, per-sample min-max
    stretch with +1e-8."""
    x = np.asarray(pred, dtype=F32)
    if out_hw is not None and tuple(out_hw) != x.shape[2:]:
        x = bilinear_resize(x, out_hw)
    p = sigmoid(x).astype(F32)
    mn = p.min((1, 2, 3), keepdims=True)
    mx = p.max((1, 2, 3), keepdims=True)
    return ((p - mn) / (mx - mn + F32(1e-8))).astype(F32)


def vailder_postprocess(pred, gt_hw) -> np.ndarray:
    """Offline evaluator order, utils/vailder.py:427-430 then :466: sigmoid + per-sample min-max at the logit
    size, then ``cv2.resize(..., INTER_LINEAR)`` of each normalised map to the ground-truth size.  For float32
    input cv2's INTER_LINEAR samples with the same half-pixel rule and edge clamps as align_corners=False."""
    return bilinear_resize(val_postprocess(pred), gt_hw)


def binarize(p: np.ndarray) -> np.ndarray:
    """utils/vailder.py:473: (p > 0.5) * 255 as uint8."""
    return ((np.asarray(p) > 0.5).astype(np.uint8) * 255).astype(np.uint8)


def soft_metrics(pred: np.ndarray, gt: np.ndarray, smooth: float = 1e-5):
    """``compute_dice / mae / iou / mdice / miou`` utils/trainer_v3_g.py:381-443 -> dict of [B]."""
    p = np.asarray(pred, dtype=np.float64).reshape(pred.shape[0], -1)
    g = np.asarray(gt, dtype=np.float64).reshape(gt.shape[0], -1)

    def dice(a, b):
        return (2.0 * (a * b).sum(1) + smooth) / (a.sum(1) + b.sum(1) + smooth)

    def iou(a, b):
        i = (a * b).sum(1)
        return (i + smooth) / (a.sum(1) + b.sum(1) - i + smooth)

    return {
        "dice": dice(p, g).astype(F32),
        "mae": np.abs(p - g).mean(1).astype(F32),
        "iou": iou(p, g).astype(F32),
        "mdice": ((dice(p, g) + dice(1 - p, 1 - g)) / 2).astype(F32),
        "miou": ((iou(p, g) + iou(1 - p, 1 - g)) / 2).astype(F32),
    }


# --------------------------------------------------------------------------------------
# Class N: north-star extensions (no reference implementation; parity unpinned by the
# reference, pinned by reduction identities to Class R in tests/test_oracle_classn.py)
# --------------------------------------------------------------------------------------
def to_bf16(x: np.ndarray) -> np.ndarray:
    """Round float32 to bfloat16 (round-to-nearest-even) and return it widened back to float32."""
    u = np.ascontiguousarray(np.asarray(x, dtype=F32)).view(np.uint32)
    lsb = (u >> np.uint32(16)) & np.uint32(1)
    r = (u + np.uint32(0x7FFF) + lsb) & np.uint32(0xFFFF0000)
    nan = np.isnan(np.asarray(x, dtype=F32))
    out = r.view(F32).copy()
    out[nan] = np.nan
    return out


def multi_mask_pool(emb, masks, clamp: bool = True, normalize: bool = True, background: bool = False):
    """Multi-mask region pooling via the reference's own M=1 functions (SURVEY 8c recipe):
    emb [B,C,h,w], masks [B,M,H,W] -> [B,M,C]; row (b,m) == mask_pooling(emb[b], masks[b,m])."""
    emb = np.asarray(emb, dtype=F32)
    masks = np.asarray(masks, dtype=F32)
    B, M = masks.shape[:2]
    out = np.empty((B, M, emb.shape[1]), dtype=F32)
    for b in range(B):
        m = masks[b][:, None]
        if background:
            m = (F32(1.0) - m).astype(F32)
        e = np.broadcast_to(emb[b], (M,) + emb.shape[1:])
        if clamp and normalize:
            out[b] = mask_pooling(e, m)[:, 0, :]
        else:
            mm = bilinear_resize(m, emb.shape[2:]) if m.shape[2:] != emb.shape[2:] else m
            if clamp:
                mm = np.clip(mm, 0, 1)
            num = (e.astype(np.float64) * mm.astype(np.float64)).sum((2, 3))
            den = mm.astype(np.float64).sum((2, 3)) + 1e-8
            p = (num / den).astype(F32)
            out[b] = l2_normalize(p) if normalize else p
    return out


def region_query_similarity(regions: np.ndarray, queries: np.ndarray, bf16_operands: bool = True) -> np.ndarray:
    """S[q, r] = <Qn[q], Rn[r]> for unit rows, [N_q, N_r] float32.

    With ``bf16_operands`` both operands are first rounded to bf16 (what the tensor-core path
    consumes); products of two bf16 values are exact in fp32 and the K-sum is accumulated in
    float64 then rounded once, so this is the canonical, order-independent score used for the
    bit-identical top-k bar."""
    r = np.asarray(regions, dtype=F32)
    q = np.asarray(queries, dtype=F32)
    if bf16_operands:
        r, q = to_bf16(r), to_bf16(q)
    return (q.astype(np.float64) @ r.astype(np.float64).T).astype(F32)


def infonce_loss(regions, queries, targets, tau: float = 0.07, bf16_operands: bool = True) -> np.float32:
    """Region-level InfoNCE: mean_q CE(S[q,:]/tau, targets[q]).  tau is NOT defined by the
    reference (build choice, explicit argument)."""
    s = region_query_similarity(regions, queries, bf16_operands).astype(np.float64) / tau
    m = s.max(1, keepdims=True)
    lse = m[:, 0] + np.log(np.exp(s - m).sum(1))
    t = np.asarray(targets, dtype=np.int64)
    return F32((lse - s[np.arange(s.shape[0]), t]).mean())


def topk_retrieve(regions, queries, k: int, bf16_operands: bool = True):
    """Per query the k best regions under the total order (score desc, index asc).
    Returns (indices [N_q,k] int64, scores [N_q,k] float32)."""
    s = region_query_similarity(regions, queries, bf16_operands)
    order = np.argsort(-s.astype(np.float64), axis=1, kind="stable")[:, :k]
    return order.astype(np.int64), np.take_along_axis(s, order, 1)


def dice_loss(pred, target, smooth: float = 1.0) -> np.float32:
    """Soft dice on sigmoid(pred): mean_n 1 - (2*sum(p*t)+s)/(sum(p)+sum(t)+s).  Class N: the
    reference ships only the *name* ``bce_with_dice_loss`` in a stale .pyc; constants are a build
    choice."""
    p = sigmoid(np.asarray(pred, dtype=F32))
    t = np.asarray(target, dtype=F32).astype(np.float64)
    i = (p * t).sum((2, 3))
    return F32((1.0 - (2.0 * i + smooth) / (p.sum((2, 3)) + t.sum((2, 3)) + smooth)).mean())


def focal_loss(pred, target, alpha: float = 0.25, gamma: float = 2.0) -> np.float32:
    """Sigmoid focal loss, mean over all elements: alpha_t * (1-p_t)^gamma * bce.  Class N."""
    x = np.asarray(pred, dtype=F32).astype(np.float64)
    t = np.asarray(target, dtype=F32).astype(np.float64)
    p = sigmoid(x)
    bce = (1.0 - t) * x - log_sigmoid(x)
    p_t = p * t + (1 - p) * (1 - t)
    a_t = alpha * t + (1 - alpha) * (1 - t)
    return F32((a_t * (1 - p_t) ** gamma * bce).mean())
