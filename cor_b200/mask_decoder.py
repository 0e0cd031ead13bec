"""Mask-logit producer of the SAM decoder on the B200 path.

The reference ends ``MaskDecoder.predict_masks`` with the hypernetwork product

    masks = (hyper_in @ upscaled_embedding.view(b, c, h * w)).view(b, -1, h, w)      lib/sam_model/mask_decoder.py:135-137

for all four mask tokens and ``forward`` then keeps ``masks[:, 0:1]`` (``multimask_output=False``, :97-102) -- the
logits ``wbce_with_wiou_loss`` reads straight back (utils/trainer_v3_g.py:67-68).  :func:`hyper_mask_logits` produces
only the consumed token(s), in one pass over the upscaled embedding, forward and backward, on the kernels of
``csrc/hyper_logits.cu``; :func:`logits_and_seg_loss` chains it into the segmentation-loss kernel so the producer
writes bf16 logits once and the loss reads them once.  ``hooks.install()`` can rebind ``MaskDecoder.predict_masks``
to :func:`predict_masks` (same arguments and return values as the reference method).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib as L
from . import ops
from ._lib import CorError

__all__ = ["hyper_mask_logits", "logits_and_seg_loss", "predict_masks"]


class _HyperLogitsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hyper_in, upscaled, t0, T, out_dtype):
        dev = L.require_cuda(hyper_in, upscaled)
        if hyper_in.dim() != 3 or upscaled.dim() != 4 or hyper_in.shape[0] != upscaled.shape[0] or hyper_in.shape[2] != upscaled.shape[1]:
            raise CorError(f"hyper_mask_logits: hyper_in {tuple(hyper_in.shape)} must be [B,T,C] and upscaled {tuple(upscaled.shape)} [B,C,H,W]")
        B, T_all, Cc = hyper_in.shape
        H, W = upscaled.shape[2:]
        P = H * W
        h = hyper_in.float().contiguous()
        up = ops._as_supported_float(upscaled)
        out = torch.empty((B, T, H, W), dtype=out_dtype, device=dev)
        ops._call("cor_hyper_logits_fwd", dev, ops.ptr(h), ops.ptr(up), L.dtype_code(up), ops.ptr(out), L.dtype_code(out), B, T_all, int(t0),
                  int(T), Cc, ops._ll(P))
        ctx.save_for_backward(h, up)
        ctx.cfg = (int(t0), int(T), hyper_in.dtype, upscaled.dtype, upscaled.requires_grad)
        return out

    @staticmethod
    def backward(ctx, g):
        h, up = ctx.saved_tensors
        t0, T, h_dtype, up_dtype, need_up = ctx.cfg
        dev = h.device
        B, T_all, Cc = h.shape
        P = up.shape[2] * up.shape[3]
        gc = ops._as_supported_float(g)
        d_up = torch.empty_like(up) if need_up else None
        d_h = torch.empty_like(h)
        work = ops._work(L.load().cor_hyper_logits_work_bytes(B, T, Cc, P), dev)
        ops._call("cor_hyper_logits_bwd", dev, ops.ptr(h), ops.ptr(up), L.dtype_code(up), ops.ptr(gc), L.dtype_code(gc), ops.ptr(d_up),
                  ops.ptr(d_h), B, T_all, t0, T, Cc, ops._ll(P), ops.ptr(work))
        return d_h.to(h_dtype), (d_up.to(up_dtype) if d_up is not None else None), None, None, None


def hyper_mask_logits(hyper_in: torch.Tensor, upscaled_embedding: torch.Tensor, tokens: Tuple[int, int] = (0, 1),
                      out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """``(hyper_in @ upscaled.view(b, c, h*w)).view(b, -1, h, w)[:, t0:t1]`` (mask_decoder.py:135-137 + the slice at
    :97-102): hyper_in [B,T,C] x upscaled [B,C,H,W] -> [B, t1-t0, H, W].  ``tokens=(0, 4)`` gives all of SAM's masks."""
    t0, t1 = tokens
    if out_dtype is None:
        out_dtype = upscaled_embedding.dtype if upscaled_embedding.dtype in (torch.float32, torch.bfloat16) else torch.float32
    return _HyperLogitsFn.apply(hyper_in, upscaled_embedding, int(t0), int(t1 - t0), out_dtype)


def logits_and_seg_loss(hyper_in: torch.Tensor, upscaled_embedding: torch.Tensor, query_mask: torch.Tensor, w1: float = 1.0,
                        w2: float = 1.0):
    """Producer -> loss without the detour through four fp32 mask planes: bf16 logits of token 0 written once, read once by
    the segmentation-loss kernel (mask resample fused, utils/trainer_v3_g.py:67-68).  Returns (loss, logits [B,1,H,W])."""
    logits = hyper_mask_logits(hyper_in, upscaled_embedding, (0, 1), torch.bfloat16)
    return ops.seg_loss(logits, query_mask, w1, w2), logits


def predict_masks(self, image_embeddings, image_pe, sparse_prompt_embeddings, dense_prompt_embeddings):
    """Drop-in for ``MaskDecoder.predict_masks`` (lib/sam_model/mask_decoder.py:107-142): the module's own parameters and
    sub-modules do everything up to the hypernetwork product, which runs on :func:`hyper_mask_logits` (all tokens, so
    ``forward``'s slicing keeps working unchanged)."""
    B = image_embeddings.size(0)
    output_tokens = torch.cat([self.iou_token.weight, self.mask_tokens.weight], dim=0).unsqueeze(0).expand(B, -1, -1)
    tokens = torch.cat((output_tokens, sparse_prompt_embeddings), dim=1)
    src = image_embeddings + dense_prompt_embeddings
    pos_src = image_pe.expand(B, -1, -1, -1)
    hs, src = self.transformer(src, pos_src, tokens)
    iou_token_out = hs[:, 0, :]
    mask_tokens_out = hs[:, 1:(1 + self.num_mask_tokens), :]
    src = src.transpose(1, 2).view(B, -1, 64, 64)
    upscaled_embedding = self.output_upscaling(src)
    hyper_in = torch.stack([mlp(mask_tokens_out[:, i, :]) for i, mlp in enumerate(self.output_hypernetworks_mlps)], dim=1)
    masks = hyper_mask_logits(hyper_in, upscaled_embedding, (0, self.num_mask_tokens), torch.float32)
    return masks, self.iou_prediction_head(iou_token_out), src
