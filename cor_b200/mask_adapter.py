"""Drop-in for the pooling modules of the reference's ``lib/support_model/mask_adapter.py``.

``MaskedPooling`` and ``MaskAdapterPooling`` keep the reference's constructor arguments, forward
signatures, output shapes and PARAMETER NAMES (``channel_clip_to_maskadapter.*``, ``get_mask_map.*``),
so checkpoints written by the reference load with ``strict=True`` (my_test.py:145).  The pooling
tails (mask_adapter.py:19-24 and :62-79) run in libcor_b200.so.

The learned half (``ChannelReduction`` :83-94 and ``GenerateMaskAdapterMap`` :97-179) runs through
:func:`adapter_maps` on CUDA: channels-last rows end to end, every 1x1 convolution / ``nn.Linear`` (channel reduction,
``fuse``, the 4x point-wise MLP of the three ConvNeXt blocks with GELU, layer scale and the skip connection in the GEMM
epilogue, the final 8-map head) forward AND backward on the tcgen05 GEMM (``cor_gemm_bf16``), every LayerNorm (+GELU) on
``cor_ln_rows_fwd/bwd``.  The reference adds the mask feature to a feature map REPEATED once per mask and then applies
``fuse`` (:155-163); a 1x1 convolution is linear, so ``fuse(x + m) = fuse(x) + fuse_w m``: the feature half is computed
once per IMAGE and the mask half by a [256 x 16] matrix composed from ``fuse`` and the last mask-downscaling conv --
the Q-fold repeat and the [B Q, 512, h, w] mask feature never exist.  The depth-wise 7x7 convolutions run on
``cor_dwconv7_cl`` (channels-last, forward, input gradient = flipped kernel, weight gradient).  What stays on cuDNN: the two
tiny stride-2 3x3 convolutions of ``mask_downscaling`` (1 -> 4 -> 16 channels, with their LayerNorm + GELU).
The module classes below keep the reference's eager definition as their CPU-importable form and parameter container.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops

__all__ = ["MaskedPooling", "MaskAdapterPooling", "masked_pool_tail", "softmax_map_pool_tail", "adapter_maps"]


def masked_pool_tail(clip_feature: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """mask_adapter.py:19-24: [B,C,h,w] x [B,1,H,W] -> [B,C]  (no clamp, no L2, eps 1e-8)."""
    out = ops.region_pool(clip_feature, mask, transform=ops.W_PLAIN, normalize=False, engine="stream").fg
    return out.reshape(clip_feature.shape[0], -1).to(torch.promote_types(clip_feature.dtype, mask.dtype))


class _AdapterTailFn(torch.autograd.Function):
    """softmax_P(logsigmoid(maps)) @ feat^T, mean over groups of G maps, as one batched tcgen05 GEMM per direction
    (csrc/adapter_tail.cu + csrc/gemm_umma.cu).  maps [B,R,h,w] f32, feat [B,C,h,w] -> [B, R/G, C] f32."""

    @staticmethod
    def forward(ctx, maps, feat, G):
        from . import _lib as L
        from .linear import cast_bf16, gemm
        dev = L.require_cuda(maps, feat)
        B, R, h, w = maps.shape
        C, P, Q = feat.shape[1], h * w, R // G
        Qp = (Q + 63) // 64 * 64
        m32 = maps.float().contiguous()
        wq = torch.zeros((B * Qp, P), dtype=torch.bfloat16, device=dev)
        den = torch.empty((B * R,), dtype=torch.float32, device=dev)
        ops._call("cor_adapter_tail_weights", dev, ops.ptr(m32), B, R, P, G, Qp, ops.ptr(wq), ops.ptr(den))
        f16 = cast_bf16(feat.detach().reshape(B * C, P)) if feat.dtype != torch.bfloat16 else feat.detach().reshape(B * C, P).contiguous()
        out = gemm(wq, f16, Q, C, P, batch=B, a_batch_rows=Qp, b_batch_rows=C)              # [B*Q, C] = Wq[b] feat[b]^T
        ctx.save_for_backward(m32, den, wq, f16)
        ctx.cfg = (B, R, P, C, G, Q, Qp, feat.dtype, maps.dtype, tuple(feat.shape), tuple(maps.shape))
        return out.view(B, Q, C)

    @staticmethod
    def backward(ctx, g):
        from .linear import cast_bf16, gemm
        m32, den, wq, f16 = ctx.saved_tensors
        B, R, P, C, G, Q, Qp, fdt, mdt, fshape, mshape = ctx.cfg
        dev = m32.device
        g16 = torch.empty((B * Qp, C), dtype=torch.bfloat16, device=dev)                    # zero rows: Qp is the K extent of d feat
        g32 = g.float().contiguous()
        ops._call("cor_cast_pad_rows_bf16", dev, ops.ptr(g32), B, Q, Qp, C, ops.ptr(g16))
        gf = gm = None
        if ctx.needs_input_grad[1]:
            # d feat[b] (C x P) = g[b]^T (C x Q) Wq[b] (Q x P): both operands stored [K = Qp][rows]
            gf = gemm(g16, wq, C, P, Qp, a_mn=True, b_mn=True, batch=B, a_batch_rows=Qp, b_batch_rows=Qp).view(fshape).to(fdt)
        if ctx.needs_input_grad[0]:
            # d Wq[b] (Q x P) = g[b] (Q x C) feat[b] (C x P): feat stored [K = C][P]
            gwq = gemm(g16, f16, Q, P, C, b_mn=True, batch=B, a_batch_rows=Qp, b_batch_rows=C)
            gm = torch.empty((B * R, P), dtype=torch.float32, device=dev)
            ops._call("cor_adapter_tail_bwd", dev, ops.ptr(m32), ops.ptr(den), ops.ptr(gwq), B, R, P, G, ops.ptr(gm))
            gm = gm.view(mshape).to(mdt)
        return gm, gf, None


def _tail_gemm_ok(maps: torch.Tensor, feat: torch.Tensor, G: int) -> bool:
    """Shapes the batched-GEMM tail serves: maps already at the feature resolution, TMA-legal row pitches (P, C multiples
    of 8) and C a whole number of 64-row K blocks (the MN-major feature operand of d Wq must not run into the next image)."""
    if not (maps.is_cuda and feat.is_cuda) or maps.dim() != 4 or feat.dim() != 4 or maps.shape[-2:] != feat.shape[-2:]:
        return False
    P, C = feat.shape[2] * feat.shape[3], feat.shape[1]
    return P % 8 == 0 and C % 64 == 0 and maps.shape[1] % G == 0 and G <= 64


def softmax_map_pool_tail(maps: torch.Tensor, clip_feature: torch.Tensor, num_output_maps: int) -> torch.Tensor:
    """mask_adapter.py:62-79: softmax_P(logsigmoid(maps)) @ feat^T, mean over groups of maps ->
    [B, N/num_output_maps, C].  softmax(logsigmoid(x)) == sigmoid(x) / sum sigmoid(x).  Tensor-core path: the mean and the
    normalisation fold into one bf16 weight row per mask, then one batched GEMM; other shapes take the exact fp32 streaming
    kernel."""
    if _tail_gemm_ok(maps, clip_feature, num_output_maps):
        return _AdapterTailFn.apply(maps, clip_feature, int(num_output_maps))
    return ops.region_pool(clip_feature, maps, transform=ops.W_SIGMOID, normalize=False, group=num_output_maps, eps=0.0,
                           engine="stream").fg


class MaskedPooling(nn.Module):
    def forward(self, clip_feature: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
        return masked_pool_tail(clip_feature, mask)


class LayerNorm(nn.Module):
    """LayerNorm over the channel axis for channels_last ([...,C]) or channels_first ([N,C,H,W])."""

    def __init__(self, normalized_shape, eps=1e-6, data_format="channels_last"):
        super().__init__()
        if data_format not in ("channels_last", "channels_first"):
            raise NotImplementedError(data_format)
        self.weight = nn.Parameter(torch.ones(normalized_shape))
        self.bias = nn.Parameter(torch.zeros(normalized_shape))
        self.eps, self.data_format, self.normalized_shape = eps, data_format, (normalized_shape,)

    def forward(self, x):
        if self.data_format == "channels_last":
            return F.layer_norm(x, self.normalized_shape, self.weight, self.bias, self.eps)
        mu = x.mean(1, keepdim=True)
        var = (x - mu).pow(2).mean(1, keepdim=True)
        return self.weight[:, None, None] * ((x - mu) / torch.sqrt(var + self.eps)) + self.bias[:, None, None]


class ChannelReduction(nn.Module):
    def __init__(self, in_channel, out_channel):
        super().__init__()
        self.conv = nn.Conv2d(in_channel, out_channel, 1)
        self.norm = LayerNorm(out_channel, data_format="channels_first")
        self.act = nn.GELU()

    def forward(self, x):
        return self.act(self.norm(self.conv(x)))


class ConvNextBlock(nn.Module):
    def __init__(self, dim, kernel_size=7, layer_scale_init_value=1e-6):
        super().__init__()
        self.dwconv = nn.Conv2d(dim, dim, kernel_size=kernel_size, padding=kernel_size // 2, groups=dim)
        self.norm = LayerNorm(dim, eps=1e-6)
        self.pwconv1 = nn.Linear(dim, 4 * dim)
        self.act = nn.GELU()
        self.pwconv2 = nn.Linear(4 * dim, dim)
        self.gamma = nn.Parameter(layer_scale_init_value * torch.ones(dim)) if layer_scale_init_value > 0 else None
        self.drop_path = nn.Identity()

    def forward(self, x):
        y = self.pwconv2(self.act(self.pwconv1(self.norm(self.dwconv(x).permute(0, 2, 3, 1)))))
        if self.gamma is not None:
            y = self.gamma * y
        return x + y.permute(0, 3, 1, 2)


class GenerateMaskAdapterMap(nn.Module):
    """Learned map generator (mask_adapter.py:97-179): per (image, mask) a stack of activation maps."""

    def __init__(self, clip_in_channel=768, mask_downscaling_mid_channel=16, mid_channel=768, num_output_maps=16):
        super().__init__()
        self.clip_in_channel = clip_in_channel
        self.fuse = nn.Conv2d(clip_in_channel, mid_channel, 1)
        self.cnext1, self.cnext2, self.cnext3 = ConvNextBlock(mid_channel), ConvNextBlock(mid_channel), ConvNextBlock(mid_channel)
        self.norm = LayerNorm(mid_channel)
        self.final = nn.Conv2d(mid_channel, num_output_maps, 1)
        m = mask_downscaling_mid_channel
        self.mask_downscaling = nn.Sequential(
            nn.Conv2d(1, m // 4, 3, stride=2, padding=1), LayerNorm(m // 4, data_format="channels_first"), nn.GELU(),
            nn.Conv2d(m // 4, m, 3, stride=2, padding=1), LayerNorm(m, data_format="channels_first"), nn.GELU(),
            nn.Conv2d(m, clip_in_channel, 1))

    def forward(self, clip_feature, masks):
        B, Q = masks.shape[:2]
        H, W = clip_feature.shape[-2:]
        m = masks.reshape(B * Q, 1, *masks.shape[2:]).float()
        m = self.mask_downscaling(F.interpolate(m, size=(H * 4, W * 4), mode="bilinear", align_corners=False))
        x = clip_feature.repeat_interleave(Q, dim=0) + m
        x = self.cnext3(self.cnext2(self.cnext1(self.fuse(x))))
        x = self.norm(x.permute(0, 2, 3, 1).contiguous()).permute(0, 3, 1, 2)
        x = self.final(x.contiguous())
        return x.reshape(B, Q * x.shape[1], H, W)


def _rows(x_nchw: torch.Tensor) -> torch.Tensor:
    """[N,C,h,w] -> channels-last rows [N*h*w, C] (a view when the tensor already is channels_last)."""
    n, c, h, w = x_nchw.shape
    return x_nchw.permute(0, 2, 3, 1).reshape(n * h * w, c)


def _convnext_rows(blk, y: torch.Tensor, n: int, h: int, w: int) -> torch.Tensor:
    """One ConvNeXt block (mask_adapter.py:182-223) on channels-last rows y [n*h*w, C] (f32)."""
    from .linear import dwconv7_ok, dwconv7_rows, linear, ln_rows
    c = y.shape[1]
    if blk.dwconv.kernel_size == (7, 7) and dwconv7_ok(c, h, w):
        t = dwconv7_rows(y, blk.dwconv.weight, blk.dwconv.bias, n, h, w)               # csrc/dwconv.cu, rows in, rows out
    else:
        img = y.view(n, h, w, c).permute(0, 3, 1, 2)                   # NCHW shape, channels_last strides: no copy
        t = _rows(F.conv2d(img, blk.dwconv.weight.float(), blk.dwconv.bias.float(), padding=blk.dwconv.padding, groups=c))
    t = ln_rows(t, blk.norm.weight, blk.norm.bias, blk.norm.eps, None, out_bf16=True)
    hdn = linear(t, blk.pwconv1.weight, blk.pwconv1.bias, "gelu", out_bf16=True)       # [rows, 4C] stays bf16 between the two GEMMs
    # gamma * pwconv2(.) + input: layer scale and skip connection in the GEMM epilogue
    return linear(hdn, blk.pwconv2.weight, blk.pwconv2.bias, None, None, None, blk.gamma, y)


def _md_fusable(md) -> bool:
    """mask_downscaling as the reference builds it (mask_adapter.py:128-142): conv, LayerNorm2d, GELU, conv, LayerNorm2d, GELU,
    conv with at most 32 channels inside."""
    return (len(md) == 7 and isinstance(md[1], LayerNorm) and isinstance(md[4], LayerNorm) and md[1].data_format == "channels_first"
            and md[4].data_format == "channels_first" and isinstance(md[2], nn.GELU) and isinstance(md[5], nn.GELU)
            and getattr(md[2], "approximate", "none") == "none" and getattr(md[5], "approximate", "none") == "none"
            and md[1].weight.numel() <= 32 and md[4].weight.numel() <= 32)


def adapter_maps(adapter, clip_feature: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """``get_mask_map(channel_clip_to_maskadapter(clip_feature), mask)`` (mask_adapter.py:59-60) with ``adapter``'s own
    parameters: clip_feature [B,C,h,w], mask [B,Q,h,w] (already at the feature resolution) -> maps [B, Q*num_output_maps, h, w]."""
    from .linear import linear, ln_rows
    cr, gm = adapter.channel_clip_to_maskadapter, adapter.get_mask_map
    B, C, h, w = clip_feature.shape
    Q = mask.shape[1]
    P = h * w
    # ChannelReduction (:83-94): 1x1 conv -> LayerNorm(channels_first) -> GELU, on channels-last rows
    x = linear(_rows(clip_feature.float()), cr.conv.weight.view(cr.conv.out_channels, C), cr.conv.bias)
    x = ln_rows(x, cr.norm.weight, cr.norm.bias, cr.norm.eps, "gelu")                                     # [B*P, 512]
    # mask branch (:157-158): x4 bilinear up (our resample kernel: the mask carries no gradient), two stride-2 3x3 convs
    # (cuDNN: a few hundred kFLOP per mask) each followed by LayerNorm2d + GELU in ONE launch of ours (the reference's
    # channels-first LayerNorm is ~8 element-wise launches forward, ~20 backward) -> 16 channels at h x w
    md = gm.mask_downscaling
    if mask.requires_grad:
        m = F.interpolate(mask.reshape(B * Q, 1, h, w).float(), size=(4 * h, 4 * w), mode="bilinear", align_corners=False)
    else:
        m = ops.mask_prep(mask.reshape(B * Q, h, w).float().contiguous(), (4 * h, 4 * w))[0].view(B * Q, 1, 4 * h, 4 * w)
    if _md_fusable(md):
        from .linear import ln_channels_first
        t = ln_channels_first(md[0](m), md[1].weight, md[1].bias, md[1].eps, "gelu")
        m16 = ln_channels_first(md[3](t), md[4].weight, md[4].bias, md[4].eps, "gelu")                    # [B*Q, 16, h, w]
    else:
        m16 = md[5](md[4](md[3](md[2](md[1](md[0](m))))))
    # fuse(x + conv(m16)) = fuse(x) + (fuse_w conv_w) m16 + fuse_w conv_b + fuse_b   (:161-163; 1x1 convs are linear)
    mid = gm.fuse.out_channels
    fuse_w = gm.fuse.weight.view(mid, -1).float()
    conv_w = md[6].weight.view(md[6].out_channels, -1).float()
    w_comp = fuse_w @ conv_w                                                                              # [256, 16]
    b_comp = fuse_w @ md[6].bias.float() + gm.fuse.bias.float()
    fx = linear(x, gm.fuse.weight.view(mid, -1), None)                                                    # once per image: [B*P, 256]
    fm = linear(_rows(m16), w_comp, b_comp)                                                               # [B*Q*P, 256]
    y = (fm.view(B, Q, P, mid) + fx.view(B, 1, P, mid)).reshape(B * Q * P, mid)
    for blk in (gm.cnext1, gm.cnext2, gm.cnext3):
        y = _convnext_rows(blk, y, B * Q, h, w)
    y = ln_rows(y, gm.norm.weight, gm.norm.bias, gm.norm.eps, None, out_bf16=True)
    nmaps = gm.final.out_channels
    o = linear(y, gm.final.weight.view(nmaps, mid), gm.final.bias)                                        # [B*Q*P, 8]
    return o.view(B, Q, P, nmaps).permute(0, 1, 3, 2).reshape(B, Q * nmaps, h, w)


def _adapter_maps_ok(clip_feature: torch.Tensor) -> bool:
    """Shapes the GEMM path takes: CUDA, channel counts that give 16-byte TMA row pitches (all of the reference's)."""
    return clip_feature.is_cuda and clip_feature.shape[1] % 8 == 0


class MaskAdapterPooling(nn.Module):
    def __init__(self, x_in_channel=1152, mask_adatpet_network_in_channel=256, mask_downscaling_mid_channel=16,
                 mask_adatpet_network_mid_channel=256, num_output_maps=16):
        super().__init__()
        self.channel_clip_to_maskadapter = ChannelReduction(x_in_channel, mask_adatpet_network_in_channel)
        self.get_mask_map = GenerateMaskAdapterMap(mask_adatpet_network_in_channel, mask_downscaling_mid_channel,
                                                   mask_adatpet_network_mid_channel, num_output_maps)
        self.num_output_maps = num_output_maps

    def forward(self, clip_feature: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
        if mask.shape[-2:] != clip_feature.shape[-2:]:
            mask = F.interpolate(mask, size=clip_feature.shape[-2:], mode="bilinear", align_corners=False)
        if _adapter_maps_ok(clip_feature):
            maps = adapter_maps(self, clip_feature, mask)
        else:   # CPU import / odd channel counts: the eager definition (the pooling tail below still refuses CPU tensors)
            maps = self.get_mask_map(self.channel_clip_to_maskadapter(clip_feature), mask)
        return softmax_map_pool_tail(maps, clip_feature, self.num_output_maps)
