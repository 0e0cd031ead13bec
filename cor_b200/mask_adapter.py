"""Drop-in for the pooling modules of the reference's ``lib/support_model/mask_adapter.py``.

``MaskedPooling`` and ``MaskAdapterPooling`` keep the reference's constructor arguments, forward
signatures, output shapes and PARAMETER NAMES (``channel_clip_to_maskadapter.*``, ``get_mask_map.*``),
so checkpoints written by the reference load with ``strict=True`` (my_test.py:145).  The pooling
tails (mask_adapter.py:19-24 and :62-79) run in libcor_b200.so; the learned map generator is dense
conv / linear work and stays on cuDNN / cuBLAS.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops

__all__ = ["MaskedPooling", "MaskAdapterPooling", "masked_pool_tail", "softmax_map_pool_tail"]


def masked_pool_tail(clip_feature: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """mask_adapter.py:19-24: [B,C,h,w] x [B,1,H,W] -> [B,C]  (no clamp, no L2, eps 1e-8)."""
    out = ops.region_pool(clip_feature, mask, transform=ops.W_PLAIN, normalize=False, engine="stream").fg
    return out.reshape(clip_feature.shape[0], -1).to(torch.promote_types(clip_feature.dtype, mask.dtype))


def softmax_map_pool_tail(maps: torch.Tensor, clip_feature: torch.Tensor, num_output_maps: int) -> torch.Tensor:
    """mask_adapter.py:62-79: softmax_P(logsigmoid(maps)) @ feat^T, mean over groups of maps ->
    [B, N/num_output_maps, C].  softmax(logsigmoid(x)) == sigmoid(x) / sum sigmoid(x)."""
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
        maps = self.get_mask_map(self.channel_clip_to_maskadapter(clip_feature), mask)
        return softmax_map_pool_tail(maps, clip_feature, self.num_output_maps)
