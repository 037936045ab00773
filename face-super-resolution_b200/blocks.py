"""Parameter containers with the reference's sub-module tree (src/models/blocks.py:44-263).

The reference's callers reach inside the model (`group.blocks[b].channel_attention.fc`,
`named_modules()` filtered on 'residual_groups'; custom.py:209-217, explainability.py:119-139), and
`state_dict()` keys are derived from this tree, so the same named modules exist here.  They only
HOLD parameters: the arithmetic runs in the fused CUDA kernels driven by FaceEnhanceNet.forward, so
calling a sub-module on its own raises instead of silently running a PyTorch fallback (forward hooks
on sub-modules therefore never fire; use FaceEnhanceNet.get_attention_maps for the SE weights).
"""
from __future__ import annotations

import torch
import torch.nn as nn


class _FusedOnly(nn.Module):
    def forward(self, *args, **kwargs):  # noqa: D401
        raise RuntimeError(
            f"{type(self).__name__} is a parameter container of the fused B200 path; it is evaluated inside "
            "FaceEnhanceNet.forward by CUDA kernels and has no standalone (PyTorch/CPU) forward.")


class ChannelAttention(_FusedOnly):
    """blocks.py:44-92.  fc.0 = Linear(C -> max(C // r, 8), no bias), fc.2 = Linear(back, no bias)."""

    def __init__(self, num_channels: int, reduction_ratio: int = 4):
        super().__init__()
        self.num_channels = num_channels
        self.reduction_ratio = reduction_ratio
        hidden = max(num_channels // reduction_ratio, 8)
        self.global_pool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Sequential(
            nn.Linear(num_channels, hidden, bias=False), nn.ReLU(inplace=True),
            nn.Linear(hidden, num_channels, bias=False), nn.Sigmoid())


class RCAB(_FusedOnly):
    """blocks.py:95-153.  conv1 -> PReLU(C) -> conv2 -> channel attention, * res_scale + x."""

    def __init__(self, num_channels: int = 64, kernel_size: int = 3, reduction_ratio: int = 4,
                 bias: bool = True, res_scale: float = 0.2):
        super().__init__()
        self.res_scale = res_scale
        pad = kernel_size // 2
        self.conv1 = nn.Conv2d(num_channels, num_channels, kernel_size, padding=pad, bias=bias)
        self.prelu = nn.PReLU(num_channels)
        self.conv2 = nn.Conv2d(num_channels, num_channels, kernel_size, padding=pad, bias=bias)
        self.channel_attention = ChannelAttention(num_channels, reduction_ratio)


class ResidualGroup(_FusedOnly):
    """blocks.py:156-189.  `blocks` RCABs, then conv, plus the group input."""

    def __init__(self, num_channels: int = 64, num_blocks: int = 4, kernel_size: int = 3,
                 reduction_ratio: int = 4, res_scale: float = 0.2):
        super().__init__()
        self.blocks = nn.Sequential(*[
            RCAB(num_channels, kernel_size, reduction_ratio, res_scale=res_scale) for _ in range(num_blocks)])
        self.conv = nn.Conv2d(num_channels, num_channels, kernel_size, padding=kernel_size // 2)


class PixelShuffleUpsample(_FusedOnly):
    """blocks.py:192-227.  conv C -> 4C, PixelShuffle(2), PReLU(C)."""

    def __init__(self, in_channels: int, scale_factor: int = 2):
        super().__init__()
        self.scale_factor = scale_factor
        self.conv = nn.Conv2d(in_channels, in_channels * scale_factor ** 2, kernel_size=3, padding=1)
        self.pixel_shuffle = nn.PixelShuffle(scale_factor)
        self.prelu = nn.PReLU(in_channels)


class UpsampleModule(_FusedOnly):
    """blocks.py:230-263.  log2(scale) PixelShuffleUpsample stages."""

    def __init__(self, num_channels: int = 64, scale_factor: int = 4):
        super().__init__()
        self.scale_factor = scale_factor
        n, s = 0, scale_factor
        while s > 1:
            s //= 2
            n += 1
        self.stages = nn.Sequential(*[PixelShuffleUpsample(num_channels, 2) for _ in range(n)])
