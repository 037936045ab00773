"""Deterministic weight recipes for parity tests.  TEST INFRASTRUCTURE ONLY (see oracle/fen_oracle.py).

The reference initialises with Kaiming-normal(fan_out, relu), zero biases, PReLU slope 0.25 and a
ZERO conv_last (src/models/custom.py:130-145), which makes forward(x) == clamp(bicubic(x)) exactly:
a parity test on the literal init is vacuous.  The tiers below fix that:

  T0  literal init statistics (conv_last = 0)        -> checks skip + clamp + plumbing
  T1  conv_last ~ N(0, 1e-3), conv biases ~ N(0, 0.01), PReLU slopes ~ U(0.05, 0.45)   (the bar)
  T2  as T1 with conv_last ~ N(0, 1e-2)              (stress, report only)

Weights come from numpy's PCG64 stream (platform independent), never from replaying torch's RNG, so
the same state_dict can be rebuilt on the GPU box, loaded into the real reference here
(tests/golden/make_golden.py) and into the product module.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict

import numpy as np
import torch


def state_dict_schema(num_groups=6, blocks_per_group=10, num_channels=64, reduction_ratio=4,
                      in_channels=3, out_channels=3, num_stages=2) -> "OrderedDict[str, tuple]":
    """Key -> shape, in the reference module's registration order."""
    C = num_channels
    R = max(C // reduction_ratio, 8)
    s: "OrderedDict[str, tuple]" = OrderedDict()
    s["conv_first.weight"] = (C, in_channels, 3, 3)
    s["conv_first.bias"] = (C,)
    for g in range(num_groups):
        for b in range(blocks_per_group):
            p = f"residual_groups.{g}.blocks.{b}"
            s[p + ".conv1.weight"] = (C, C, 3, 3)
            s[p + ".conv1.bias"] = (C,)
            s[p + ".prelu.weight"] = (C,)
            s[p + ".conv2.weight"] = (C, C, 3, 3)
            s[p + ".conv2.bias"] = (C,)
            s[p + ".channel_attention.fc.0.weight"] = (R, C)
            s[p + ".channel_attention.fc.2.weight"] = (C, R)
        s[f"residual_groups.{g}.conv.weight"] = (C, C, 3, 3)
        s[f"residual_groups.{g}.conv.bias"] = (C,)
    s["conv_after_body.weight"] = (C, C, 3, 3)
    s["conv_after_body.bias"] = (C,)
    for st in range(num_stages):
        s[f"upsample.stages.{st}.conv.weight"] = (4 * C, C, 3, 3)
        s[f"upsample.stages.{st}.conv.bias"] = (4 * C,)
        s[f"upsample.stages.{st}.prelu.weight"] = (C,)
    s["conv_last.weight"] = (out_channels, C, 3, 3)
    s["conv_last.bias"] = (out_channels,)
    return s


def make_state_dict(seed: int = 0, tier: str = "T1", **cfg) -> Dict[str, torch.Tensor]:
    """Build a full fp32 state_dict for the given config and parity tier."""
    if tier not in ("T0", "T1", "T2"):
        raise ValueError(tier)
    rng = np.random.default_rng(seed)
    out: Dict[str, torch.Tensor] = OrderedDict()
    for key, shape in state_dict_schema(**cfg).items():
        if key.endswith("prelu.weight"):
            v = np.full(shape, 0.25, np.float32) if tier == "T0" else rng.uniform(0.05, 0.45, shape)
        elif key.startswith("conv_last"):
            sigma = {"T0": 0.0, "T1": 1e-3, "T2": 1e-2}[tier]
            v = rng.normal(0.0, 1.0, shape) * sigma
        elif key.endswith(".bias"):
            v = np.zeros(shape) if tier == "T0" else rng.normal(0.0, 0.01, shape)
        else:  # conv / linear weight: Kaiming normal, mode=fan_out, gain sqrt(2)
            fan_out = shape[0] * int(np.prod(shape[2:])) if len(shape) == 4 else shape[0]
            v = rng.normal(0.0, 1.0, shape) * np.sqrt(2.0 / fan_out)
        out[key] = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))
    return out
