"""CPU fp32 oracle for the FaceEnhanceNet forward path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product path (face-super-resolution_b200/) never does and fails loudly when its
CUDA library is missing.

A functional restatement, in plain fp32 PyTorch ops on the CPU, of

    src/models/custom.py:147-190   FaceEnhanceNet.forward
    src/models/blocks.py:75-92     ChannelAttention.forward
    src/models/blocks.py:135-153   RCAB.forward
    src/models/blocks.py:185-189   ResidualGroup.forward
    src/models/blocks.py:223-227   PixelShuffleUpsample.forward

driven directly by a state_dict with the reference's key schema, so it
needs neither /root/reference nor the product package.  Floating point: the product is compared to
this oracle with the tolerance BASELINE.json states (PSNR >= 50 dB, max-abs <= 2e-2 on [0,1]).

Parity pin: tests/golden/fen_*.npz hold outputs of the real reference module (imported from
/root/reference by tests/golden/make_golden.py) for weights made by oracle/weights.py;
tests/test_oracle_fen.py checks this restatement against them (<= 1e-5 max-abs, fp32 reassociation
only).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn.functional as F


def _conv(sd: Dict[str, torch.Tensor], name: str, x: torch.Tensor) -> torch.Tensor:
    return F.conv2d(x, sd[name + ".weight"], sd[name + ".bias"], padding=1)


def count_groups_blocks(sd: Dict[str, torch.Tensor]):
    """Recover (num_groups, blocks_per_group) from the keys, as scripts/test_model.py:35-79 does."""
    groups, blocks = set(), set()
    for k in sd:
        p = k.split(".")
        if p[0] == "residual_groups":
            groups.add(int(p[1]))
            if p[2] == "blocks":
                blocks.add(int(p[3]))
    return len(groups), len(blocks)


def channel_attention(sd, prefix: str, o: torch.Tensor) -> torch.Tensor:
    """blocks.py:86-89: mean over H,W -> Linear(no bias) -> ReLU -> Linear(no bias) -> Sigmoid."""
    y = o.mean(dim=(2, 3))
    y = F.relu(F.linear(y, sd[prefix + ".fc.0.weight"]))
    return torch.sigmoid(F.linear(y, sd[prefix + ".fc.2.weight"]))


def fen_forward(
    sd: Dict[str, torch.Tensor],
    x: torch.Tensor,
    training: bool = False,
    res_scale: float = 0.2,
    scale_factor: int = 4,
    taps: Optional[Dict[str, torch.Tensor]] = None,
) -> torch.Tensor:
    """x: [B,3,H,W] fp32 in [0,1] -> [B,3,4H,4W] fp32.  `taps`, if given, receives intermediate
    tensors: 'conv_first', 'group{g}', 'body', 'up{s}', 'se' ([B, n_rcab, C] attention scales)."""
    sd = {k: v.detach().to(torch.float32).cpu() for k, v in sd.items()}
    x = x.detach().to(torch.float32).cpu()
    return _forward(sd, x, training, res_scale, scale_factor, taps)


def _forward(sd, x, training, res_scale, scale_factor, taps):
    n_groups, n_blocks = count_groups_blocks(sd)
    bicubic = F.interpolate(x, scale_factor=scale_factor, mode="bicubic", align_corners=False)
    feat = _conv(sd, "conv_first", x)
    long_skip = feat
    if taps is not None:
        taps["conv_first"] = feat
    se_all: List[torch.Tensor] = []
    for g in range(n_groups):
        group_in = feat
        for b in range(n_blocks):
            p = f"residual_groups.{g}.blocks.{b}"
            h = F.prelu(_conv(sd, p + ".conv1", feat), sd[p + ".prelu.weight"])
            o = _conv(sd, p + ".conv2", h)
            s = channel_attention(sd, p + ".channel_attention", o)
            se_all.append(s)
            feat = (o * s[:, :, None, None]) * res_scale + feat
        feat = _conv(sd, f"residual_groups.{g}.conv", feat) + group_in
        if taps is not None:
            taps[f"group{g}"] = feat
    feat = _conv(sd, "conv_after_body", feat) + long_skip
    if taps is not None:
        taps["body"] = feat
        taps["se"] = torch.stack(se_all, dim=1) if se_all else torch.zeros(x.shape[0], 0, feat.shape[1])
    s_idx = 0
    while f"upsample.stages.{s_idx}.conv.weight" in sd:
        p = f"upsample.stages.{s_idx}"
        feat = F.prelu(F.pixel_shuffle(_conv(sd, p + ".conv", feat), 2), sd[p + ".prelu.weight"])
        if taps is not None:
            taps[f"up{s_idx}"] = feat
        s_idx += 1
    out = _conv(sd, "conv_last", feat) + bicubic
    if not training:
        out = torch.clamp(out, 0.0, 1.0)
    return out


def fen_backward(sd: Dict[str, torch.Tensor], x: torch.Tensor, dout: torch.Tensor, res_scale: float = 0.2,
                 scale_factor: int = 4):
    """What `sr = model(lr); sr.backward(dout)` leaves in .grad of every state_dict tensor in train() mode
    (the network side of loss.backward() in src/training/trainer.py:462-488): fp32 autograd over the
    restatement above.  Returns (sr, {key: grad})."""
    leaves = {k: v.detach().to(torch.float32).cpu().clone().requires_grad_(True) for k, v in sd.items()}
    with torch.enable_grad():
        sr = _forward(leaves, x.detach().to(torch.float32).cpu(), True, res_scale, scale_factor, None)
        sr.backward(dout.detach().to(torch.float32).cpu())
    return sr.detach(), {k: v.grad for k, v in leaves.items()}


def l1_grad(sr: torch.Tensor, hr: torch.Tensor) -> torch.Tensor:
    """d mean|sr - hr| / d sr (nn.L1Loss, src/losses/combined.py:38-47)."""
    return torch.sign(sr - hr) / sr.numel()


def bicubic_x4_weights():
    """The 4 phase filters of F.interpolate(scale_factor=4, 'bicubic', align_corners=False), A=-0.75,
    times 2048 (exact integers).  Output d = 4q + r reads input q+off-1 .. q+off+2 (index-clamped),
    off = -1,-1,0,0 for r = 0..3.  (custom.py:158-161)"""
    w = [[-135, 873, 1535, -225], [-21, 235, 1981, -147], [-147, 1981, 235, -21], [-225, 1535, 873, -135]]
    off = [-1, -1, 0, 0]
    return w, off


def psnr(a: torch.Tensor, b: torch.Tensor, max_val: float = 1.0) -> float:
    """src/evaluation/metrics.py:17-34 formula: 10 log10(max^2 / mse)."""
    mse = torch.mean((a.double() - b.double()) ** 2).item()
    if mse == 0:
        return float("inf")
    import math

    return 10.0 * math.log10(max_val * max_val / mse)
