"""CPU oracle for the bicubic low-resolution generator.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product path (face-super-resolution_b200/) never does.

What it restates
----------------
The reference produces its LR images with one third-party call (opencv-python >= 4.8, unpinned in
requirements.txt:13; cv2 4.13.0 is what this image ships and what the goldens were made with):

    src/data/dataset.py:292-296     lr = cv2.resize(hr, (w // 4, h // 4), interpolation=cv2.INTER_CUBIC)
    src/data/prepare_data.py:36-39  create_lr_image(hr, lr_size=64, method='bicubic')
    src/data/transforms.py:260-279  to_tensor: HWC uint8 -> CHW float32 / 255.0

cv2's source is not under /root/reference, so this file restates its published algorithm for the
exact /4 case: the source coordinate of output x is 4x + 1.5, so the four cubic taps (A = -0.75)
always sit at fraction 0.5 on pixels 4x .. 4x+3 (never out of range, no border handling) with
weights [-3, 19, 19, -3] / 32.  cv2 evaluates the separable filter in fixed point and rounds the
final value half-to-even, which makes the whole operation the integer formula below.

Parity pin: tests/golden/lr_*.npz hold cv2.resize outputs made in this image by
tests/golden/make_golden.py; tests/test_oracle_lr.py checks this restatement against them
bit-for-bit (and against cv2 live when cv2 is importable).
"""
from __future__ import annotations

import numpy as np

_TAPS = np.array([-3, 19, 19, -3], dtype=np.int32)


def lr_from_hr_u8(hr: np.ndarray) -> np.ndarray:
    """Integer restatement of ``cv2.resize(hr, (W//4, H//4), interpolation=cv2.INTER_CUBIC)``.

    hr: uint8 array [..., H, W, C] with H, W multiples of 4.  Returns uint8 [..., H//4, W//4, C].

        u  = sum_{i,j in 0..3} a_i a_j hr[4y+i, 4x+j, c],   a = [-3, 19, 19, -3]
        lr = clamp(round_half_even(u / 1024), 0, 255)
    """
    if hr.dtype != np.uint8:
        raise TypeError("hr must be uint8")
    *lead, H, W, C = hr.shape
    if H % 4 or W % 4:
        raise ValueError("H and W must be multiples of 4")
    x = hr.reshape(-1, H // 4, 4, W // 4, 4, C).astype(np.int32)
    # separable: rows (i) then columns (j); all exact in int32 (|u| <= 377400)
    u = np.einsum("nyixjc,i,j->nyxc", x, _TAPS, _TAPS, optimize=True).astype(np.int32)
    q = (u + 511 + ((u >> 10) & 1)) >> 10  # round half to even of u / 1024 (arithmetic shift)
    return np.clip(q, 0, 255).astype(np.uint8).reshape(*lead, H // 4, W // 4, C)


def lr_from_hr_u8_loops(hr: np.ndarray) -> np.ndarray:
    """Same formula written as plain loops (small inputs only) - an independent second statement."""
    H, W, C = hr.shape
    out = np.zeros((H // 4, W // 4, C), dtype=np.uint8)
    a = (-3, 19, 19, -3)
    for y in range(H // 4):
        for x in range(W // 4):
            for c in range(C):
                u = 0
                for i in range(4):
                    for j in range(4):
                        u += a[i] * a[j] * int(hr[4 * y + i, 4 * x + j, c])
                # Python's round() is half-to-even; u/1024 is exact in binary floating point
                out[y, x, c] = min(255, max(0, round(u / 1024)))
    return out


def to_tensor_chw(lr_u8: np.ndarray) -> np.ndarray:
    """``to_tensor`` of src/data/transforms.py:260-279: HWC uint8 -> CHW float32 divided by 255."""
    return (np.moveaxis(lr_u8, -1, -3).astype(np.float32) / np.float32(255.0)).astype(np.float32)


# ---------------------------------------------------------------------------------------------------
# Float LR generator of the trainer and of the evaluation scripts.
# The reference code IS a single PyTorch call, so the oracle executes exactly that call on the CPU:
#   src/training/trainer.py:416-421  lr = F.interpolate(hr, scale_factor=0.25, mode='bicubic', align_corners=False)
#   scripts/test_model.py:139-156    generate_lr: same taps, then np.clip(lr * 255, 0, 255).astype(np.uint8)
#   scripts/test_model.py:176-190    to_numpy: np.clip(sr * 255, 0, 255).astype(np.uint8), CHW -> HWC, RGB -> BGR
# lr_from_hr_float_taps restates the algorithm (16 taps w_i w_j, w = [-3, 19, 19, -3] / 32) and
# tests/test_oracle_lr.py pins it against the PyTorch call to 2.4e-7 (fp32 summation order).
def lr_from_hr_float(hr):
    """hr: float32 torch tensor [B,C,H,W] -> [B,C,H/4,W/4], the trainer's exact call."""
    import torch.nn.functional as F
    return F.interpolate(hr.detach().float().cpu(), scale_factor=0.25, mode="bicubic", align_corners=False)


def lr_from_hr_float_taps(hr: np.ndarray) -> np.ndarray:
    """The same operation as explicit taps (numpy, float32): rows first, then columns."""
    w = np.array([-0.09375, 0.59375, 0.59375, -0.09375], dtype=np.float32)
    hr = np.asarray(hr, dtype=np.float32)
    rows = [sum(hr[..., i::4, j::4] * w[j] for j in range(4)).astype(np.float32) for i in range(4)]
    return sum(rows[i] * w[i] for i in range(4)).astype(np.float32)


def quantize_u8_hwc(x_chw: np.ndarray, bgr: bool = False) -> np.ndarray:
    """np.clip(x * 255, 0, 255).astype(np.uint8) on [..., C, H, W] float32 -> [..., H, W, C] (scripts' to_numpy /
    generate_lr; astype truncates), channel order reversed when bgr (cv2.COLOR_RGB2BGR)."""
    q = np.clip(np.asarray(x_chw, dtype=np.float32) * np.float32(255), 0, 255).astype(np.uint8)
    q = np.moveaxis(q, -3, -1)
    return q[..., ::-1].copy() if bgr else q
