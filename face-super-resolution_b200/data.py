"""GPU replacement for the bicubic LR generation of the reference's src/data
(`create_lr_image`, prepare_data.py:23-59; the inline `cv2.resize` of dataset.py:292-296;
`to_tensor`, transforms.py:260-279).  Integer kernel, bit-exact against cv2.INTER_CUBIC."""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib


def lr_from_hr(hr: torch.Tensor, want_u8: bool = True, want_f32: bool = True
               ) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """Batched LR generation on the GPU.

    hr: uint8 CUDA tensor [B,H,W,C] (HWC, as cv2 / the dataset hold images), H, W multiples of 4.
    Returns (lr_u8 [B,H/4,W/4,C] uint8 HWC, lr_f32 [B,C,H/4,W/4] float32 = to_tensor(lr_u8))."""
    if hr.dtype != torch.uint8 or hr.dim() != 4:
        raise TypeError("hr must be a uint8 tensor [B,H,W,C]")
    if not hr.is_cuda:
        raise RuntimeError("lr_from_hr needs a CUDA tensor: there is no CPU fallback")
    if not (want_u8 or want_f32):
        raise ValueError("nothing requested")
    hr = hr.contiguous()
    B, H, W, Cn = hr.shape
    if H % 4 or W % 4:
        raise ValueError("H and W must be multiples of 4")
    lib = _lib.load()
    with torch.cuda.device(hr.device):
        u8 = torch.empty((B, H // 4, W // 4, Cn), dtype=torch.uint8, device=hr.device) if want_u8 else None
        f32 = torch.empty((B, Cn, H // 4, W // 4), dtype=torch.float32, device=hr.device) if want_f32 else None
        rc = lib.fen_lr_from_hr_u8(hr.data_ptr(), u8.data_ptr() if want_u8 else None,
                                   f32.data_ptr() if want_f32 else None, B, H, W, Cn,
                                   torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "fen_lr_from_hr_u8")
    return u8, f32


def lr_from_hr_float(hr: torch.Tensor, want_f32: bool = True, want_u8: bool = False, bgr: bool = False
                     ) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """The trainer's / the scripts' float LR generator on the GPU:
    F.interpolate(hr, scale_factor=0.25, mode='bicubic', align_corners=False) (trainer.py:416-421) and,
    with want_u8, generate_lr of scripts/test_model.py:139-156 (np.clip(lr * 255, 0, 255).astype(uint8), HWC,
    BGR when bgr).  hr: float32 CUDA tensor [B,C,H,W] in [0,1].  Returns (lr_f32 [B,C,H/4,W/4], lr_u8 [B,H/4,W/4,C])."""
    if hr.dtype != torch.float32 or hr.dim() != 4:
        raise TypeError("hr must be a float32 tensor [B,C,H,W]")
    if not hr.is_cuda:
        raise RuntimeError("lr_from_hr_float needs a CUDA tensor: there is no CPU fallback")
    if not (want_u8 or want_f32):
        raise ValueError("nothing requested")
    hr = hr.contiguous()
    B, Cn, H, W = hr.shape
    if H % 4 or W % 4:
        raise ValueError("H and W must be multiples of 4")
    lib = _lib.load()
    with torch.cuda.device(hr.device):
        f32 = torch.empty((B, Cn, H // 4, W // 4), dtype=torch.float32, device=hr.device) if want_f32 else None
        u8 = torch.empty((B, H // 4, W // 4, Cn), dtype=torch.uint8, device=hr.device) if want_u8 else None
        rc = lib.fen_lr_from_hr_f32(hr.data_ptr(), f32.data_ptr() if want_f32 else None,
                                    u8.data_ptr() if want_u8 else None, B, Cn, H, W, int(bgr),
                                    torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "fen_lr_from_hr_f32")
    return f32, u8


def sr_to_uint8(sr: torch.Tensor, bgr: bool = False) -> torch.Tensor:
    """to_numpy of the evaluation scripts (scripts/test_model.py:176-190) on the GPU: float32 [B,C,H,W] ->
    uint8 [B,H,W,C], np.clip(x * 255, 0, 255).astype(uint8) (truncation), BGR channel order when bgr."""
    if sr.dtype != torch.float32 or sr.dim() != 4:
        raise TypeError("sr must be a float32 tensor [B,C,H,W]")
    if not sr.is_cuda:
        raise RuntimeError("sr_to_uint8 needs a CUDA tensor: there is no CPU fallback")
    sr = sr.contiguous()
    B, Cn, H, W = sr.shape
    lib = _lib.load()
    with torch.cuda.device(sr.device):
        out = torch.empty((B, H, W, Cn), dtype=torch.uint8, device=sr.device)
        rc = lib.fen_sr_to_u8(sr.data_ptr(), out.data_ptr(), B, Cn, H, W, int(bgr), torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "fen_sr_to_u8")
    return out


def create_lr_image(hr_image: np.ndarray, lr_size: int = 64, method: str = "bicubic") -> np.ndarray:
    """Same signature as the reference's create_lr_image (prepare_data.py:23-59): HWC uint8 numpy in,
    HWC uint8 numpy out, computed by the CUDA kernel.  Only the path's method ('bicubic') at the exact
    /4 ratio is implemented; other methods are outside the hot path and raise."""
    if method != "bicubic":
        if method in ("bilinear", "realistic"):
            raise NotImplementedError(f"method '{method}' is outside the B200 hot path")
        raise ValueError(f"Unknown degradation method: {method}")
    if hr_image.dtype != np.uint8 or hr_image.ndim != 3:
        raise TypeError("hr_image must be uint8 (H, W, C)")
    H, W, _ = hr_image.shape
    if H != 4 * lr_size or W != 4 * lr_size:
        raise NotImplementedError("only the exact 4:1 ratio of the reference pipeline is implemented")
    hr = torch.from_numpy(np.ascontiguousarray(hr_image)).cuda().unsqueeze(0)
    u8, _ = lr_from_hr(hr, want_u8=True, want_f32=False)
    return u8[0].cpu().numpy()


def to_tensor(image, normalize: bool = True) -> torch.Tensor:
    """transforms.py:260-279: HWC uint8 -> CHW, float / 255.0 when normalize."""
    t = torch.as_tensor(image)
    t = t.permute(2, 0, 1).contiguous()
    return t.float() / 255.0 if normalize else t
