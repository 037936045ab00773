"""ctypes binding of the C ABI declared in include/fen_b200.h.  Loading fails loudly when the CUDA
library has not been built: there is no CPU or PyTorch fallback for this path."""
from __future__ import annotations

import ctypes as C
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# FEN_B200_LIB: developer override (e.g. a -DFEN_BODY_DEBUG build next to the product library)
LIB_PATH = os.environ.get("FEN_B200_LIB") or os.path.join(PKG_DIR, "libfen_b200.so")

FEN_OK, FEN_EINVAL, FEN_ENODEV, FEN_ENOMEM, FEN_ECUDA = 0, -1, -2, -3, -4


class FenConfig(C.Structure):
    _fields_ = [
        ("num_channels", C.c_int32), ("num_groups", C.c_int32), ("blocks_per_group", C.c_int32),
        ("reduction_ratio", C.c_int32), ("scale_factor", C.c_int32), ("res_scale", C.c_float),
    ]


# name -> (restype, argtypes); must list every symbol include/fen_b200.h declares
SIGNATURES = {
    "fen_abi_version": (C.c_int, []),
    "fen_last_error": (C.c_char_p, []),
    "fen_param_count": (C.c_int64, [C.POINTER(FenConfig)]),
    "fen_packed_bytes": (C.c_int64, [C.POINTER(FenConfig)]),
    "fen_pack_weights": (C.c_int, [C.POINTER(FenConfig), C.c_void_p, C.c_void_p, C.c_void_p]),
    "fen_forward_workspace_bytes": (C.c_int64, [C.POINTER(FenConfig), C.c_int, C.c_int, C.c_int]),
    "fen_forward": (C.c_int, [C.POINTER(FenConfig), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                              C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "fen_forward_u8": (C.c_int, [C.POINTER(FenConfig), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                 C.c_int, C.c_void_p, C.c_int64, C.c_void_p]),
    "fen_forward_tap": (C.c_int64, [C.POINTER(FenConfig), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.POINTER(C.c_void_p)]),
    "fen_lr_from_hr_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_void_p]),
    "fen_lr_from_hr_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p]),
    "fen_sr_to_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "fen_train_workspace_bytes": (C.c_int64, [C.c_int64]),
    "fen_l1_loss": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                              C.c_void_p]),
    "fen_psnr": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "fen_grad_norm": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "fen_clip_adamw_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                      C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int,
                                      C.c_void_p]),
    "fen_ssim_workspace_bytes": (C.c_int64, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "fen_ssim": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int,
                           C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "fen_packed_bwd_bytes": (C.c_int64, [C.POINTER(FenConfig)]),
    "fen_pack_weights_bwd": (C.c_int, [C.POINTER(FenConfig), C.c_void_p, C.c_void_p, C.c_void_p]),
    "fen_step_workspace_bytes": (C.c_int64, [C.POINTER(FenConfig), C.c_int, C.c_int, C.c_int]),
    "fen_forward_train": (C.c_int, [C.POINTER(FenConfig), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                    C.c_int, C.c_void_p, C.c_int64, C.c_void_p]),
    "fen_backward": (C.c_int, [C.POINTER(FenConfig), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                               C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]),
    "fen_backward_num_stages": (C.c_int, [C.POINTER(FenConfig)]),
    "fen_backward_stage_range": (C.c_int, [C.POINTER(FenConfig), C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "fen_backward_stages": (C.c_int, [C.POINTER(FenConfig), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p]),
    "fen_conv3x3_c64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "fen_pack_conv3x3": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "fen_last_launch_count": (C.c_int, []),
    "fen_profile_body": (C.c_int, [C.c_int]),
    "fen_last_body_ms": (C.c_float, []),
}

_lib = None


def load() -> C.CDLL:
    """Load libfen_b200.so and bind every entry point; raises RuntimeError if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). This package has no CPU / PyTorch fallback for the FaceEnhanceNet path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    if lib.fen_abi_version() != 1:
        raise RuntimeError("libfen_b200.so ABI version mismatch; rebuild it")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    """Raise the reference-style RuntimeError for a failed call."""
    if rc < 0:
        msg = load().fen_last_error().decode("utf-8", "replace")
        if rc == FEN_EINVAL:
            raise ValueError(f"{what}: {msg}")
        raise RuntimeError(f"{what} failed ({rc}): {msg}")
