"""Import shim: the product package lives in ``face-super-resolution_b200/`` (a directory name Python
cannot import directly); this package re-roots its module search path there, so
``import fsr_b200`` / ``from fsr_b200.model import FaceEnhanceNet`` load those files."""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "face-super-resolution_b200")
__path__.insert(0, _PKG_DIR)

from .model import (  # noqa: E402,F401
    FaceEnhanceNet,
    FaceEnhanceNetConfig,
    FaceEnhanceNetLite,
    create_face_enhance_net,
)
from .blocks import (  # noqa: E402,F401
    ChannelAttention,
    RCAB,
    ResidualGroup,
    PixelShuffleUpsample,
    UpsampleModule,
)
from .training import ClipAdamW, SSIMLoss, Stage1Step, allreduce_mean_, l1_loss, psnr, ssim  # noqa: E402,F401
from .data import create_lr_image, lr_from_hr, lr_from_hr_float, sr_to_uint8, to_tensor  # noqa: E402,F401

__all__ = [
    "FaceEnhanceNet", "FaceEnhanceNetConfig", "FaceEnhanceNetLite", "create_face_enhance_net",
    "ChannelAttention", "RCAB", "ResidualGroup", "PixelShuffleUpsample", "UpsampleModule",
    "ClipAdamW", "SSIMLoss", "ssim", "Stage1Step", "allreduce_mean_", "l1_loss", "psnr", "create_lr_image", "lr_from_hr", "lr_from_hr_float", "sr_to_uint8", "to_tensor",
]
