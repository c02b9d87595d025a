"""synt_isic_b200 -- B200-native (sm_100a) implementation of the SYNT_ISIC hot path:
the DDPM reverse-diffusion loop (UNet2D forward + DDPMScheduler.step per timestep) and the
ResNet18 logit evaluations issued by Time-SHAP / patch-SHAP / CSI.

Python here is host glue that mirrors the reference's object protocol; every arithmetic
operation of the path runs in ``libsynt_isic_b200.so`` (hand-written CUDA, C ABI in
include/synt_isic.h).  There is no CPU or PyTorch fallback.
"""
from ._lib import LIB_PATH, lib  # noqa: F401
from .scheduler import DDPMScheduler  # noqa: F401
from .unet import SUPPORTED_CONFIG, UNet2DModel  # noqa: F401
from .classifier import CLASS_NAMES, MelanomaClassifierAdaptive  # noqa: F401

__all__ = ["UNet2DModel", "DDPMScheduler", "MelanomaClassifierAdaptive", "CLASS_NAMES", "SUPPORTED_CONFIG", "lib",
           "LIB_PATH"]
