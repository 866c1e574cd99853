"""Public surface of ``torch_em_b200``: the torch-em names for the accelerated path."""
from . import _lib, distributed, multi_gpu_training, util
from .multi_gpu_training import train_multi_gpu
from .loss import (AffinityLoss, ApplyAndRemoveMask, ApplyMask, BCEDiceLoss, BCEDiceLossWithLogits, CombinedLoss, DiceBasedDistanceLoss,
                   DiceLoss, DiceLossWithLogits, DistanceLoss, LossWrapper, MaskIgnoreLabel, dice_score)
from .model import AnisotropicUNet, UNet2d, UNet3d
from .transform import (AffinityTransform, BoundaryTransform, BoundaryTransformWithIgnoreLabel, NoToBackgroundBoundaryTransform,
                        OneHotTransform, segmentation_to_affinities)

__all__ = ["UNet2d", "UNet3d", "AnisotropicUNet", "DiceLoss", "DiceLossWithLogits", "BCEDiceLoss", "BCEDiceLossWithLogits", "CombinedLoss",
           "DistanceLoss", "DiceBasedDistanceLoss", "dice_score", "LossWrapper", "ApplyMask", "ApplyAndRemoveMask",
           "MaskIgnoreLabel", "AffinityLoss", "AffinityTransform", "BoundaryTransform", "NoToBackgroundBoundaryTransform",
           "BoundaryTransformWithIgnoreLabel", "OneHotTransform", "segmentation_to_affinities", "launch_count",
           "reset_launch_count", "distributed", "multi_gpu_training", "train_multi_gpu", "util"]

launch_count = _lib.launch_count
reset_launch_count = _lib.reset_launch_count
