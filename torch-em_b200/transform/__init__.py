from .label import (AffinityTransform, BoundaryTransform, BoundaryTransformWithIgnoreLabel, NoToBackgroundBoundaryTransform,
                    OneHotTransform, labels_to_binary, segmentation_to_affinities)

__all__ = ["AffinityTransform", "BoundaryTransform", "NoToBackgroundBoundaryTransform", "BoundaryTransformWithIgnoreLabel",
           "OneHotTransform", "segmentation_to_affinities", "labels_to_binary"]
