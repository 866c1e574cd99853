from .label import AffinityTransform, BoundaryTransform, labels_to_binary

__all__ = ["AffinityTransform", "BoundaryTransform", "labels_to_binary"]
