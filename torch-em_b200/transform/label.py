"""``AffinityTransform`` / ``BoundaryTransform`` with torch-em's arguments (torch_em/transform/label.py:100-129,
248-327), running as integer stencils on the GPU (``csrc/labels.cu``) over label volumes that are already on the
device -- instead of numpy/C++ in dataloader workers followed by a host->device copy of the float target.

Input: integer labels as a CUDA tensor ``(D, H, W)`` (one sample, like the reference's per-sample call; returns
``(channels, D, H, W)``) or batched ``(N, D, H, W)`` / ``(N, 1, D, H, W)`` (returns ``(N, channels, D, H, W)``);
2-D offsets work on ``(H, W)`` / ``(N, H, W)`` / ``(N, 1, H, W)``.  Output float32, channel order
``[fg?][C disaffinities][fg-mask?][C masks]`` (label.py:311-325) resp. ``[foreground?][boundary]`` (label.py:123-128).
"""
import ctypes
from typing import List, Optional

import torch

from .._lib import call

__all__ = ["AffinityTransform", "BoundaryTransform", "NoToBackgroundBoundaryTransform", "BoundaryTransformWithIgnoreLabel",
           "OneHotTransform", "segmentation_to_affinities", "labels_to_binary"]


def _stream(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _vp(t):
    return ctypes.c_void_p(t.data_ptr())


def _canon_labels(labels, ndim):
    """-> (int64 contiguous (N, D, H, W), batched?, had_channel_axis?, is2d)."""
    if not torch.is_tensor(labels) or labels.device.type != "cuda":
        raise RuntimeError("b200em label transforms take CUDA tensors (the CPU/numpy path is the reference's own)")
    if labels.is_floating_point():
        raise TypeError("labels must be an integer tensor")
    x = labels
    batched = True
    if x.dim() == ndim:
        x, batched = x[None], False
    elif x.dim() == ndim + 2:
        if x.shape[1] != 1:
            raise ValueError(f"expected a single label channel, got {x.shape[1]}")
        x = x[:, 0]
    elif x.dim() != ndim + 1:
        raise ValueError(f"labels of shape {tuple(labels.shape)} do not match ndim={ndim}")
    if ndim == 2:
        x = x[:, None]
    return x.to(torch.int64).contiguous(), batched


def offsets_to_3d(offsets):
    ndim = len(offsets[0])
    assert ndim in (2, 3)
    flat = []
    for off in offsets:
        assert len(off) == ndim
        flat.extend(([0] if ndim == 2 else []) + [int(o) for o in off])
    return ndim, (ctypes.c_int * len(flat))(*flat)


def labels_to_binary(labels: torch.Tensor, background_label: int = 0) -> torch.Tensor:
    """label.py:34-44."""
    return (labels != background_label).to(labels.dtype)


class AffinityTransform:
    """Instance labels -> (dis)affinity targets [+ validity masks] (label.py:248-327)."""

    def __init__(self, offsets: List[List[int]], ignore_label: Optional[int] = None, add_binary_target: bool = False,
                 add_mask: bool = False, include_ignore_transitions: bool = False):
        self.offsets = offsets
        self.ndim, self._c_offsets = offsets_to_3d(offsets)
        self.ignore_label = ignore_label
        self.add_binary_target = add_binary_target
        self.add_mask = add_mask
        self.include_ignore_transitions = include_ignore_transitions

    @property
    def n_channels(self):
        return (len(self.offsets) + int(self.add_binary_target)) * (2 if self.add_mask else 1)

    def __call__(self, labels: torch.Tensor) -> torch.Tensor:
        x, batched = _canon_labels(labels, self.ndim)
        N, D, H, W = x.shape
        out = torch.empty((N, self.n_channels, D, H, W), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            call("b200em_affinity_targets", _vp(x), _vp(out), N, D, H, W, self._c_offsets, len(self.offsets),
                 int(self.ignore_label is not None), int(self.ignore_label or 0), int(self.add_binary_target),
                 int(self.add_mask), int(self.include_ignore_transitions), _stream(x))
        if self.ndim == 2:
            out = out[:, :, 0]
        return out if batched else out[0]


class BoundaryTransform:
    """Instance labels -> boundary target, find_boundaries(mode="thick") semantics (label.py:100-129)."""

    def __init__(self, mode: str = "thick", add_binary_target: bool = False, ndim: Optional[int] = None):
        if mode != "thick":
            raise NotImplementedError("only mode='thick' (the reference default) is implemented on the GPU path")
        self.mode = mode
        self.add_binary_target = add_binary_target
        self.ndim = ndim

    def __call__(self, labels: torch.Tensor) -> torch.Tensor:
        ndim = self.ndim if self.ndim is not None else (labels.dim() if labels.dim() <= 3 else 3)
        x, batched = _canon_labels(labels, ndim)
        N, D, H, W = x.shape
        out = torch.empty((N, 2 if self.add_binary_target else 1, D, H, W), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            call("b200em_boundary_targets", _vp(x), _vp(out), N, D, H, W, int(self.add_binary_target), _stream(x))
        if ndim == 2:
            out = out[:, :, 0]
        return out if batched else out[0]


class _MaskedBoundaryTransform:
    _mode = 0

    def _run(self, labels, aux, bg):
        ndim = self.ndim if self.ndim is not None else (labels.dim() if labels.dim() <= 3 else 3)
        x, batched = _canon_labels(labels, ndim)
        N, D, H, W = x.shape
        out = torch.empty((N, 2 if self.add_binary_target else 1, D, H, W), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            call("b200em_boundary_targets_masked", _vp(x), _vp(out), N, D, H, W, int(self.add_binary_target), self._mode,
                 int(aux), int(bg), _stream(x))
        if ndim == 2:
            out = out[:, :, 0]
        return out if batched else out[0]


class NoToBackgroundBoundaryTransform(_MaskedBoundaryTransform):
    """Boundaries, with the boundaries TO the background set to ``mask_label`` (label.py:133-189).  float32 output
    (the reference returns int8 holding the same -1 / 0 / 1 values)."""
    _mode = 1

    def __init__(self, bg_label: int = 0, mask_label: int = -1, mode: str = "thick", add_binary_target: bool = False,
                 ndim: Optional[int] = None):
        if mode != "thick":
            raise NotImplementedError("only mode='thick' (the reference default) is implemented on the GPU path")
        self.bg_label = bg_label
        self.mask_label = mask_label
        self.mode = mode
        self.ndim = ndim
        self.add_binary_target = add_binary_target

    def __call__(self, labels: torch.Tensor) -> torch.Tensor:
        return self._run(labels, self.mask_label, self.bg_label)


class BoundaryTransformWithIgnoreLabel(_MaskedBoundaryTransform):
    """Boundaries, with the boundaries of the ignore region set to ``ignore_label`` (label.py:192-244).  float32 output."""
    _mode = 2

    def __init__(self, ignore_label: int = -1, mode: str = "thick", add_binary_target: bool = False, ndim: Optional[int] = None):
        if mode != "thick":
            raise NotImplementedError("only mode='thick' (the reference default) is implemented on the GPU path")
        self.ignore_label = ignore_label
        self.mode = mode
        self.ndim = ndim
        self.add_binary_target = add_binary_target

    def __call__(self, labels: torch.Tensor) -> torch.Tensor:
        return self._run(labels, self.ignore_label, 0)


class OneHotTransform:
    """Semantic labels -> one-hot float32 channels (label.py:330-353).  Like the reference, the channel axis is prepended to
    whatever shape the labels have: ``labels.shape -> (n_classes,) + labels.shape``; ``class_ids=None`` takes the sorted
    unique labels of the input (one device->host synchronisation, as np.unique is data dependent)."""

    def __init__(self, class_ids=None):
        self.class_ids = list(range(class_ids)) if isinstance(class_ids, int) else class_ids

    def __call__(self, labels: torch.Tensor) -> torch.Tensor:
        if not torch.is_tensor(labels) or labels.device.type != "cuda":
            raise RuntimeError("b200em label transforms take CUDA tensors (the CPU/numpy path is the reference's own)")
        x = labels.to(torch.int64).contiguous()
        if self.class_ids is None:
            ids = torch.unique(x)
        else:
            ids = torch.tensor([int(c) for c in self.class_ids], dtype=torch.int64, device=x.device)
        n_classes, S = int(ids.numel()), x.numel()
        out = torch.empty((n_classes,) + tuple(x.shape), dtype=torch.float32, device=x.device)
        if n_classes and S:
            with torch.cuda.device(x.device):
                call("b200em_one_hot", _vp(x), _vp(ids), n_classes, _vp(out), 1, S, _stream(x))
        return out


def segmentation_to_affinities(segmentation: torch.Tensor, offsets: List[List[int]]) -> torch.Tensor:
    """(N, 1, *spatial) segmentation -> (N, len(offsets), *spatial) float32 AFFINITIES (1 = same segment) with replication
    at the border (torch_em/loss/affinity_side_loss.py:70-89), as one integer stencil pass."""
    assert segmentation.shape[1] == 1, f"{segmentation.shape}"
    if segmentation.device.type != "cuda":
        raise RuntimeError("b200em label transforms take CUDA tensors (the CPU/numpy path is the reference's own)")
    ndim, c_offsets = offsets_to_3d(offsets)
    assert segmentation.dim() == ndim + 2
    seg = segmentation[:, 0]
    if seg.is_floating_point():
        seg = (seg.float() + 0.0).contiguous().view(torch.int32)     # equal floats <=> equal bit patterns (-0 folded into +0)
    x = seg.to(torch.int64)
    if ndim == 2:
        x = x[:, None]
    x = x.contiguous()
    N, D, H, W = x.shape
    out = torch.empty((N, len(offsets), D, H, W), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        call("b200em_segmentation_affinities", _vp(x), _vp(out), N, D, H, W, c_offsets, len(offsets), _stream(x))
    return out[:, :, 0] if ndim == 2 else out
