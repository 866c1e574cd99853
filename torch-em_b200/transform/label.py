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

__all__ = ["AffinityTransform", "BoundaryTransform", "labels_to_binary"]


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
