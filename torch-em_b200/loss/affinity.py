"""``AffinityLoss``: masked Dice against affinity targets computed on the fly from instance labels.

The reference has no class of this name; its affinity-training idiom is
``LossWrapper(DiceLoss(), ApplyAndRemoveMask("multiply"))`` on a target built on the CPU by
``AffinityTransform(offsets, add_mask=True)`` (torch_em/cli.py:263-267, transform/label.py:248-327).  This module is
that idiom in two forms:

* ``AffinityLoss()`` called with a float target ``[C disaffinities | C masks]`` is exactly the reference idiom
  (fused masked Dice);
* ``AffinityLoss(offsets=...)`` called with integer instance labels computes target and mask inside the loss kernels
  (``csrc/labels.cu``: affinity_dice_sums / affinity_dice_bwd) -- the 2C-channel fp32 target never exists in HBM.
"""
import ctypes
from typing import List, Optional

import torch
import torch.nn as nn

from .._lib import BF16, F32, call
from ..transform.label import offsets_to_3d
from .dice import masked_dice

__all__ = ["AffinityLoss"]


def _vp(t):
    return ctypes.c_void_p(t.data_ptr())


def _dt(t):
    return BF16 if t.dtype == torch.bfloat16 else F32


def _stream(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


class _AffinityDiceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, labels, c_offsets, n_off, ignore_label, include_ignore_transitions, eps, reduce):
        if pred.device.type != "cuda":
            raise RuntimeError("b200em AffinityLoss: tensors must live on a CUDA device (no CPU fallback on this path)")
        p = pred.detach()
        if p.dtype not in (torch.float32, torch.bfloat16):
            p = p.float()
        p = p.contiguous()
        N, C, D, H, W = p.shape
        lab = labels.detach()
        if lab.dim() == 5:
            lab = lab[:, 0]
        lab = lab.to(torch.int64).contiguous()
        if tuple(lab.shape) != (N, D, H, W):
            raise ValueError(f"Expect labels of shape {(N, D, H, W)}, got: {tuple(lab.shape)}.")
        if C != n_off:
            raise ValueError(f"Expect one prediction channel per offset, got: {C} channels, {n_off} offsets.")
        dev = p.device
        sums = torch.zeros((C, 3), dtype=torch.float32, device=dev)
        coef = torch.empty((C, 2), dtype=torch.float32, device=dev)
        loss = torch.empty((1,), dtype=torch.float32, device=dev)
        has_ign, ign = int(ignore_label is not None), int(ignore_label or 0)
        red = {"sum": 0, "mean": 1, "max": 2, "min": 3}[reduce]
        with torch.cuda.device(dev):
            st = _stream(p)
            call("b200em_affinity_dice_sums", _vp(p), _dt(p), _vp(lab), N, D, H, W, c_offsets, n_off, has_ign, ign,
                 int(include_ignore_transitions), _vp(sums), st)
            call("b200em_dice_finalize", _vp(sums), C, float(eps), 1, red, _vp(loss), _vp(coef), st)
        ctx.saved = (p, lab, coef, c_offsets, n_off, has_ign, ign, int(include_ignore_transitions), pred.dtype)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, gout):
        p, lab, coef, c_offsets, n_off, has_ign, ign, iit, in_dtype = ctx.saved
        N, C, D, H, W = p.shape
        g = torch.empty_like(p)
        go = gout.detach().float().contiguous().reshape(-1)
        with torch.cuda.device(p.device):
            call("b200em_affinity_dice_bwd", _vp(p), _dt(p), _vp(lab), N, D, H, W, c_offsets, n_off, has_ign, ign, iit,
                 _vp(coef), _vp(go), _vp(g), _dt(g), _stream(p))
        if g.dtype != in_dtype:
            g = g.to(in_dtype)
        return g, None, None, None, None, None, None, None


class AffinityLoss(nn.Module):
    """Masked Dice on affinity targets; see the module docstring."""

    def __init__(self, offsets: Optional[List[List[int]]] = None, ignore_label: Optional[int] = None,
                 include_ignore_transitions: bool = False, eps: float = 1e-7, reduce_channel: str = "sum"):
        super().__init__()
        if reduce_channel not in ("sum", "mean", "max", "min"):
            raise ValueError(f"Unsupported channel reduction {reduce_channel}")
        self.offsets = offsets
        self.ignore_label = ignore_label
        self.include_ignore_transitions = include_ignore_transitions
        self.eps = eps
        self.reduce_channel = reduce_channel
        if offsets is not None:
            ndim, self._c_offsets = offsets_to_3d(offsets)
            if ndim != 3:
                raise NotImplementedError("on-the-fly affinity targets are implemented for 3-D offsets")
        self.init_kwargs = {"offsets": offsets, "ignore_label": ignore_label,
                            "include_ignore_transitions": include_ignore_transitions, "eps": eps,
                            "reduce_channel": reduce_channel}

    def forward(self, prediction: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        if target.is_floating_point():
            return masked_dice(prediction, target, True, self.eps, self.reduce_channel)
        if self.offsets is None:
            raise ValueError("AffinityLoss got integer labels but was built without offsets")
        return _AffinityDiceFn.apply(prediction, target, self._c_offsets, len(self.offsets), self.ignore_label,
                                     self.include_ignore_transitions, self.eps, self.reduce_channel)
