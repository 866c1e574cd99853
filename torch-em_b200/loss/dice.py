"""``DiceLoss`` / ``dice_score`` with torch-em's signature (torch_em/loss/dice.py:34-133), computed by the fused
sm_100a reductions of ``csrc/dice.cu``: one pass over prediction, target (and mask) forward, one pass backward.
"""
import ctypes
from typing import Optional

import torch
import torch.nn as nn

from .. import _lib
from .._lib import BF16, F32, call

__all__ = ["DiceLoss", "DiceLossWithLogits", "BCEDiceLoss", "BCEDiceLossWithLogits", "dice_score", "masked_dice"]

_REDUCE = {"sum": 0, "mean": 1, "max": 2, "min": 3, None: 4}


def _vp(t, offset_elems=0):
    return ctypes.c_void_p(t.data_ptr() + offset_elems * t.element_size())


def _dt(t):
    return BF16 if t.dtype == torch.bfloat16 else F32


def _stream(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _canon_pred(p):
    if p.dtype not in (torch.float32, torch.bfloat16):
        p = p.float()
    return p.contiguous()


class _DiceFn(torch.autograd.Function):
    """loss(pred (N,C,S), target/mask laid out with a common sample stride)."""

    @staticmethod
    def forward(ctx, pred, target, mask, mask_in_target, channelwise, eps, reduce):
        if pred.device.type != "cuda":
            raise RuntimeError("b200em DiceLoss: tensors must live on a CUDA device (no CPU fallback on this path)")
        p = _canon_pred(pred.detach())
        N, C = p.shape[0], p.shape[1]
        S = p[0, 0].numel()
        t = target.detach()
        if t.dtype != torch.float32:
            t = t.float()
        t = t.contiguous()
        nstride = t.shape[1] * S
        m_ptr = None
        m = None
        if mask_in_target:                       # ApplyAndRemoveMask: target = [C targets | C masks]
            m_ptr = _vp(t, C * S)
        elif mask is not None:
            m = mask.detach()
            m = (m if m.dtype == torch.float32 else m.float()).expand(t.shape).contiguous()
            m_ptr = _vp(m)
        sums = torch.zeros((C, 3), dtype=torch.float32, device=p.device)
        coef = torch.empty((C, 2), dtype=torch.float32, device=p.device)
        loss = torch.empty((C if (channelwise and reduce is None) else 1,), dtype=torch.float32, device=p.device)
        with torch.cuda.device(p.device):
            st = _stream(p)
            call("b200em_dice_sums", _vp(p), _dt(p), _vp(t), m_ptr, nstride, N, C, S, _vp(sums), st)
            call("b200em_dice_finalize", _vp(sums), C, float(eps), int(bool(channelwise)), _REDUCE[reduce], _vp(loss),
                 _vp(coef), st)
        ctx.saved = (p, t, m, mask_in_target, coef, nstride, (N, C, S), pred.dtype, pred.shape)
        ctx.per_channel = bool(channelwise and reduce is None)
        return loss if ctx.per_channel else loss.reshape(())

    @staticmethod
    def backward(ctx, gout):
        p, t, m, mask_in_target, coef, nstride, (N, C, S), in_dtype, in_shape = ctx.saved
        g = torch.empty(p.shape, dtype=p.dtype, device=p.device)
        go = gout.detach().float().contiguous().reshape(-1)
        m_ptr = _vp(t, C * S) if mask_in_target else (_vp(m) if m is not None else None)
        with torch.cuda.device(p.device):
            call("b200em_dice_bwd", _vp(p), _dt(p), _vp(t), m_ptr, nstride, _vp(coef), _vp(go), int(ctx.per_channel),
                 _vp(g), _dt(g), N, C, S, _stream(p))
        g = g.reshape(in_shape)
        if g.dtype != in_dtype:
            g = g.to(in_dtype)
        return g, None, None, None, None, None, None


def _flatten_to_ncs(x):
    """(N, C, *spatial) stays; a cropped (Nvalid, C) tensor becomes (1, C, Nvalid): Dice pools every axis but C."""
    if x.dim() == 2:
        return x.t().unsqueeze(0)
    return x


def dice_score(input_: torch.Tensor, target: torch.Tensor, invert: bool = False, channelwise: bool = True,
               reduce_channel: Optional[str] = "sum", eps: float = 1e-7) -> torch.Tensor:
    """Dice score between input and target (torch_em/loss/dice.py:34-93)."""
    if input_.shape != target.shape:
        raise ValueError(f"Expect input and target of same shape, got: {input_.shape}, {target.shape}.")
    if channelwise and reduce_channel not in _REDUCE:
        raise ValueError(f"Unsupported channel reduction {reduce_channel}")
    loss = _DiceFn.apply(_flatten_to_ncs(input_), _flatten_to_ncs(target), None, False, channelwise, eps,
                         reduce_channel if channelwise else "sum")
    if invert:
        return loss
    # score = 1 - loss per channel; for the reduced forms undo the inversion on the reduced value
    if not channelwise or reduce_channel is None or reduce_channel == "mean":
        return 1.0 - loss
    if reduce_channel == "sum":
        return input_.shape[1] - loss
    # max(1 - s) = 1 - min(s): the extremum flips, so evaluate the opposite reduction
    other = "min" if reduce_channel == "max" else "max"
    return 1.0 - _DiceFn.apply(_flatten_to_ncs(input_), _flatten_to_ncs(target), None, False, True, eps, other)


def masked_dice(prediction, target_with_mask, channelwise=True, eps=1e-7, reduce_channel="sum"):
    """LossWrapper(DiceLoss, ApplyAndRemoveMask('multiply')) in one kernel: target = [C targets | C masks]
    (loss/wrapper.py:84-87, 129-152)."""
    assert target_with_mask.dim() == prediction.dim(), f"{target_with_mask.dim()}, {prediction.dim()}"
    assert target_with_mask.size(1) == 2 * prediction.size(1), f"{target_with_mask.size(1)}, {prediction.size(1)}"
    assert target_with_mask.shape[2:] == prediction.shape[2:], f"{str(target_with_mask.shape)}, {str(prediction.shape)}"
    return _DiceFn.apply(prediction, target_with_mask, None, True, channelwise, eps, reduce_channel if channelwise else "sum")


class DiceLoss(nn.Module):
    """Dice error between a binary input and binary target (torch_em/loss/dice.py:96-133)."""

    def __init__(self, channelwise: bool = True, eps: float = 1e-7, reduce_channel: Optional[str] = "sum"):
        if reduce_channel not in ("sum", "mean", "max", "min", None):
            raise ValueError(f"Unsupported channel reduction {reduce_channel}")
        super().__init__()
        self.channelwise = channelwise
        self.eps = eps
        self.reduce_channel = reduce_channel
        self.init_kwargs = {"channelwise": channelwise, "eps": self.eps, "reduce_channel": self.reduce_channel}

    def forward(self, input_: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        return dice_score(input_=input_, target=target, invert=True, channelwise=self.channelwise, eps=self.eps,
                          reduce_channel=self.reduce_channel)

    def forward_masked(self, input_, target, mask):
        """Dice of (input*mask, target*mask) without materialising the products."""
        if input_.shape != target.shape:
            raise ValueError(f"Expect input and target of same shape, got: {input_.shape}, {target.shape}.")
        return _DiceFn.apply(_flatten_to_ncs(input_), _flatten_to_ncs(target), _flatten_to_ncs(mask), False,
                             self.channelwise, self.eps, self.reduce_channel if self.channelwise else "sum")


def _check_same_shape(input_, target):
    if input_.shape != target.shape:
        raise ValueError(f"Expect input and target of same shape, got: {input_.shape}, {target.shape}.")


class DiceLossWithLogits(nn.Module):
    """Dice error between logits and a binary target (torch_em/loss/dice.py:136-173): the sigmoid is applied inside the
    fused reduction and its derivative inside the fused backward (csrc/segloss.cu)."""

    def __init__(self, channelwise: bool = True, eps: float = 1e-7, reduce_channel: Optional[str] = "sum"):
        if reduce_channel not in ("sum", "mean", "max", "min", None):
            raise ValueError(f"Unsupported channel reduction {reduce_channel}")
        super().__init__()
        self.channelwise = channelwise
        self.eps = eps
        self.reduce_channel = reduce_channel
        self.init_kwargs = {"channelwise": channelwise, "eps": self.eps, "reduce_channel": self.reduce_channel}

    def forward(self, input_: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        from .segloss import SegLossFn
        _check_same_shape(input_, target)
        C = input_.shape[1]
        return SegLossFn.apply(input_, target, [[1.0, 0.0, 0.0, 0.0]] * C, True, None, self.channelwise, self.eps,
                               self.reduce_channel if self.channelwise else "sum")


class BCEDiceLoss(nn.Module):
    """alpha * Dice + beta * binary cross entropy between binary inputs and target (torch_em/loss/dice.py:176-214), one
    fused pass forward and one backward."""
    _logits = False

    def __init__(self, alpha: float = 1.0, beta: float = 1.0, channelwise: bool = True, eps: float = 1e-7):
        super().__init__()
        self.alpha = alpha
        self.beta = beta
        self.channelwise = channelwise
        self.eps = eps
        self.init_kwargs = {"alpha": alpha, "beta": beta, "channelwise": channelwise, "eps": self.eps}

    def forward(self, input_: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        from .segloss import SegLossFn
        _check_same_shape(input_, target)
        C = input_.shape[1]
        # the BCE is a mean over ALL elements (N*C*S): per-channel weight beta / C on a per-channel mean over N*S
        chan = [[float(self.alpha), float(self.beta) / C, 0.0, 0.0]] * C
        return SegLossFn.apply(input_, target, chan, self._logits, None, self.channelwise, self.eps, "sum")


class BCEDiceLossWithLogits(BCEDiceLoss):
    """alpha * Dice(sigmoid(x)) + beta * BCE-with-logits(x) (torch_em/loss/dice.py:217-256)."""
    _logits = True
