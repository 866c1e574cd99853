"""Autograd node over the fused Dice / BCE / MSE reductions of ``csrc/segloss.cu`` (the Dice-family losses beyond DiceLoss)."""
import ctypes

import torch

from .._lib import BF16, F32, call

_REDUCE = {"sum": 0, "mean": 1, "max": 2, "min": 3, None: 4}


def _vp(t, offset_elems=0):
    return ctypes.c_void_p(t.data_ptr() + offset_elems * t.element_size())


def _dt(t):
    return BF16 if t.dtype == torch.bfloat16 else F32


def _stream(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


class SegLossFn(torch.autograd.Function):
    """loss(pred (N, C, *spatial), target (N, C, *spatial)); ``chan``: C x (w_dice, w_bce, w_mse, use_mask) host list;
    ``mask_channel``: None (no mask) or the index of the TARGET channel that masks the channels with use_mask = 1."""

    @staticmethod
    def forward(ctx, pred, target, chan, logits, mask_channel, channelwise, eps, reduce):
        if pred.device.type != "cuda":
            raise RuntimeError("b200em loss: tensors must live on a CUDA device (no CPU fallback on this path)")
        p = pred.detach()
        if p.dtype not in (torch.float32, torch.bfloat16):
            p = p.float()
        p = p.contiguous()
        N, C = p.shape[0], p.shape[1]
        S = p[0, 0].numel()
        t = target.detach()
        t = (t if t.dtype == torch.float32 else t.float()).contiguous()
        dev = p.device
        chan_t = torch.tensor(chan, dtype=torch.float32, device=dev).reshape(C, 4)
        sums = torch.zeros((C, 5), dtype=torch.float32, device=dev)
        coef = torch.empty((C, 4), dtype=torch.float32, device=dev)
        per_channel = bool(channelwise and reduce is None)
        loss = torch.empty((C if per_channel else 1,), dtype=torch.float32, device=dev)
        nstride = C * S
        mask_ptr = _vp(t, mask_channel * S) if mask_channel is not None else None
        with torch.cuda.device(dev):
            st = _stream(p)
            call("b200em_segloss_sums", _vp(p), _dt(p), _vp(t), mask_ptr, nstride, nstride, 0, _vp(chan_t), int(logits), N, C, S,
                 _vp(sums), st)
            call("b200em_segloss_finalize", _vp(sums), _vp(chan_t), C, float(eps), int(bool(channelwise)), _REDUCE[reduce],
                 float(N * S), _vp(loss), _vp(coef), st)
        ctx.saved = (p, t, chan_t, coef, mask_channel, int(logits), (N, C, S), pred.dtype, pred.shape, per_channel)
        return loss if per_channel else loss.reshape(())

    @staticmethod
    def backward(ctx, gout):
        p, t, chan_t, coef, mask_channel, logits, (N, C, S), in_dtype, in_shape, per_channel = ctx.saved
        g = torch.empty(p.shape, dtype=p.dtype, device=p.device)
        go = gout.detach().float().contiguous().reshape(-1)
        nstride = C * S
        mask_ptr = _vp(t, mask_channel * S) if mask_channel is not None else None
        with torch.cuda.device(p.device):
            call("b200em_segloss_bwd", _vp(p), _dt(p), _vp(t), mask_ptr, nstride, nstride, 0, _vp(chan_t), _vp(coef), _vp(go),
                 int(per_channel), logits, _vp(g), _dt(g), N, C, S, _stream(p))
        g = g.reshape(in_shape)
        if g.dtype != in_dtype:
            g = g.to(in_dtype)
        return g, None, None, None, None, None, None, None
