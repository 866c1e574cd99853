"""``UNet3d`` / ``AnisotropicUNet`` with torch-em's constructor, ``forward`` signature, ``init_kwargs`` and
``state_dict`` key names (torch_em/model/unet.py:584-728), computed by the sm_100a kernels of ``libb200em``.

Drop-in contract: pass an instance to ``torch_em.default_segmentation_trainer`` unchanged; reference
checkpoints load through ``load_state_dict`` (same keys and shapes: ``encoder.blocks.{l}.block.{1,4}.*``,
``base.block.{1,4}.*``, ``decoder.blocks.{l}.block.{1,4}.*``, ``decoder.samplers.{l}.conv.*``, ``out_conv.*``, plus
``block.{0,3}.*`` for GroupNorm; ``block.{0,2}`` when ``norm=None``).

The sub-modules below hold parameters only -- they are never called.  ``forward`` runs the whole network as ONE
``torch.autograd.Function`` whose forward / backward are explicit kernel schedules (``engine.py``); parameters stay
ordinary fp32 ``nn.Parameter`` objects, so AdamW, DDP, ``state_dict`` and checkpoints are untouched.
"""
import math
from typing import List, Optional, Union

import torch
import torch.nn as nn

from .. import engine
from ..backend import default_backend

__all__ = ["UNet3d", "AnisotropicUNet", "ConvBlock3d", "Upsampler3d"]


class _ConvParams(nn.Module):
    """Weight and bias of one nn.Conv3d (same shapes, same default init as torch: kaiming_uniform(a=sqrt(5)))."""

    def __init__(self, in_channels, out_channels, kernel_size):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = tuple(kernel_size)
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels, *self.kernel_size))
        self.bias = nn.Parameter(torch.empty(out_channels))
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        fan_in = self.in_channels * self.kernel_size[0] * self.kernel_size[1] * self.kernel_size[2]
        bound = 1 / math.sqrt(fan_in) if fan_in > 0 else 0
        nn.init.uniform_(self.bias, -bound, bound)

    def extra_repr(self):
        return f"{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}"


class _AffineParams(nn.Module):
    """gamma / beta of one nn.GroupNorm."""

    def __init__(self, channels):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(channels))
        self.bias = nn.Parameter(torch.zeros(channels))


class _Slot(nn.Module):
    """Parameter-free position in a block (ReLU / InstanceNorm): keeps the reference's Sequential indices."""


_SUPPORTED_NORMS = ("InstanceNorm", "GroupNorm", None)


class ConvBlock3d(nn.Module):
    """Parameter holder for Norm -> Conv3d -> ReLU -> Norm -> Conv3d -> ReLU (unet.py:409-441, 570-574)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, padding=1, norm="InstanceNorm"):
        super().__init__()
        if norm not in _SUPPORTED_NORMS:
            raise NotImplementedError(
                f"norm={norm!r}: the B200 path implements 'InstanceNorm' (reference default), 'GroupNorm' and None; "
                "'BatchNorm' / 'InstanceNormTrackStats' couple samples through running statistics and are not built yet")
        k = (kernel_size,) * 3 if isinstance(kernel_size, int) else tuple(kernel_size)
        p = (padding,) * 3 if isinstance(padding, int) else tuple(padding)
        if any(kk not in (1, 3) for kk in k) or any(pp != kk // 2 for kk, pp in zip(k, p)):
            raise NotImplementedError(f"kernel_size={kernel_size}, padding={padding}: only 'same' 3x3x3 / 1x3x3 kernels")
        self.in_channels, self.out_channels, self.kernel_size, self.norm = in_channels, out_channels, k, norm
        if norm is None:
            mods = [_ConvParams(in_channels, out_channels, k), _Slot(), _ConvParams(out_channels, out_channels, k), _Slot()]
        else:
            n1 = _AffineParams(in_channels) if norm == "GroupNorm" else _Slot()
            n2 = _AffineParams(out_channels) if norm == "GroupNorm" else _Slot()
            mods = [n1, _ConvParams(in_channels, out_channels, k), _Slot(), n2,
                    _ConvParams(out_channels, out_channels, k), _Slot()]
        self.block = nn.Sequential(*mods)

    def forward(self, x):
        raise RuntimeError("ConvBlock3d holds parameters only; the network runs through UNet3d.forward")


class Upsampler3d(nn.Module):
    """Parameter holder for trilinear interpolate -> Conv3d(1x1x1) (unet.py:444-458, 577-581)."""

    def __init__(self, scale_factor, in_channels, out_channels, mode="trilinear"):
        super().__init__()
        if mode != "trilinear":
            raise NotImplementedError("only mode='trilinear' (the reference's Upsampler3d default)")
        self.scale_factor = scale_factor
        self.conv = _ConvParams(in_channels, out_channels, (1, 1, 1))


class _Encoder(nn.Module):
    def __init__(self, features, scale_factors, kernels, **kw):
        super().__init__()
        self.blocks = nn.ModuleList([ConvBlock3d(i, o, kernel_size=k, padding=tuple(kk // 2 for kk in k), **kw)
                                     for i, o, k in zip(features[:-1], features[1:], kernels)])
        self.in_channels, self.out_channels = features[0], features[-1]

    def __len__(self):
        return len(self.blocks)


class _Decoder(nn.Module):
    def __init__(self, features, scale_factors, kernels, **kw):
        super().__init__()
        self.blocks = nn.ModuleList([ConvBlock3d(i, o, kernel_size=k, padding=tuple(kk // 2 for kk in k), **kw)
                                     for i, o, k in zip(features[:-1], features[1:], kernels)])
        self.samplers = nn.ModuleList([Upsampler3d(f, i, o) for f, i, o in zip(scale_factors, features[:-1], features[1:])])
        self.in_channels, self.out_channels = features[0], features[-1]

    def __len__(self):
        return len(self.blocks)


def _device_ctx(device):
    import contextlib
    return torch.cuda.device(device) if device.type == "cuda" else contextlib.nullcontext()


class _LazyPacks:
    """Mapping key -> pack for backends without a batched pack set (the CPU emulation used by the host-logic tests)."""

    def __init__(self, B, P):
        self._B, self._P = B, P

    def __getitem__(self, key):
        return self._B.pack(key, self._P[key + ".weight"])

    def refresh_fwd(self, weights=None, bf16=True):
        pass

    def refresh_dgrad(self, bf16=True):
        pass


class _UNetFunction(torch.autograd.Function):
    """The whole network as one autograd node: forward_pass / backward_pass are explicit kernel schedules."""

    @staticmethod
    def forward(ctx, model, act_dtype, x, *params):
        if ctx.needs_input_grad[2]:
            raise NotImplementedError(
                "the fused U-Net node does not compute the gradient w.r.t. its input (the reference's trainer never asks for "
                "it, default_trainer.py:805-831); detach the input or set requires_grad=False")
        names = model._param_names
        P = dict(zip(names, params))
        B = model._backend()
        bf16 = act_dtype == torch.bfloat16
        with _device_ctx(x.device):
            packs = model._packs(B, P)
            packs.refresh_fwd({k[:-len(".weight")]: v for k, v in P.items() if v.dim() == 5}, bf16=bf16)
            pred, fctx = engine.forward_pass(B, model._plan, P, x.detach(), act_dtype, packs)
        ctx.model, ctx.P, ctx.packs, ctx.bf16 = model, P, packs, bf16
        ctx.fctx = fctx if any(ctx.needs_input_grad) else None   # no_grad / eval inference keeps nothing
        ctx.ran_backward = False
        if ctx.fctx is not None:
            # the head's backward reads the prediction (sigmoid / tanh derivative): saving it through autograd makes an
            # in-place edit of the returned tensor (pred.clamp_()) raise instead of silently corrupting the gradient
            ctx.save_for_backward(pred)
            fctx.misc.pop("pred", None)          # (a direct reference would close a cycle pred -> grad_fn -> ctx -> pred)
        return pred

    @staticmethod
    def backward(ctx, grad_pred):
        model = ctx.model
        if ctx.fctx is None:
            if ctx.ran_backward:
                raise RuntimeError(
                    "the fused U-Net node releases its activations after the first backward pass: a second backward through "
                    "the same forward (retain_graph=True) is not supported -- run the forward again")
            raise RuntimeError("backward through a forward that ran without grad")
        (pred,) = ctx.saved_tensors
        ctx.fctx.misc["pred"] = pred
        B = model._backend()
        with _device_ctx(grad_pred.device):
            ctx.packs.refresh_dgrad(bf16=ctx.bf16)
            grads = engine.backward_pass(B, model._plan, ctx.P, ctx.fctx, grad_pred, ctx.packs)
            if model.grad_sync is not None:
                model.grad_sync(grads.flat)      # ONE collective over the flat gradient buffer (distributed.py)
        ctx.fctx = None
        ctx.ran_backward = True
        out = [None, None, None]
        for name in model._param_names:
            g = grads.get(name)
            if g is not None:
                g = g.reshape(ctx.P[name].shape)
            out.append(g)
        return tuple(out)


class AnisotropicUNet(nn.Module):
    """3D U-Net with per-level (possibly anisotropic) pooling factors; same arguments as
    ``torch_em.model.AnisotropicUNet`` (unet.py:610-624)."""

    def __init__(
        self,
        in_channels: int,
        out_channels: int,
        scale_factors: List[List[int]],
        initial_features: int = 32,
        gain: int = 2,
        final_activation: Optional[Union[str, nn.Module]] = None,
        return_side_outputs: bool = False,
        conv_block_impl: nn.Module = ConvBlock3d,
        anisotropic_kernel: bool = False,
        postprocessing: Optional[Union[str, nn.Module]] = None,
        check_shape: bool = True,
        **conv_block_kwargs,
    ):
        super().__init__()
        if getattr(conv_block_impl, "__name__", None) != "ConvBlock3d":
            raise NotImplementedError(
                f"conv_block_impl={conv_block_impl!r}: only the default ConvBlock3d is fused; a user-defined block cannot "
                "be accelerated and there is deliberately no silent fallback")
        if return_side_outputs:
            raise NotImplementedError("return_side_outputs=True is not built yet on the B200 path")
        if postprocessing is not None:
            raise NotImplementedError("postprocessing (bioimage.io channel accumulators) is not built yet on the B200 path")
        if out_channels is None:
            raise NotImplementedError("out_channels=None (return decoder features) is not built yet on the B200 path")
        unknown = set(conv_block_kwargs) - {"norm", "kernel_size", "padding"}
        if unknown:
            raise TypeError(f"unexpected conv block arguments: {sorted(unknown)}")
        if conv_block_kwargs.get("kernel_size", 3) != 3 or conv_block_kwargs.get("padding", 1) != 1:
            raise NotImplementedError("only kernel_size=3, padding=1 conv blocks")
        norm = conv_block_kwargs.get("norm", "InstanceNorm")
        act_name = self._activation_name(final_activation)

        depth = len(scale_factors)
        sfs = [engine._as_factor(sf) for sf in scale_factors]
        features_encoder = [in_channels] + [initial_features * gain ** i for i in range(depth)]
        features_decoder = [initial_features * gain ** i for i in range(depth + 1)][::-1]
        ek = engine.level_kernels(sfs, anisotropic_kernel)
        dk = engine.level_kernels(sfs[::-1], anisotropic_kernel)
        self.encoder = _Encoder(features_encoder, sfs, ek, norm=norm)
        self.base = ConvBlock3d(features_encoder[-1], features_encoder[-1] * gain, norm=norm)
        self.decoder = _Decoder(features_decoder, sfs[::-1], dk, norm=norm)
        self.out_conv = _ConvParams(features_decoder[-1], out_channels, (1, 1, 1))
        self._out_channels = out_channels
        self.return_decoder_outputs = False
        self.check_shape = check_shape
        self.final_activation = final_activation if isinstance(final_activation, nn.Module) else (
            None if final_activation is None else getattr(nn, final_activation)())
        self.postprocessing = None
        self.init_kwargs = {"in_channels": in_channels, "out_channels": out_channels, "scale_factors": scale_factors,
                            "initial_features": initial_features, "gain": gain,
                            "final_activation": final_activation, "return_side_outputs": return_side_outputs,
                            "conv_block_impl": conv_block_impl, "anisotropic_kernel": anisotropic_kernel,
                            "postprocessing": postprocessing, **conv_block_kwargs}
        self._plan = engine.make_plan(in_channels, out_channels, sfs, initial_features, gain, norm, act_name,
                                      anisotropic_kernel)
        self._param_names = [n for n, _ in self.named_parameters()]
        self._backend_override = None
        self.grad_sync = None       # callable(flat fp32 gradient buffer); set by torch_em_b200.distributed.sync_gradients
        self.compute_dtype = None   # None: follow torch.autocast (bf16) / fp32 otherwise; or torch.bfloat16 / torch.float32

    # ---- reference surface ----------------------------------------------------------------------------------
    @staticmethod
    def _activation_name(activation):
        if activation is None:
            return None
        name = activation if isinstance(activation, str) else type(activation).__name__
        if isinstance(activation, str) and getattr(nn, activation, None) is None:
            raise ValueError(f"Invalid activation: {activation}")
        if name not in ("Sigmoid", "ReLU", "Tanh"):
            raise NotImplementedError(f"final_activation={name!r}: the fused head implements Sigmoid, ReLU, Tanh or None")
        return name

    @property
    def in_channels(self):
        return self.encoder.in_channels

    @property
    def out_channels(self):
        return self._out_channels

    @property
    def depth(self):
        return len(self.encoder)

    def load_encoder_state(self, state):
        self.encoder.load_state_dict(state)

    def load_decoder_state(self, state):
        self.decoder.load_state_dict(state)

    def load_base_state(self, state):
        self.base.load_state_dict(state)

    def _check_shape(self, x):
        spatial_shape = tuple(x.shape)[2:]
        engine.check_shape(spatial_shape, self._plan.scale_factors)

    # ---- execution ------------------------------------------------------------------------------------------
    def _backend(self):
        return self._backend_override if self._backend_override is not None else default_backend()

    def _packs(self, B, P):
        """The model's operand images (backend.PackSet: one pack launch per direction and pass); owned by the model so that
        they are freed with it.  One set per backend object (tests swap backends on a live model)."""
        if not hasattr(B, "pack_set"):
            return _LazyPacks(B, P)
        store = self.__dict__.get("_pack_store")
        if store is None or store[0] is not B:
            weights = {k[:-len(".weight")]: v for k, v in P.items() if v.dim() == 5}
            store = (B, B.pack_set(weights))
            self.__dict__["_pack_store"] = store
        return store[1]

    def _activation_dtype(self, x):
        if self.compute_dtype is not None:
            return self.compute_dtype
        if torch.is_autocast_enabled(x.device.type):
            dt = torch.get_autocast_dtype(x.device.type)
            if dt == torch.float16:
                raise NotImplementedError(
                    "fp16 autocast (+GradScaler) is not supported by the B200 path; pass "
                    "mixed_precision_dtype='bfloat16' (or mixed_precision=False) to default_segmentation_trainer")
            return dt
        return torch.float32

    @torch.compiler.disable
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """(N, in_channels, D, H, W) -> (N, out_channels, D, H, W) fp32 (unet.py:237-253).

        Opted out of Dynamo tracing: ``DefaultTrainer`` wraps the model in ``torch.compile`` by default
        (default_trainer.py:541, util/util.py:38-74); the network already is one hand-scheduled autograd node, so the
        compiled wrapper simply calls this method eagerly."""
        if getattr(self, "check_shape", True):
            self._check_shape(x)
        params = [p for _, p in self.named_parameters()]
        return _UNetFunction.apply(self, self._activation_dtype(x), x, *params)

    def __deepcopy__(self, memo):
        # predict_with_halo deep-copies the model once per device (prediction.py:188-192)
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k == "_pack_store":
                continue                          # derived data: the copy packs its own operands on its own device
            new.__dict__[k] = v if k == "_backend_override" else copy.deepcopy(v, memo)
        return new


class UNet3d(AnisotropicUNet):
    """3D U-Net with isotropic 2x pooling per level; same arguments as ``torch_em.model.UNet3d`` (unet.py:701-714)."""

    def __init__(
        self,
        in_channels: int,
        out_channels: int,
        depth: int = 4,
        initial_features: int = 32,
        gain: int = 2,
        final_activation: Optional[Union[str, nn.Module]] = None,
        return_side_outputs: bool = False,
        conv_block_impl: nn.Module = ConvBlock3d,
        postprocessing: Optional[Union[str, nn.Module]] = None,
        check_shape: bool = True,
        **conv_block_kwargs,
    ):
        scale_factors = depth * [2]
        super().__init__(in_channels, out_channels, scale_factors, initial_features=initial_features, gain=gain,
                         final_activation=final_activation, return_side_outputs=return_side_outputs,
                         anisotropic_kernel=False, postprocessing=postprocessing, conv_block_impl=conv_block_impl,
                         check_shape=check_shape, **conv_block_kwargs)
        self.init_kwargs = {"in_channels": in_channels, "out_channels": out_channels, "depth": depth,
                            "initial_features": initial_features, "gain": gain,
                            "final_activation": final_activation, "return_side_outputs": return_side_outputs,
                            "conv_block_impl": conv_block_impl, "postprocessing": postprocessing, **conv_block_kwargs}
