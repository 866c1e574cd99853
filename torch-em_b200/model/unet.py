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

__all__ = ["UNet2d", "UNet3d", "AnisotropicUNet", "ConvBlock2d", "ConvBlock3d", "Upsampler2d", "Upsampler3d"]


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
        fan_in = self.in_channels * math.prod(self.kernel_size)
        bound = 1 / math.sqrt(fan_in) if fan_in > 0 else 0
        nn.init.uniform_(self.bias, -bound, bound)

    def extra_repr(self):
        return f"{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}"


class _AffineParams(nn.Module):
    """gamma / beta of one nn.GroupNorm; with ``running=True`` also the running statistics of nn.BatchNorm3d /
    nn.InstanceNorm3d(track_running_stats=True) under the reference's buffer names (unet.py:391-406)."""

    def __init__(self, channels, running=False):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(channels))
        self.bias = nn.Parameter(torch.zeros(channels))
        if running:
            self.register_buffer("running_mean", torch.zeros(channels))
            self.register_buffer("running_var", torch.ones(channels))
            self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))


class _Slot(nn.Module):
    """Parameter-free position in a block (ReLU / InstanceNorm): keeps the reference's Sequential indices."""


_SUPPORTED_NORMS = ("InstanceNorm", "GroupNorm", "BatchNorm", "InstanceNormTrackStats", None)


class ConvBlock3d(nn.Module):
    """Parameter holder for Norm -> Conv3d -> ReLU -> Norm -> Conv3d -> ReLU (unet.py:409-441, 570-574)."""
    _dim = 3

    def __init__(self, in_channels, out_channels, kernel_size=3, padding=1, norm="InstanceNorm"):
        super().__init__()
        if norm not in _SUPPORTED_NORMS:
            raise ValueError(f"Invalid norm: expect one of 'InstanceNorm', 'BatchNorm' or 'GroupNorm', got {norm}")
        dim = self._dim
        k = (kernel_size,) * dim if isinstance(kernel_size, int) else tuple(kernel_size)
        p = (padding,) * dim if isinstance(padding, int) else tuple(padding)
        if len(k) != dim or any(kk not in (1, 3) for kk in k) or any(pp != kk // 2 for kk, pp in zip(k, p)):
            raise NotImplementedError(f"kernel_size={kernel_size}, padding={padding}: only 'same' kernels of extent 1 or 3")
        self.in_channels, self.out_channels, self.kernel_size, self.norm = in_channels, out_channels, k, norm
        if norm is None:
            mods = [_ConvParams(in_channels, out_channels, k), _Slot(), _ConvParams(out_channels, out_channels, k), _Slot()]
        else:
            affine, running = norm in engine.AFFINE_NORMS, norm in engine.RUNNING_NORMS
            n1 = _AffineParams(in_channels, running) if affine else _Slot()
            n2 = _AffineParams(out_channels, running) if affine else _Slot()
            mods = [n1, _ConvParams(in_channels, out_channels, k), _Slot(), n2,
                    _ConvParams(out_channels, out_channels, k), _Slot()]
        self.block = nn.Sequential(*mods)

    def forward(self, x):
        raise RuntimeError("ConvBlock holds parameters only; the network runs through the U-Net's forward")


class ConvBlock2d(ConvBlock3d):
    """Parameter holder of the 2-D block (unet.py:464-468): Conv2d weights (Cout, Cin, 3, 3)."""
    _dim = 2


class Upsampler3d(nn.Module):
    """Parameter holder for trilinear interpolate -> Conv3d(1x1x1) (unet.py:444-458, 577-581)."""

    def __init__(self, scale_factor, in_channels, out_channels, mode="trilinear"):
        super().__init__()
        if mode != "trilinear":
            raise NotImplementedError("only mode='trilinear' (the reference's Upsampler3d default)")
        self.scale_factor = scale_factor
        self.mode = mode
        self.conv = _ConvParams(in_channels, out_channels, (1, 1, 1))


class Upsampler2d(nn.Module):
    """Parameter holder for bilinear interpolate -> Conv2d(1x1) (unet.py:471-478)."""

    def __init__(self, scale_factor, in_channels, out_channels, mode="bilinear"):
        super().__init__()
        if mode != "bilinear":
            raise NotImplementedError("only mode='bilinear' (the reference's Upsampler2d default)")
        self.scale_factor = scale_factor
        self.mode = mode
        self.conv = _ConvParams(in_channels, out_channels, (1, 1))


class _Encoder(nn.Module):
    def __init__(self, features, scale_factors, kernels, block=ConvBlock3d, **kw):
        super().__init__()
        self.blocks = nn.ModuleList([block(i, o, kernel_size=k, padding=tuple(kk // 2 for kk in k), **kw)
                                     for i, o, k in zip(features[:-1], features[1:], kernels)])
        self.in_channels, self.out_channels = features[0], features[-1]

    def __len__(self):
        return len(self.blocks)


class _Decoder(nn.Module):
    def __init__(self, features, scale_factors, kernels, block=ConvBlock3d, sampler=Upsampler3d, **kw):
        super().__init__()
        self.blocks = nn.ModuleList([block(i, o, kernel_size=k, padding=tuple(kk // 2 for kk in k), **kw)
                                     for i, o, k in zip(features[:-1], features[1:], kernels)])
        self.samplers = nn.ModuleList([sampler(f, i, o) for f, i, o in zip(scale_factors, features[:-1], features[1:])])
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


def _forcing_h16(B, on):
    """backend.forcing_h16(on) where the backend has one (the emulation backend of the CPU tests does not)."""
    import contextlib
    f = getattr(B, "forcing_h16", None)
    return f(on) if f is not None else contextlib.nullcontext()


class _UNetFunction(torch.autograd.Function):
    """The whole network as one autograd node: forward_pass / backward_pass are explicit kernel schedules."""

    @staticmethod
    def forward(ctx, model, act_dtype, x, *params):
        if ctx.needs_input_grad[2]:
            raise NotImplementedError(
                "the fused U-Net node does not compute the gradient w.r.t. its input (the reference's trainer never asks for "
                "it, default_trainer.py:805-831); detach the input or set requires_grad=False")
        names = model._param_names
        # 2-D models hold Conv2d-shaped weights (Cout, Cin, kh, kw): the kernels see them as (Cout, Cin, 1, kh, kw) views
        P = {k: (v.unsqueeze(2) if v.dim() == 4 else v) for k, v in zip(names, params)}
        B = model._backend()
        # fp16 autocast (the reference trainer's default mixed precision, default_trainer.py:132-142): fp32 activations with fp16
        # tensor-core operands -- the h16 path, forced on for this node's forward and backward
        fp16 = act_dtype == torch.float16
        if fp16:
            act_dtype = torch.float32
        bf16 = act_dtype == torch.bfloat16
        two_d = model._dim == 2
        with _device_ctx(x.device), _forcing_h16(B, fp16):
            packs = model._packs(B, P)
            packs.refresh_fwd({k[:-len(".weight")]: v for k, v in P.items() if v.dim() == 5}, bf16=bf16)
            bufs = dict(model.named_buffers())
            preds, fctx = engine.forward_pass(B, model._plan, P, x.detach().unsqueeze(2) if two_d else x.detach(), act_dtype, packs,
                                              bufs=bufs, training=model.training)
        if two_d:
            preds = [p.squeeze(2) for p in preds]
        ctx.model, ctx.P, ctx.packs, ctx.bf16, ctx.fp16 = model, P, packs, bf16, fp16
        ctx.shapes = [tuple(v.shape) for v in params]
        ctx.fctx = fctx if any(ctx.needs_input_grad) else None   # no_grad / eval inference keeps nothing
        ctx.ran_backward = False
        if ctx.fctx is not None:
            # the head's backward reads the prediction (sigmoid / tanh derivative): saving it through autograd makes an
            # in-place edit of the returned tensor (pred.clamp_()) raise instead of silently corrupting the gradient
            ctx.save_for_backward(*preds)
        return tuple(preds)

    @staticmethod
    def backward(ctx, *grad_preds):
        model = ctx.model
        if ctx.fctx is None:
            if ctx.ran_backward:
                raise RuntimeError(
                    "the fused U-Net node releases its activations after the first backward pass: a second backward through "
                    "the same forward (retain_graph=True) is not supported -- run the forward again")
            raise RuntimeError("backward through a forward that ran without grad")
        two_d = model._dim == 2
        preds = [p.unsqueeze(2) if two_d else p for p in ctx.saved_tensors]
        grad_preds = [None if g is None else (g.unsqueeze(2) if two_d else g) for g in grad_preds]
        ctx.fctx.misc["preds"] = preds
        B = model._backend()
        dev = preds[0].device
        with _device_ctx(dev), _forcing_h16(B, ctx.fp16):
            ctx.packs.refresh_dgrad(bf16=ctx.bf16)
            sync = model.grad_sync
            chunked = sync is not None and hasattr(sync, "ready")
            grads = engine.backward_pass(B, model._plan, ctx.P, ctx.fctx, grad_preds, ctx.packs, sync=sync if chunked else None)
            if sync is not None and not chunked:
                sync(grads.flat)                 # plain callable: one collective over the flat gradient buffer at the end
        ctx.fctx = None
        ctx.ran_backward = True
        out = [None, None, None]
        for name, shape in zip(model._param_names, ctx.shapes):
            g = grads.get(name)
            if g is not None:
                g = g.reshape(shape)
            out.append(g)
        return tuple(out)


# ---- model-internal post-processing (unet.py:15-95): channel accumulators for bioimage.io models -------------------------
class AccumulateChannels(nn.Module):
    """``cat([x[:, i0:i1], accumulate(x[:, c0:c1], dim=1, keepdim=True)], 1)`` with accumulate in mean / min / max
    (unet.py:15-45).  A few output channels of the prediction: runs as torch indexing ops on the head's output."""

    def __init__(self, invariant_channels, accumulate_channels, accumulator):
        super().__init__()
        self.invariant_channels = invariant_channels
        self.accumulate_channels = accumulate_channels
        assert accumulator in ("mean", "min", "max")
        self.accumulator = getattr(torch, accumulator)

    def _accumulate(self, x, c0, c1):
        res = self.accumulator(x[:, c0:c1], dim=1, keepdim=True)
        if not torch.is_tensor(res):
            res = res.values
        return res

    def forward(self, x):
        c0, c1 = self.accumulate_channels
        if self.invariant_channels is None:
            return self._accumulate(x, c0, c1)
        i0, i1 = self.invariant_channels
        return torch.cat([x[:, i0:i1], self._accumulate(x, c0, c1)], dim=1)


POSTPROCESSING = {
    "affinities_to_boundaries_anisotropic": lambda: AccumulateChannels(None, (1, 3), "max"),
    "affinities_to_boundaries2d": lambda: AccumulateChannels(None, (0, 2), "max"),
    "affinities_with_foreground_to_boundaries2d": lambda: AccumulateChannels((0, 1), (1, 3), "max"),
    "affinities_to_boundaries3d": lambda: AccumulateChannels(None, (0, 3), "max"),
    "affinities_with_foreground_to_boundaries3d": lambda: AccumulateChannels((0, 1), (1, 4), "max"),
}


class _UNetCommon(nn.Module):
    """Shared machinery of UNet2d / AnisotropicUNet / UNet3d: the reference's UNetBase surface (unet.py:102-253) over the
    fused kernel schedule."""
    _dim = 3

    def _setup(self, in_channels, out_channels, sfs, initial_features, gain, final_activation, return_side_outputs,
               conv_block_impl, anisotropic_kernel, postprocessing, check_shape, conv_block_kwargs, dim):
        default_block = "ConvBlock3d" if dim == 3 else "ConvBlock2d"
        if getattr(conv_block_impl, "__name__", None) != default_block:
            raise NotImplementedError(
                f"conv_block_impl={conv_block_impl!r}: only the default {default_block} is fused; a user-defined block cannot "
                "be accelerated and there is deliberately no silent fallback")
        unknown = set(conv_block_kwargs) - {"norm", "kernel_size", "padding"}
        if unknown:
            raise TypeError(f"unexpected conv block arguments: {sorted(unknown)}")
        if conv_block_kwargs.get("kernel_size", 3) != 3 or conv_block_kwargs.get("padding", 1) != 1:
            raise NotImplementedError("only kernel_size=3, padding=1 conv blocks")
        norm = conv_block_kwargs.get("norm", "InstanceNorm")
        act_name = self._activation_name(final_activation)
        depth = len(sfs)
        if return_side_outputs:
            if isinstance(out_channels, int) or out_channels is None:
                out_channels = [out_channels] * depth
            if len(out_channels) != depth:
                raise ValueError()
        if out_channels is None or (return_side_outputs and any(c is None for c in out_channels)):
            raise NotImplementedError("out_channels=None (return decoder features) is not built on the B200 path")
        features_encoder = [in_channels] + [initial_features * gain ** i for i in range(depth)]
        features_decoder = [initial_features * gain ** i for i in range(depth + 1)][::-1]
        self._plan = engine.make_plan(in_channels, out_channels, sfs, initial_features, gain, norm, act_name, anisotropic_kernel,
                                      side_outputs=return_side_outputs, dim=dim)
        pl = self._plan
        block = ConvBlock3d if dim == 3 else ConvBlock2d
        sampler = Upsampler3d if dim == 3 else Upsampler2d

        def kk(k):
            return k if dim == 3 else k[1:]

        self.encoder = _Encoder(features_encoder, sfs, [kk(b.conv1.kernel) for b in pl.enc], block=block, norm=norm)
        self.base = block(features_encoder[-1], features_encoder[-1] * gain, kernel_size=kk(pl.base.conv1.kernel),
                          padding=tuple(v // 2 for v in kk(pl.base.conv1.kernel)), norm=norm)
        self.decoder = _Decoder(features_decoder, sfs[::-1] if dim == 3 else [2] * depth, [kk(b.conv1.kernel) for b in pl.dec],
                                block=block, sampler=sampler, norm=norm)
        one = (1,) * dim
        if return_side_outputs:
            self.out_conv = nn.ModuleList([_ConvParams(f, c, one) for f, c in zip(features_decoder[1:], out_channels)])
            self.return_decoder_outputs = True
            self._out_channels = list(out_channels)
        else:
            self.out_conv = _ConvParams(features_decoder[-1], out_channels, one)
            self.return_decoder_outputs = False
            self._out_channels = out_channels
        self.check_shape = check_shape
        self.final_activation = final_activation if isinstance(final_activation, nn.Module) else (
            None if final_activation is None else getattr(nn, final_activation)())
        self.postprocessing = self._get_postprocessing(postprocessing)
        self._param_names = [n for n, _ in self.named_parameters()]
        self._backend_override = None
        self.grad_sync = None       # callable(flat fp32 gradient buffer); set by torch_em_b200.distributed.sync_gradients
        self.compute_dtype = None   # None: follow torch.autocast (bf16) / fp32 otherwise; or torch.bfloat16 / torch.float32
        return out_channels

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

    @staticmethod
    def _get_postprocessing(postprocessing):
        if postprocessing is None:
            return None
        if isinstance(postprocessing, nn.Module):
            return postprocessing
        if postprocessing in POSTPROCESSING:
            return POSTPROCESSING[postprocessing]()
        raise ValueError(f"Invalid postprocessing: {postprocessing}")

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
        if self._dim == 2:
            factor = [2 ** self.depth] * len(spatial_shape)          # unet.py:229-235
            if len(spatial_shape) != 2:
                raise ValueError(f"Invalid shape for U-Net: dimensions don't agree {len(spatial_shape)} != 2")
            if any(sh % fac != 0 for sh, fac in zip(spatial_shape, factor)):
                raise ValueError(f"Invalid shape for U-Net: {spatial_shape} is not divisible by {factor}")
            return
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
            if dt == torch.float16 and not hasattr(self._backend(), "forcing_h16"):
                raise NotImplementedError(
                    "fp16 autocast (+GradScaler) is not supported by this backend; pass "
                    "mixed_precision_dtype='bfloat16' (or mixed_precision=False) to default_segmentation_trainer")
            return dt            # float16: served as fp32 activations with fp16 tensor-core operands (_UNetFunction.forward)
        return torch.float32

    @torch.compiler.disable
    def forward(self, x: torch.Tensor):
        """(N, in_channels, *spatial) -> (N, out_channels, *spatial) fp32, or the list of side outputs with the full-resolution
        one first (unet.py:237-253, 211-227).

        Opted out of Dynamo tracing: ``DefaultTrainer`` wraps the model in ``torch.compile`` by default
        (default_trainer.py:541, util/util.py:38-74); the network already is one hand-scheduled autograd node, so the
        compiled wrapper simply calls this method eagerly."""
        if getattr(self, "check_shape", True):
            self._check_shape(x)
        elif x.dim() != self._dim + 2:
            raise ValueError(f"Invalid shape for U-Net: dimensions don't agree {x.dim() - 2} != {self._dim}")
        params = [p for _, p in self.named_parameters()]
        outs = list(_UNetFunction.apply(self, self._activation_dtype(x), x, *params))
        if self.postprocessing is not None:
            outs = [self.postprocessing(o) for o in outs]
        return outs if self.return_decoder_outputs else outs[0]

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


class AnisotropicUNet(_UNetCommon):
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
        sfs = [engine._as_factor(sf) for sf in scale_factors]
        out_channels = self._setup(in_channels, out_channels, sfs, initial_features, gain, final_activation, return_side_outputs,
                                   conv_block_impl, anisotropic_kernel, postprocessing, check_shape, conv_block_kwargs, dim=3)
        self.init_kwargs = {"in_channels": in_channels, "out_channels": out_channels, "scale_factors": scale_factors,
                            "initial_features": initial_features, "gain": gain,
                            "final_activation": final_activation, "return_side_outputs": return_side_outputs,
                            "conv_block_impl": conv_block_impl, "anisotropic_kernel": anisotropic_kernel,
                            "postprocessing": postprocessing, **conv_block_kwargs}


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


class UNet2d(_UNetCommon):
    """2D U-Net; same arguments as ``torch_em.model.UNet2d`` (unet.py:481-563).  Runs on the 3-D kernels as their D = 1 special
    case: (1,3,3) kernels, (1,2,2) max-pooling, bilinear up-sampling = trilinear with factor 1 along depth."""
    _dim = 2

    def __init__(
        self,
        in_channels: int,
        out_channels: int,
        depth: int = 4,
        initial_features: int = 32,
        gain: int = 2,
        final_activation=None,
        return_side_outputs: bool = False,
        conv_block_impl: nn.Module = ConvBlock2d,
        pooler_impl: nn.Module = nn.MaxPool2d,
        sampler_impl: nn.Module = Upsampler2d,
        postprocessing: Optional[Union[nn.Module, str]] = None,
        check_shape: bool = True,
        **conv_block_kwargs,
    ):
        super().__init__()
        if pooler_impl is not nn.MaxPool2d:
            raise NotImplementedError(f"pooler_impl={pooler_impl!r}: only nn.MaxPool2d (the reference default) is fused")
        if getattr(sampler_impl, "__name__", None) != "Upsampler2d":
            raise NotImplementedError(f"sampler_impl={sampler_impl!r}: only Upsampler2d (bilinear, the reference default) is fused")
        sfs = [[1, 2, 2]] * depth
        out_channels = self._setup(in_channels, out_channels, sfs, initial_features, gain, final_activation, return_side_outputs,
                                   conv_block_impl, False, postprocessing, check_shape, conv_block_kwargs, dim=2)
        self.init_kwargs = {"in_channels": in_channels, "out_channels": out_channels, "depth": depth,
                            "initial_features": initial_features, "gain": gain,
                            "final_activation": final_activation, "return_side_outputs": return_side_outputs,
                            "conv_block_impl": conv_block_impl, "pooler_impl": pooler_impl,
                            "sampler_impl": sampler_impl, "postprocessing": postprocessing, **conv_block_kwargs}
