"""Oracle: functional fp32 restatement of torch_em.model.unet.{UNet3d,AnisotropicUNet} (TEST INFRASTRUCTURE).

Not a copy of the reference module tree: one flat function over a ``state_dict`` keyed exactly like the
reference's, using ``torch.nn.functional`` ops on whatever device the tensors live on
(CPU fp32 for parity; the same function on ``cuda`` is the cuDNN comparison arm of ``bench.py``).

Reference lines restated:
  * block order  Norm -> Conv3d -> ReLU -> Norm -> Conv3d -> ReLU      unet.py:409-441
  * norms        InstanceNorm3d(C) / GroupNorm(min(32,C),C) / None     unet.py:391-406
  * encoder      block, keep skip, MaxPool3d(factor)                   unet.py:311-321
  * decoder      trilinear(align_corners=False) -> 1x1x1 conv -> cat([up, skip]) -> block   unet.py:375-388,444-458
  * head         out_conv (1x1x1) -> final activation                  unet.py:194-209
  * anisotropic kernels (1,3,3)/(0,1,1): see level_kernels()            unet.py:256-272
  * shape check                                                        unet.py:229-235, 671-680
"""
from typing import Dict, List, Optional, Sequence, Union

import numpy as np
import torch
import torch.nn.functional as F


def _as_factor(f) -> List[int]:
    return [f, f, f] if isinstance(f, int) else list(f)


def feature_schedule(in_channels, scale_factors, initial_features=32, gain=2):
    depth = len(scale_factors)
    enc = [in_channels] + [initial_features * gain ** i for i in range(depth)]
    dec = [initial_features * gain ** i for i in range(depth + 1)][::-1]
    return enc, dec


def level_kernels(scale_factors, anisotropic_kernel: bool):
    """Per-level (kernel, padding) for one Encoder / Decoder given ITS scale-factor order (unet.py:256-272,
    291-294, 339-342).

    The reference builds ``[conv_block_kwargs] * n`` (one shared dict) and ``_update_conv_kwargs`` mutates it in
    place, so the first anisotropic scale factor met in this order fixes the kernel of EVERY level of that
    encoder / decoder (later anisotropic factors bail out because kernel_size is then a tuple).  Verified by
    instantiating the reference: AnisotropicUNet([[1,2,2],[2,2,2]], anisotropic_kernel=True) has (1,3,3)
    kernels in all encoder and decoder blocks and (3,3,3) only in the base.  Behaviour is the spec.
    """
    k, p = (3, 3, 3), (1, 1, 1)
    if anisotropic_kernel:
        for sf in scale_factors:
            f = _as_factor(sf)
            if f.count(f[0]) != len(f):
                k = tuple(1 if s == 1 else 3 for s in f)
                p = tuple(0 if s == 1 else 1 for s in f)
                break
    return [(k, p)] * len(scale_factors)


def init_state_dict(in_channels, out_channels, scale_factors, initial_features=32, gain=2,
                    norm="InstanceNorm", anisotropic_kernel=False, seed=0, dtype=torch.float32):
    """Random weights with the reference's key names and shapes (kaiming-uniform-like bounds)."""
    g = torch.Generator().manual_seed(seed)
    enc, dec = feature_schedule(in_channels, scale_factors, initial_features, gain)
    sd = {}
    conv_idx = (1, 4) if norm is not None else (0, 2)

    def conv(prefix, cout, cin, k):
        fan_in = cin * k[0] * k[1] * k[2]
        bound = 1.0 / np.sqrt(fan_in)
        sd[prefix + ".weight"] = (torch.rand(cout, cin, *k, generator=g, dtype=dtype) * 2 - 1) * bound
        sd[prefix + ".bias"] = (torch.rand(cout, generator=g, dtype=dtype) * 2 - 1) * bound

    def block(prefix, cin, cout, k):
        if norm == "GroupNorm":
            sd[f"{prefix}.block.0.weight"] = 1 + 0.1 * torch.randn(cin, generator=g, dtype=dtype)
            sd[f"{prefix}.block.0.bias"] = 0.1 * torch.randn(cin, generator=g, dtype=dtype)
            sd[f"{prefix}.block.3.weight"] = 1 + 0.1 * torch.randn(cout, generator=g, dtype=dtype)
            sd[f"{prefix}.block.3.bias"] = 0.1 * torch.randn(cout, generator=g, dtype=dtype)
        conv(f"{prefix}.block.{conv_idx[0]}", cout, cin, k)
        conv(f"{prefix}.block.{conv_idx[1]}", cout, cout, k)

    depth = len(scale_factors)
    ek = level_kernels(scale_factors, anisotropic_kernel)
    for l in range(depth):
        block(f"encoder.blocks.{l}", enc[l], enc[l + 1], ek[l][0])
    block("base", enc[-1], enc[-1] * gain, (3, 3, 3))
    rev = list(scale_factors)[::-1]
    dk = level_kernels(rev, anisotropic_kernel)
    for l in range(depth):
        block(f"decoder.blocks.{l}", dec[l], dec[l + 1], dk[l][0])
    for l in range(depth):
        conv(f"decoder.samplers.{l}.conv", dec[l + 1], dec[l], (1, 1, 1))
    if out_channels is not None:
        conv("out_conv", out_channels, dec[-1], (1, 1, 1))
    return sd


def _norm(x, sd, key, norm):
    if norm is None:
        return x
    if norm == "InstanceNorm":
        return F.instance_norm(x, eps=1e-5)
    if norm == "GroupNorm":
        c = x.shape[1]
        return F.group_norm(x, min(32, c), sd[key + ".weight"], sd[key + ".bias"], eps=1e-5)
    raise ValueError(f"oracle: norm {norm!r} not on the in-scope path")


def _block(x, sd, prefix, norm, pad):
    ci = (1, 4) if norm is not None else (0, 2)
    x = _norm(x, sd, f"{prefix}.block.0", norm)
    x = F.relu(F.conv3d(x, sd[f"{prefix}.block.{ci[0]}.weight"], sd[f"{prefix}.block.{ci[0]}.bias"], padding=pad))
    x = _norm(x, sd, f"{prefix}.block.3", norm)
    x = F.relu(F.conv3d(x, sd[f"{prefix}.block.{ci[1]}.weight"], sd[f"{prefix}.block.{ci[1]}.bias"], padding=pad))
    return x


def check_shape(spatial_shape: Sequence[int], scale_factors):
    """unet.py:671-680 (same message)."""
    factor = [int(np.prod([_as_factor(sf)[i] for sf in scale_factors])) for i in range(3)]
    if len(spatial_shape) != 3:
        raise ValueError(f"Invalid shape for U-Net: dimensions don't agree {len(spatial_shape)} != 3")
    if any(sh % fac != 0 for sh, fac in zip(spatial_shape, factor)):
        raise ValueError(f"Invalid shape for U-Net: {tuple(spatial_shape)} is not divisible by {factor}")


def unet3d_forward(x: torch.Tensor, sd: Dict[str, torch.Tensor], scale_factors,
                   norm: Optional[str] = "InstanceNorm", final_activation: Optional[str] = None,
                   anisotropic_kernel: bool = False, return_intermediates: bool = False):
    """Forward of UNet3d / AnisotropicUNet on NCDHW ``x`` with reference-keyed ``sd``."""
    check_shape(x.shape[2:], scale_factors)
    depth = len(scale_factors)
    inter = {}
    skips = []
    ek = level_kernels(scale_factors, anisotropic_kernel)
    for l in range(depth):
        x = _block(x, sd, f"encoder.blocks.{l}", norm, ek[l][1])
        skips.append(x)
        f = _as_factor(scale_factors[l])
        x = F.max_pool3d(x, kernel_size=f, stride=f)
        inter[f"enc{l}"] = skips[-1]
    x = _block(x, sd, "base", norm, (1, 1, 1))
    inter["base"] = x
    rev = list(scale_factors)[::-1]
    dk = level_kernels(rev, anisotropic_kernel)
    for l in range(depth):
        f = _as_factor(rev[l])
        x = F.interpolate(x, scale_factor=[float(s) for s in f], mode="trilinear", align_corners=False)
        x = F.conv3d(x, sd[f"decoder.samplers.{l}.conv.weight"], sd[f"decoder.samplers.{l}.conv.bias"])
        x = torch.cat([x, skips[depth - 1 - l]], dim=1)
        x = _block(x, sd, f"decoder.blocks.{l}", norm, dk[l][1])
        inter[f"dec{l}"] = x
    if "out_conv.weight" in sd:
        x = F.conv3d(x, sd["out_conv.weight"], sd["out_conv.bias"])
    if final_activation is not None:
        if final_activation == "Sigmoid":
            x = torch.sigmoid(x)
        elif final_activation == "ReLU":
            x = F.relu(x)
        elif final_activation == "Tanh":
            x = torch.tanh(x)
        else:
            raise ValueError(f"oracle: activation {final_activation!r} not restated")
    if return_intermediates:
        return x, inter
    return x


def conv_flops_fwd(in_channels, out_channels, scale_factors, spatial, batch, initial_features=32, gain=2,
                   anisotropic_kernel=False):
    """Algorithmic conv FLOPs of one forward: sum 2*Nvox*Cin*Cout*taps."""
    enc, dec = feature_schedule(in_channels, scale_factors, initial_features, gain)
    depth = len(scale_factors)
    sp = list(spatial)
    total = 0
    first = None

    def add(cin, cout, taps, sp_):
        nonlocal total, first
        fl = 2 * batch * sp_[0] * sp_[1] * sp_[2] * cin * cout * taps
        if first is None:
            first = fl
        total += fl

    ek = level_kernels(scale_factors, anisotropic_kernel)
    for l in range(depth):
        k = ek[l][0]
        taps = k[0] * k[1] * k[2]
        add(enc[l], enc[l + 1], taps, sp)
        add(enc[l + 1], enc[l + 1], taps, sp)
        f = _as_factor(scale_factors[l])
        sp = [s // ff for s, ff in zip(sp, f)]
    add(enc[-1], enc[-1] * gain, 27, sp)
    add(enc[-1] * gain, enc[-1] * gain, 27, sp)
    rev = list(scale_factors)[::-1]
    dk = level_kernels(rev, anisotropic_kernel)
    for l in range(depth):
        f = _as_factor(rev[l])
        sp = [s * ff for s, ff in zip(sp, f)]
        add(dec[l], dec[l + 1], 1, sp)          # sampler 1x1x1 (at high res in the reference)
        k = dk[l][0]
        taps = k[0] * k[1] * k[2]
        add(dec[l], dec[l + 1], taps, sp)
        add(dec[l + 1], dec[l + 1], taps, sp)
    if out_channels is not None:
        add(dec[-1], out_channels, 1, sp)
    return total, first


def conv_flops_train(*args, **kwargs):
    """F_train = 3*F_fwd - F_fwd(first conv) (no dgrad into the network input)."""
    total, first = conv_flops_fwd(*args, **kwargs)
    return 3 * total - first
