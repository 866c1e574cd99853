"""Schedule of the fused 3D U-Net train step: which kernel runs on which buffer, forward and backward.

This is the host-side mirror of ``UNetBase._apply_default`` / ``Encoder.forward`` / ``Decoder.forward`` /
``ConvBlock.forward`` / ``Upsampler.forward`` (unet.py:194-209, 311-321, 375-388, 429-441, 455-458) and of the
autograd graph PyTorch would build for them -- restated as an explicit list of calls into the C ABI
(``include/b200em.h``).  PyTorch supplies device memory (the caching allocator), the current stream and the
``autograd.Function`` hook that calls ``forward_pass`` / ``backward_pass``; no ``torch.nn`` compute op runs here.

Data layout in HBM
  * activations: channels-last ``(N, D, H, W, C)`` ("NDHWC"), bf16 under bf16 autocast, fp32 otherwise
  * every decoder level owns ONE concat buffer ``(N, D, H, W, 2C)``: the encoder's second conv writes its output (the
    skip connection) straight into channels ``[C, 2C)``, the up-sampler writes into ``[0, C)``;
    ``torch.cat`` (unet.py:372-373) never runs
  * per-(n, c) statistics ``(sum, sum of squares)`` fp32 ``[N, C, 2]`` are produced by the epilogue of whichever
    kernel wrote the tensor and turned into ``scale/shift`` by a finalize kernel; the normalisation itself is applied
    inside the consuming convolution's operand load (padding stays zero in normalised space)
  * the 1x1x1 sampler conv runs BEFORE the trilinear interpolation (they commute exactly: both linear, interpolation
    weights sum to one), i.e. on 8x fewer voxels than the reference order (unet.py:455-458)

Reference details restated here: pre-norm block order (unet.py:429-438), GroupNorm(min(32, C), C)
(unet.py:402), MaxPool3d first-max gradient routing, ReLU'(0) = 0.
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch

EPS = 1e-5  # nn.InstanceNorm3d / nn.GroupNorm default eps (unet.py:397,402)


@dataclass
class ConvSpec:
    key: str                      # state-dict prefix, e.g. "encoder.blocks.0.block.1"
    cin: int
    cout: int
    kernel: Tuple[int, int, int]


@dataclass
class BlockSpec:
    prefix: str                   # e.g. "encoder.blocks.0"
    conv1: ConvSpec
    conv2: ConvSpec
    norm1_key: Optional[str]      # GroupNorm affine parameter prefixes (None unless norm == "GroupNorm")
    norm2_key: Optional[str]


@dataclass
class Plan:
    in_channels: int
    out_channels: int
    scale_factors: List[List[int]]
    norm: Optional[str]
    final_activation: Optional[str]
    enc: List[BlockSpec] = field(default_factory=list)
    base: BlockSpec = None
    dec: List[BlockSpec] = field(default_factory=list)
    samplers: List[ConvSpec] = field(default_factory=list)
    out_conv: ConvSpec = None
    side_convs: List[ConvSpec] = None        # return_side_outputs=True: one 1x1x1 head per decoder level (unet.py:211-227)

    @property
    def depth(self):
        return len(self.scale_factors)


def _as_factor(f):
    return [f, f, f] if isinstance(f, int) else list(f)


def level_kernels(scale_factors, anisotropic_kernel):
    """Per-level kernel for one encoder / decoder, given ITS scale-factor order (unet.py:256-272, 291-294, 339-342).

    The reference shares one kwargs dict across all levels of an Encoder / Decoder and ``_update_conv_kwargs``
    mutates it in place, so the first anisotropic factor met fixes the kernel of every level (behaviour is the spec;
    pinned by tests/golden/aniso_f4_anisokernel.npz)."""
    k = (3, 3, 3)
    if anisotropic_kernel:
        for sf in scale_factors:
            f = _as_factor(sf)
            if f.count(f[0]) != len(f):
                k = tuple(1 if s == 1 else 3 for s in f)
                break
    return [k] * len(scale_factors)


AFFINE_NORMS = ("GroupNorm", "BatchNorm", "InstanceNormTrackStats")     # norms with gamma / beta parameters (unet.py:391-406)
RUNNING_NORMS = ("BatchNorm", "InstanceNormTrackStats")                 # ... and running statistics used in eval mode
NORM_MOMENTUM = {"BatchNorm": 0.1, "InstanceNormTrackStats": 0.01}      # nn.BatchNorm3d default / unet.py:398


def make_plan(in_channels, out_channels, scale_factors, initial_features=32, gain=2, norm="InstanceNorm",
              final_activation=None, anisotropic_kernel=False, side_outputs=False, dim=3) -> Plan:
    """dim=2: the 2-D U-Net (unet.py:481-563) as the D = 1 special case -- (1,3,3) kernels everywhere (base included),
    (1,2,2) pooling, and bilinear up-sampling = trilinear with factor 1 along depth."""
    sfs = [_as_factor(sf) for sf in scale_factors]
    depth = len(sfs)
    enc_f = [in_channels] + [initial_features * gain ** i for i in range(depth)]
    dec_f = [initial_features * gain ** i for i in range(depth + 1)][::-1]
    ci = (1, 4) if norm is not None else (0, 2)
    gn = norm in AFFINE_NORMS

    def block(prefix, cin, cout, k):
        return BlockSpec(prefix,
                         ConvSpec(f"{prefix}.block.{ci[0]}", cin, cout, k),
                         ConvSpec(f"{prefix}.block.{ci[1]}", cout, cout, k),
                         f"{prefix}.block.0" if gn else None, f"{prefix}.block.3" if gn else None)

    plan = Plan(in_channels, out_channels, sfs, norm, final_activation)
    if dim == 2:
        ek = dk = [(1, 3, 3)] * depth
        base_k = (1, 3, 3)
    else:
        ek = level_kernels(sfs, anisotropic_kernel)
        dk = level_kernels(sfs[::-1], anisotropic_kernel)
        base_k = (3, 3, 3)
    plan.enc = [block(f"encoder.blocks.{l}", enc_f[l], enc_f[l + 1], ek[l]) for l in range(depth)]
    plan.base = block("base", enc_f[-1], enc_f[-1] * gain, base_k)
    plan.dec = [block(f"decoder.blocks.{l}", dec_f[l], dec_f[l + 1], dk[l]) for l in range(depth)]
    plan.samplers = [ConvSpec(f"decoder.samplers.{l}.conv", dec_f[l], dec_f[l + 1], (1, 1, 1)) for l in range(depth)]
    if side_outputs:
        plan.side_convs = [ConvSpec(f"out_conv.{l}", dec_f[l + 1], out_channels[l], (1, 1, 1)) for l in range(depth)]
    elif out_channels is not None:
        plan.out_conv = ConvSpec("out_conv", dec_f[-1], out_channels, (1, 1, 1))
    return plan


def check_shape(spatial, scale_factors):
    """unet.py:671-680: every spatial axis must be divisible by the product of its scale factors."""
    factor = [1, 1, 1]
    for sf in scale_factors:
        for i in range(3):
            factor[i] *= sf[i]
    if len(spatial) != 3:
        raise ValueError(f"Invalid shape for U-Net: dimensions don't agree {len(spatial)} != 3")
    if any(sh % fac != 0 for sh, fac in zip(spatial, factor)):
        raise ValueError(f"Invalid shape for U-Net: {tuple(spatial)} is not divisible by {factor}")


def _groups(norm, c):
    return min(32, c) if norm == "GroupNorm" else c


class _ZeroArena:
    """Zero-initialised fp32 scratch of one pass (the per-(n, c) statistics the kernels accumulate into): views of a few large
    buffers, each zero-filled by ONE launch, instead of one ``torch.zeros`` -- one fill kernel -- per norm layer (42 per cfg2 step)."""
    CHUNK = 1 << 16                                      # floats per buffer (256 KB)

    def __init__(self, device):
        self.device = device
        self.buf = None
        self.used = 0

    def zeros(self, shape):
        n = 1
        for d in shape:
            n *= int(d)
        need = (n + 63) & ~63                            # 256-byte granules keep every view 16-byte aligned
        if self.buf is None or self.used + need > self.buf.numel():
            self.buf = torch.zeros(max(self.CHUNK, need), dtype=torch.float32, device=self.device)
            self.used = 0
        v = self.buf[self.used:self.used + n].view(shape)
        self.used += need
        return v


class _Ctx:
    """Everything the backward pass needs from one forward pass."""

    def __init__(self):
        self.blocks = {}
        self.misc = {}
        self.arena = None


def _norm_forward(B, plan, P, bufs, key, sums, S, C, training):
    """Per-(n, c) ``(scale, shift)`` and ``(mean, rstd)`` of one norm layer from the per-(n, c) sums of its input.

    InstanceNorm / GroupNorm: statistics of the sample (identical in train and eval mode).  BatchNorm: statistics of the whole
    batch in training mode.  BatchNorm / InstanceNormTrackStats additionally keep running statistics (updated here in training
    mode exactly like ATen: momentum update with the UNBIASED variance) and use them in eval mode (unet.py:391-406).
    The parameter-sized bookkeeping below (C or N x C floats) is done with torch ops on the device; the per-voxel work is in
    the kernels."""
    norm = plan.norm
    gamma = P[key + ".weight"] if key is not None and norm in AFFINE_NORMS else None
    beta = P[key + ".bias"] if key is not None and norm in AFFINE_NORMS else None
    N = sums.shape[0]
    if norm in RUNNING_NORMS and not training:
        rm, rv = bufs[key + ".running_mean"], bufs[key + ".running_var"]
        rstd = torch.rsqrt(rv + EPS)
        scale = rstd * gamma.detach()
        ss = torch.stack([scale, beta.detach() - rm * scale], -1)[None].expand(N, C, 2).contiguous()
        mr = torch.stack([rm, rstd], -1)[None].expand(N, C, 2).contiguous()
        return ss, mr, "fixed"
    if norm == "BatchNorm":
        sums_b = sums.sum(0, keepdim=True).expand(N, C, 2).contiguous()
        ss, mr = B.norm_finalize(sums_b, N * S, C, gamma, beta, EPS)
        cnt = N * S
        mode = "batch"
    else:
        ss, mr = B.norm_finalize(sums, S, _groups(norm, C), gamma, beta, EPS)
        cnt = S
        mode = "sample"
    if norm in RUNNING_NORMS:
        with torch.no_grad():
            m = NORM_MOMENTUM[norm]
            mean = mr[..., 0].mean(0)
            var_b = (1.0 / (mr[..., 1] * mr[..., 1]) - EPS).clamp_(min=0.0)      # biased variance back from rstd
            var_u = (var_b * (cnt / max(cnt - 1, 1))).mean(0)
            bufs[key + ".running_mean"].mul_(1 - m).add_(mean, alpha=m)
            bufs[key + ".running_var"].mul_(1 - m).add_(var_u, alpha=m)
            if norm == "BatchNorm":                     # nn.InstanceNorm3d keeps the counter but never advances it
                bufs[key + ".num_batches_tracked"].add_(1)
    return ss, mr, mode


def _norm_backward_coef(B, plan, P, key, dsums, mr, mode, S, C, grads):
    """coef[n, c] = (c0, c1, c2) with dx = c0 * g + c1 * x + c2, and the gamma / beta gradients accumulated into ``grads``."""
    norm = plan.norm
    affine = key is not None and norm in AFFINE_NORMS
    gamma = P[key + ".weight"] if affine else None
    dgamma = grads[key + ".weight"] if affine else None
    dbeta = grads[key + ".bias"] if affine else None
    N = dsums.shape[0]
    if mode == "sample":
        return B.norm_bwd_finalize(dsums, mr, gamma, S, _groups(norm, C), dgamma, dbeta)
    sg, sgx = dsums[..., 0], dsums[..., 1]
    mean, rstd = mr[..., 0], mr[..., 1]
    dgamma += (rstd * (sgx - mean * sg)).sum(0)
    dbeta += sg.sum(0)
    if mode == "fixed":                                   # running statistics are constants: dx = gamma * rstd * g
        c0 = rstd * gamma.detach()
        return torch.stack([c0, torch.zeros_like(c0), torch.zeros_like(c0)], -1).contiguous()
    dsums_b = dsums.sum(0, keepdim=True).expand(N, C, 2).contiguous()        # "batch": one statistic over (n, voxels)
    return B.norm_bwd_finalize(dsums_b, mr, gamma, N * S, C, None, None)


def _run_block(B, plan, P, bufs, spec: BlockSpec, x_in, sums_in, out, want_out_sums, ctx, packs, training):
    N, D, H, W, _ = x_in.shape
    S = D * H * W
    dev = x_in.device
    norm = plan.norm
    rec = {"x_in": x_in}
    ss1 = mr1 = ss2 = mr2 = m1 = m2 = None
    c1, c2 = spec.conv1, spec.conv2
    if norm == "InstanceNorm" and S == 1:
        # F.instance_norm refuses a single spatial element when it uses the input statistics (torch/nn/functional.py
        # _verify_spatial_size) -- nn.InstanceNorm3d without running stats always does, also in eval mode
        raise ValueError(f"Expected more than 1 spatial element when training, got input size {(N, c1.cin, D, H, W)}")
    if norm is not None:
        ss1, mr1, m1 = _norm_forward(B, plan, P, bufs, spec.norm1_key, sums_in, S, c1.cin, training)
    y1 = torch.empty((N, D, H, W, c1.cout), dtype=x_in.dtype, device=dev)
    sums1 = ctx.arena.zeros((N, c1.cout, 2)) if norm is not None else None
    aux1 = B.conv(x_in, ss1, packs[c1.key], P[c1.key + ".bias"], y1, sums1, c1.kernel, relu=True, dgrad=False)
    if norm is not None:
        ss2, mr2, m2 = _norm_forward(B, plan, P, bufs, spec.norm2_key, sums1, S, c2.cin, training)
    sums2 = None
    if want_out_sums and norm is not None:
        sums2 = ctx.arena.zeros((N, c2.cout, 2))
    aux2 = B.conv(y1, ss2, packs[c2.key], P[c2.key + ".bias"], out, sums2, c2.kernel, relu=True, dgrad=False)
    rec.update(ss1=ss1, mr1=mr1, y1=y1, ss2=ss2, mr2=mr2, y2=out, aux1=aux1, aux2=aux2, m1=m1, m2=m2)
    ctx.blocks[spec.prefix] = rec
    return sums2


def _crop_slices(full, target):
    """Decoder._crop (unet.py:363-366): centre crop of the spatial dims ``full`` to ``target``."""
    sl = []
    for f, t in zip(full, target):
        sd = (f - t) // 2
        if f - 2 * sd != t:
            # the reference's crop keeps f - 2*sd voxels, so an ODD difference ends in torch.cat's size error (unet.py:372-373)
            raise RuntimeError(f"Sizes of tensors must match except in dimension 1. Expected size {t} but got size {f - 2 * sd} "
                               "for the cropped skip connection (Decoder._crop only handles even size differences)")
        sl.append(slice(sd, f - sd))
    return tuple(sl)


def forward_pass(B, plan: Plan, P: Dict[str, torch.Tensor], x: torch.Tensor, act_dtype, packs, bufs=None, training=True):
    """x: (N, Cin, D, H, W) fp32 NCDHW on the device.  Returns (list of NCDHW fp32 predictions -- one entry, or the side
    outputs with the full-resolution one first (unet.py:211-227) --, ctx)."""
    if x.dim() != 5:
        raise ValueError(f"Invalid shape for U-Net: dimensions don't agree {x.dim() - 2} != 3")
    N, Cin, D, H, W = x.shape
    dev = x.device
    depth = plan.depth
    norm = plan.norm
    bufs = bufs or {}
    ctx = _Ctx()
    ctx.arena = _ZeroArena(dev)
    x = x.contiguous()
    if x.dtype != torch.float32:
        x = x.float()

    a = torch.empty((N, D, H, W, Cin), dtype=act_dtype, device=dev)
    B.to_ndhwc(x, a)
    cur = a
    cur_sums = None
    if norm is not None:
        cur_sums = ctx.arena.zeros((N, Cin, 2))
        B.channel_sums(a, cur_sums)
    dims = (D, H, W)
    # spatial dims on the way down (floor division, like nn.MaxPool3d) and on the way up (x factor): they differ from the
    # encoder's only for shapes the shape check would refuse (check_shape=False); the skip is then centre-cropped
    enc_dims = [dims]
    for l in range(depth):
        enc_dims.append(tuple(d // ff for d, ff in zip(enc_dims[-1], plan.scale_factors[l])))
    if any(d < 1 for d in enc_dims[-1]):
        raise ValueError(f"Invalid shape for U-Net: {(D, H, W)} is too small for {depth} pooling levels")
    dec_dims = [None] * depth
    up = enc_dims[depth]
    for lvl in reversed(range(depth)):
        up = tuple(d * ff for d, ff in zip(up, plan.scale_factors[lvl]))
        dec_dims[lvl] = up
    cats, skip_sums, skips_full = [], [], []
    for l in range(depth):
        spec = plan.enc[l]
        C = spec.conv2.cout
        dims = enc_dims[l]
        cropped = dec_dims[l] != dims
        cat = torch.empty((N,) + dec_dims[l] + (2 * C,), dtype=act_dtype, device=dev)
        skip = torch.empty((N,) + dims + (C,), dtype=act_dtype, device=dev) if cropped else cat[..., C:]
        s2 = _run_block(B, plan, P, bufs, spec, cur, cur_sums, skip, not cropped, ctx, packs, training)
        if cropped:
            sl = _crop_slices(dims, dec_dims[l])
            cat[..., C:].copy_(skip[(slice(None),) + sl])
            if norm is not None:
                s2 = ctx.arena.zeros((N, C, 2))
                B.channel_sums(cat[..., C:], s2)
        cats.append(cat)
        skip_sums.append(s2)
        skips_full.append(skip)
        f = plan.scale_factors[l]
        pooled = torch.empty((N,) + enc_dims[l + 1] + (C,), dtype=act_dtype, device=dev)
        psums = ctx.arena.zeros((N, C, 2)) if norm is not None else None
        B.maxpool_fwd(skip, pooled, f, psums)
        cur, cur_sums = pooled, psums
    spec = plan.base
    dims = enc_dims[depth]
    base_out = torch.empty((N,) + dims + (spec.conv2.cout,), dtype=act_dtype, device=dev)
    _run_block(B, plan, P, bufs, spec, cur, cur_sums, base_out, False, ctx, packs, training)
    cur = base_out
    zlows, dec_outs = [], []
    for i in range(depth):
        lvl = depth - 1 - i
        spec = plan.dec[i]
        samp = plan.samplers[i]
        C = samp.cout
        f = plan.scale_factors[lvl]
        z_low = torch.empty((N,) + dims + (C,), dtype=act_dtype, device=dev)
        B.conv(cur, None, packs[samp.key], P[samp.key + ".bias"], z_low, None, samp.kernel, relu=False, dgrad=False)
        dims = dec_dims[lvl]
        cat = cats[lvl]
        up = cat[..., :C]
        up_sums = ctx.arena.zeros((N, C, 2)) if norm is not None else None
        B.upsample_fwd(z_low, up, f, up_sums)
        zlows.append((cur, z_low))
        cat_sums = torch.cat([up_sums, skip_sums[lvl]], dim=1) if norm is not None else None
        out = torch.empty((N,) + dims + (spec.conv2.cout,), dtype=act_dtype, device=dev)
        _run_block(B, plan, P, bufs, spec, cat, cat_sums, out, False, ctx, packs, training)
        cur = out
        dec_outs.append(out)
    preds = []
    if plan.side_convs is not None:
        for i, oc in enumerate(plan.side_convs):
            y = dec_outs[i]
            pr = torch.empty((N, oc.cout) + tuple(y.shape[1:4]), dtype=torch.float32, device=dev)
            B.head_fwd(y, P[oc.key + ".weight"], P[oc.key + ".bias"], pr, plan.final_activation)
            preds.append(pr)
        preds = preds[::-1]                               # full-resolution output first (unet.py:226-227)
    else:
        if plan.out_conv is None:
            raise NotImplementedError("out_channels=None (feature output) is not on the accelerated path")
        oc = plan.out_conv
        pred = torch.empty((N, oc.cout) + tuple(cur.shape[1:4]), dtype=torch.float32, device=dev)
        B.head_fwd(cur, P[oc.key + ".weight"], P[oc.key + ".bias"], pred, plan.final_activation)
        preds = [pred]
    ctx.misc.update(cats=cats, enc_dims=enc_dims, dec_dims=dec_dims, sampler_in=zlows, dec_outs=dec_outs, act_dtype=act_dtype,
                    first_input=a, skips_full=skips_full)
    return preds, ctx


def _block_backward(B, plan, P, spec: BlockSpec, rec, dz2, need_dx, grads, packs, lazy_dx=False):
    """dz2 = gradient w.r.t. the PRE-ReLU output of conv2 (i.e. already multiplied by [y2 > 0]).
    Returns the gradient w.r.t. the block input (before any mask of the producer), or None.
    lazy_dx: return ``(g, coef)`` instead -- the raw data gradient of conv1 and the norm-backward coefficients with
    d(input) = c0 * g + c1 * x_in + c2 -- so that the consumers (up-sampling backward, max-pool backward) apply the first norm's
    backward on the fly and the block-input gradient is never written (coef is None without a norm: d(input) = g)."""
    norm = plan.norm
    c1, c2 = spec.conv1, spec.conv2
    x_in, y1 = rec["x_in"], rec["y1"]
    N, D, H, W, _ = y1.shape
    S = D * H * W
    dev = y1.device
    f32 = dict(dtype=torch.float32, device=dev)

    def dgrad_and_norm_back(dz, conv, x, mr, mode, norm_key, out, relu_mask, lazy=False):
        """g = dgrad(dz) (gradient w.r.t. the norm OUTPUT), then the norm backward (+ the producer's ReLU mask) -> out.
        The two norm-backward reductions (sum g, sum g*x) come out of the dgrad kernel's epilogue."""
        C = conv.cin
        g = torch.empty(x.shape, dtype=x.dtype, device=dev)
        if norm is None:
            B.conv(dz, None, packs[conv.key], None, g, None, conv.kernel, relu=False, dgrad=True)
            if lazy:
                return g, None
            if out is None:
                return g
            B.norm_bwd_apply(g, x, None, None, out, relu_mask)
            return out
        dsums = grads.arena.zeros((N, C, 2))
        B.conv(dz, None, packs[conv.key], None, g, dsums, conv.kernel, relu=False, dgrad=True, dot_x=x)
        coef = _norm_backward_coef(B, plan, P, norm_key, dsums, mr, mode, S, C, grads)
        if lazy:
            return g, coef
        if out is None:
            out = torch.empty_like(g)
        B.norm_bwd_apply(g, x, coef, None, out, relu_mask)
        return out

    # conv2
    B.wgrad(y1, rec["ss2"], dz2, grads[c2.key + ".weight"], grads[c2.key + ".bias"], c2.kernel, aux=rec.get("aux2"))
    dz1 = dgrad_and_norm_back(dz2, c2, y1, rec["mr2"], rec["m2"], spec.norm2_key, torch.empty_like(y1), relu_mask=1)
    # conv1
    B.wgrad(x_in, rec["ss1"], dz1, grads[c1.key + ".weight"], grads[c1.key + ".bias"], c1.kernel, aux=rec.get("aux1"))
    if not need_dx:
        return None
    return dgrad_and_norm_back(dz1, c1, x_in, rec["mr1"], rec["m1"], spec.norm1_key, None, relu_mask=0, lazy=lazy_dx)


class FlatGrads(dict):
    """Parameter gradients as views of ONE contiguous fp32 buffer (``.flat``): the weight-gradient kernels accumulate
    straight into it and the data-parallel exchange is a single all-reduce over it."""

    def __init__(self, P: Dict[str, torch.Tensor]):
        super().__init__()
        total = sum(p.numel() for p in P.values())
        dev = next(iter(P.values())).device
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.offsets = {}
        off = 0
        for k, p in P.items():
            dict.__setitem__(self, k, self.flat[off:off + p.numel()].view(p.shape))
            self.offsets[k] = (off, off + p.numel())
            off += p.numel()

    def range_of(self, prefix):
        """[lo, hi) of the parameters whose key starts with ``prefix`` (contiguous: state-dict order groups a block's tensors)."""
        r = [v for k, v in self.offsets.items() if k.startswith(prefix)]
        return (min(a for a, _ in r), max(b for _, b in r)) if r else None

    def __setitem__(self, k, v):
        self[k].copy_(v.reshape(self[k].shape))

    def setdefault(self, k, default=None):
        return self[k]


def backward_pass(B, plan: Plan, P: Dict[str, torch.Tensor], ctx: _Ctx, grad_preds, packs, sync=None):
    """grad_preds: list aligned with the predictions forward_pass returned (entries may be None for unused side outputs).
    Returns {state-dict key: fp32 gradient} (a FlatGrads) for every parameter.

    sync (optional, data-parallel training): object with ``begin(flat)``, ``ready(lo, hi)`` and ``finish()``.  ``ready`` is called
    as soon as a contiguous range of the flat gradient buffer is final, in descending address order -- first everything from the
    base block to the end (base, decoder, samplers, heads: ~95 % of the parameters, done when the most expensive, shallow encoder
    levels are still ahead), then one encoder block at a time -- so the gradient exchange overlaps the rest of the backward."""
    m = ctx.misc
    depth = plan.depth
    grads = FlatGrads(P)
    if sync is not None:
        sync.begin(grads.flat)
    preds = m["preds"]
    dec_outs = m["dec_outs"]
    dev = dec_outs[-1].device if dec_outs else preds[0].device
    grads.arena = _ZeroArena(dev)                        # zero-initialised reduction scratch of this backward pass
    if torch.is_tensor(grad_preds):
        grad_preds = [grad_preds]

    def head_back(oc, gp, pred, y, relu_mask):
        gp = gp.contiguous()
        if gp.dtype != torch.float32:
            gp = gp.float()
        dx = torch.empty_like(y)
        B.head_bwd(gp, pred, y, P[oc.key + ".weight"], dx, grads[oc.key + ".weight"], grads[oc.key + ".bias"],
                   plan.final_activation, relu_mask=relu_mask)
        return dx

    # gradient entering every decoder block's output from ITS head (side outputs) -- the last block always has one
    side = plan.side_convs is not None
    head_dx = [None] * depth
    if side:
        for i in range(depth):
            gp = grad_preds[depth - 1 - i]                 # predictions are returned full-resolution first
            if gp is not None:
                head_dx[i] = head_back(plan.side_convs[i], gp, preds[depth - 1 - i], dec_outs[i], relu_mask=1)
    elif depth > 0:
        head_dx[depth - 1] = head_back(plan.out_conv, grad_preds[0], preds[0], dec_outs[-1], relu_mask=1)
    if depth == 0:
        raise NotImplementedError("a U-Net without pooling levels is not on the accelerated path")
    dz = head_dx[depth - 1]
    if dz is None:                                         # no gradient reached the full-resolution output
        dz = torch.zeros_like(dec_outs[-1])

    skip_grads = [None] * depth
    for i in reversed(range(depth)):
        lvl = depth - 1 - i
        spec, samp = plan.dec[i], plan.samplers[i]
        C = samp.cout
        rec = ctx.blocks[spec.prefix]
        cat = rec["x_in"]
        f = plan.scale_factors[lvl]
        x_low, z_low = m["sampler_in"][i]
        d_zlow = torch.empty(z_low.shape, dtype=x_low.dtype, device=dev)
        # The gradient w.r.t. the concat buffer is consumed twice -- [..., :C] by the up-sampling backward, [..., C:] (the skip)
        # by the max-pool backward of the encoder -- and both kernels apply the block's first norm backward themselves:
        # d_cat = c0 * g + c1 * cat + c2 is then never written (the up-sampling backward does it on the low-resolution tensor by
        # linearity, see b200em_upsample_trilinear_bwd).  Cropped skips, odd factors: materialise it as before.
        lazy = B.fused_up_bwd_ok(cat[..., :C], f) and m["enc_dims"][lvl] == m["dec_dims"][lvl]
        if lazy:
            g_cat, coef = _block_backward(B, plan, P, spec, rec, dz, True, grads, packs, lazy_dx=True)
            skip_grads[lvl] = (g_cat[..., C:], None if coef is None else coef[:, C:])
            B.upsample_bwd(g_cat[..., :C], d_zlow, f, zlow=z_low, coef=None if coef is None else coef[:, :C])
        else:
            d_cat = _block_backward(B, plan, P, spec, rec, dz, True, grads, packs)
            skip_grads[lvl] = (d_cat[..., C:], None)
            B.upsample_bwd(d_cat[..., :C], d_zlow, f)
        B.wgrad(x_low, None, d_zlow, grads[samp.key + ".weight"], grads[samp.key + ".bias"], samp.kernel)
        g = torch.empty_like(x_low)
        B.conv(d_zlow, None, packs[samp.key], None, g, None, samp.kernel, relu=False, dgrad=True)
        # x_low is the post-ReLU output of the block below: add the gradient from that block's own head (side outputs), then
        # apply its ReLU mask in place
        add = head_dx[i - 1] if i > 0 else None
        B.norm_bwd_apply(g, x_low, None, add, g, 1)
        dz = g
    need_first_dx = plan.norm in AFFINE_NORMS          # the first norm's gamma/beta need the gradient w.r.t. its output
    d_p = _block_backward(B, plan, P, plan.base, ctx.blocks["base"], dz, depth > 0 or need_first_dx, grads, packs)
    if sync is not None:
        sync.ready(grads.range_of("base.")[0], grads.flat.numel())
    for l in reversed(range(depth)):
        spec = plan.enc[l]
        rec = ctx.blocks[spec.prefix]
        skip = rec["y2"]
        dz = torch.empty(skip.shape, dtype=skip.dtype, device=dev)
        f = plan.scale_factors[l]
        sg, sg_coef = skip_grads[l]
        dims, ddims = m["enc_dims"][l], m["dec_dims"][l]
        if ddims != dims:
            # check_shape=False with a non-divisible shape: the skip was centre-cropped (Decoder._crop), so its gradient is
            # zero outside the crop; voxels beyond the last pooling window get no pooled gradient either
            full = torch.zeros(skip.shape, dtype=skip.dtype, device=dev)
            full[(slice(None),) + _crop_slices(dims, ddims)] = sg
            B.norm_bwd_apply(full, skip, None, None, dz, 1)
            sg = full
        B.maxpool_bwd(skip, d_p, sg, dz, f, 1, coef=sg_coef)
        d_p = _block_backward(B, plan, P, spec, rec, dz, l > 0 or need_first_dx, grads, packs)
        if sync is not None:
            sync.ready(*grads.range_of(spec.prefix + "."))
    if sync is not None:
        sync.finish()
    return grads
