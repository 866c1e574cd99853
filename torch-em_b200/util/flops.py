"""Algorithmic convolution FLOPs of the U-Net train step -- the figure ``roofline.achieved`` uses.

``F_fwd = sum over convs of 2 * voxels * Cin * Cout * taps`` in the REFERENCE's order (the 1x1x1 sampler conv counted at
high resolution, unet.py:455-458, although this implementation runs it before the interpolation);
``F_train = 3 * F_fwd - F_fwd(first conv)`` (no data gradient into the network input).
"""
from ..engine import _as_factor, level_kernels


def conv_flops_fwd(in_channels, out_channels, scale_factors, spatial, batch, initial_features=32, gain=2,
                   anisotropic_kernel=False):
    """-> (total forward FLOPs, FLOPs of the first conv)."""
    sfs = [_as_factor(sf) for sf in scale_factors]
    depth = len(sfs)
    enc = [in_channels] + [initial_features * gain ** i for i in range(depth)]
    dec = [initial_features * gain ** i for i in range(depth + 1)][::-1]
    sp = list(spatial)
    convs = []                                           # (cin, cout, taps, voxels)

    def vox():
        return batch * sp[0] * sp[1] * sp[2]

    ek = level_kernels(sfs, anisotropic_kernel)
    for l in range(depth):
        taps = ek[l][0] * ek[l][1] * ek[l][2]
        convs += [(enc[l], enc[l + 1], taps, vox()), (enc[l + 1], enc[l + 1], taps, vox())]
        sp = [s // f for s, f in zip(sp, sfs[l])]
    convs += [(enc[-1], enc[-1] * gain, 27, vox()), (enc[-1] * gain, enc[-1] * gain, 27, vox())]
    rev = sfs[::-1]
    dk = level_kernels(rev, anisotropic_kernel)
    for l in range(depth):
        sp = [s * f for s, f in zip(sp, rev[l])]
        taps = dk[l][0] * dk[l][1] * dk[l][2]
        convs += [(dec[l], dec[l + 1], 1, vox()), (dec[l], dec[l + 1], taps, vox()), (dec[l + 1], dec[l + 1], taps, vox())]
    if out_channels is not None:
        convs.append((dec[-1], out_channels, 1, vox()))
    fl = [2 * v * ci * co * t for ci, co, t, v in convs]
    return sum(fl), fl[0]


def conv_flops_train(*args, **kwargs):
    total, first = conv_flops_fwd(*args, **kwargs)
    return 3 * total - first
