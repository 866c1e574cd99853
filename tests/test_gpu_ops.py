"""GPU parity tests, kernel by kernel: every C-ABI entry point (called through torch-em_b200/backend.py) against the
plain-PyTorch statement of its contract (tests/emu_backend.py, run on CPU in fp32) on the same seeded inputs.

Tolerances: fp32 kernels vs fp32 CPU -- rtol 1e-4 / atol 1e-5 (reduction-order noise only).  bf16
kernels compute in fp32 from bf16 inputs and round once on store: one bf16 ulp (2^-8 relative) on outputs, fp32
tolerances on statistics and parameter gradients computed from the same bf16 inputs.
"""
import numpy as np
import pytest
import torch

from tests.emu_backend import TorchEmuBackend

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(scope="module")
def B():
    from torch_em_b200 import _lib
    from torch_em_b200.backend import CudaBackend
    _lib.load()
    return CudaBackend(use_umma=False)      # this module pins the direct / elementwise kernels; tcgen05: test_gpu_umma.py


EMU = TorchEmuBackend()


def act(shape, dtype, seed, scale=1.0, relu=False):
    g = torch.Generator().manual_seed(seed)
    t = torch.randn(shape, generator=g) * scale
    if relu:
        t = t.clamp(min=0)
    return t.to(dtype)


def tol(dtype):
    return dict(rtol=1e-4, atol=1e-5) if dtype == torch.float32 else dict(rtol=1e-2, atol=1e-2)


def close(a, b, **kw):
    np.testing.assert_allclose(a.detach().float().cpu().numpy(), b.detach().float().cpu().numpy(), **kw)


class P:
    def __init__(self, w):
        self.w = w


CONV_CASES = [
    # N, D, H, W, Cin, Cout, kernel
    (2, 6, 9, 11, 1, 8, (3, 3, 3)),
    (1, 8, 8, 8, 5, 7, (3, 3, 3)),
    (2, 4, 10, 12, 16, 40, (1, 3, 3)),
    (1, 5, 6, 7, 12, 4, (1, 1, 1)),
    (1, 9, 17, 8, 32, 32, (3, 3, 3)),
]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_forward_dgrad_wgrad(B, case, dtype):
    N, D, H, W, Cin, Cout, k = case
    x = act((N, D, H, W, Cin), dtype, 1)
    w = act((Cout, Cin) + k, torch.float32, 2, scale=0.2)
    b = act((Cout,), torch.float32, 3)
    ss = torch.stack([1 + 0.1 * act((N, Cin), torch.float32, 4), 0.1 * act((N, Cin), torch.float32, 5)], -1).contiguous()
    dz = act((N, D, H, W, Cout), dtype, 6)
    wd = w.to(DEV)
    pk = B.pack(("test", case, str(dtype)), wd)
    for in_ss, relu in ((None, False), (ss, True)):
        y_ref = torch.empty((N, D, H, W, Cout), dtype=dtype)
        s_ref = torch.zeros((N, Cout, 2))
        EMU.conv(x, in_ss, P(w), b, y_ref, s_ref, k, relu, False)
        # output written into a channel slice of a wider buffer (concat-buffer path)
        ybuf = torch.zeros((N, D, H, W, Cout + 8), dtype=dtype, device=DEV)
        y = ybuf[..., 8:]
        s = torch.zeros((N, Cout, 2), device=DEV)
        B.conv(x.to(DEV), None if in_ss is None else in_ss.to(DEV), pk, b.to(DEV), y, s, k, relu, False)
        close(y, y_ref, **tol(dtype))
        close(s, s_ref, rtol=2e-3 if dtype == torch.bfloat16 else 1e-4, atol=0.3 if dtype == torch.bfloat16 else 1e-3)
        assert float(ybuf[..., :8].abs().max()) == 0.0
    # data gradient
    g_ref = torch.empty((N, D, H, W, Cin), dtype=dtype)
    EMU.conv(dz, None, P(w), None, g_ref, None, k, False, True)
    g = torch.empty((N, D, H, W, Cin), dtype=dtype, device=DEV)
    B.conv(dz.to(DEV), None, pk, None, g, None, k, False, True)
    close(g, g_ref, **tol(dtype))
    # weight gradient (with the fused norm apply on x)
    dw_ref, db_ref = torch.zeros_like(w), torch.zeros(Cout)
    EMU.wgrad(x, ss, dz, dw_ref, db_ref, k)
    dw, db = torch.zeros_like(wd), torch.zeros(Cout, device=DEV)
    B.wgrad(x.to(DEV), ss.to(DEV), dz.to(DEV), dw, db, k)
    close(dw, dw_ref, rtol=1e-3, atol=1e-3 * float(dw_ref.abs().max()))
    close(db, db_ref, rtol=1e-3, atol=1e-3 * float(db_ref.abs().max()))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("groups_of", ["instance", "group"])
def test_norm_statistics_and_backward(B, dtype, groups_of):
    N, D, H, W, C = 2, 5, 6, 7, 16
    S = D * H * W
    groups = C if groups_of == "instance" else 4
    x = act((N, D, H, W, C), dtype, 1, relu=True) + 0.5
    g = act((N, D, H, W, C), dtype, 2)
    gamma = (1 + 0.1 * act((C,), torch.float32, 3)) if groups_of == "group" else None
    beta = (0.1 * act((C,), torch.float32, 4)) if groups_of == "group" else None
    s_ref = torch.zeros((N, C, 2)); EMU.channel_sums(x, s_ref)
    s = torch.zeros((N, C, 2), device=DEV); B.channel_sums(x.to(DEV), s)
    close(s, s_ref, rtol=1e-4, atol=1e-3)
    ss_ref, mr_ref = EMU.norm_finalize(s_ref, S, groups, gamma, beta, 1e-5)
    ss, mr = B.norm_finalize(s, S, groups, None if gamma is None else gamma.to(DEV), None if beta is None else beta.to(DEV), 1e-5)
    close(ss, ss_ref, rtol=1e-4, atol=1e-5)
    close(mr, mr_ref, rtol=1e-4, atol=1e-5)
    d_ref = torch.zeros((N, C, 2)); EMU.channel_dot_sums(g, x, d_ref)
    d = torch.zeros((N, C, 2), device=DEV); B.channel_dot_sums(g.to(DEV), x.to(DEV), d)
    close(d, d_ref, rtol=1e-4, atol=1e-3)
    dg_ref, db_ref = (torch.zeros(C), torch.zeros(C)) if gamma is not None else (None, None)
    coef_ref = EMU.norm_bwd_finalize(d_ref, mr_ref, gamma, S, groups, dg_ref, db_ref)
    dg, db = (torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)) if gamma is not None else (None, None)
    coef = B.norm_bwd_finalize(d, mr, None if gamma is None else gamma.to(DEV), S, groups, dg, db)
    close(coef, coef_ref, rtol=1e-3, atol=1e-5)
    if gamma is not None:
        close(dg, dg_ref, rtol=1e-3, atol=1e-3)
        close(db, db_ref, rtol=1e-3, atol=1e-3)
    add = act((N, D, H, W, C), dtype, 5)
    for c_, a_, relu in ((coef_ref, None, 1), (None, add, 0), (coef_ref, add, 1), (None, None, 1)):
        o_ref = torch.empty((N, D, H, W, C), dtype=dtype)
        EMU.norm_bwd_apply(g, x, c_, a_, o_ref, relu)
        o = torch.empty((N, D, H, W, C), dtype=dtype, device=DEV)
        B.norm_bwd_apply(g.to(DEV), x.to(DEV), None if c_ is None else c_.to(DEV), None if a_ is None else a_.to(DEV), o, relu)
        close(o, o_ref, **tol(dtype))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("f", [(2, 2, 2), (1, 2, 2)])
@pytest.mark.parametrize("C", [8, 3])
def test_pool_and_upsample(B, dtype, f, C):
    N, D, H, W = 2, 4, 6, 8
    # few distinct values -> many ties: the first maximum in (d, h, w) scan order must receive the gradient
    gen = torch.Generator().manual_seed(7)
    x = (torch.randint(0, 3, (N, D, H, W, C), generator=gen).float() * 0.5).to(dtype)
    Do, Ho, Wo = D // f[0], H // f[1], W // f[2]
    y_ref = torch.empty((N, Do, Ho, Wo, C), dtype=dtype); s_ref = torch.zeros((N, C, 2))
    EMU.maxpool_fwd(x, y_ref, f, s_ref)
    y = torch.empty((N, Do, Ho, Wo, C), dtype=dtype, device=DEV); s = torch.zeros((N, C, 2), device=DEV)
    B.maxpool_fwd(x.to(DEV), y, f, s)
    close(y, y_ref, rtol=0, atol=0)
    close(s, s_ref, rtol=1e-5, atol=1e-4)
    dp = act((N, Do, Ho, Wo, C), dtype, 8)
    add = act((N, D, H, W, C), dtype, 9)
    for a_, relu in ((None, 0), (add, 1)):
        o_ref = torch.empty((N, D, H, W, C), dtype=dtype)
        EMU.maxpool_bwd(x, dp, a_, o_ref, f, relu)
        o = torch.empty((N, D, H, W, C), dtype=dtype, device=DEV)
        B.maxpool_bwd(x.to(DEV), dp.to(DEV), None if a_ is None else a_.to(DEV), o, f, relu)
        close(o, o_ref, **tol(dtype))
    # trilinear, align_corners=False
    z = act((N, Do, Ho, Wo, C), dtype, 10)
    u_ref = torch.empty((N, D, H, W, C), dtype=dtype); us_ref = torch.zeros((N, C, 2))
    EMU.upsample_fwd(z, u_ref, f, us_ref)
    u = torch.empty((N, D, H, W, C), dtype=dtype, device=DEV); us = torch.zeros((N, C, 2), device=DEV)
    B.upsample_fwd(z.to(DEV), u, f, us)
    close(u, u_ref, **tol(dtype))
    close(us, us_ref, rtol=2e-3, atol=0.2 if dtype == torch.bfloat16 else 1e-3)
    du = act((N, D, H, W, C), dtype, 11)
    dz_ref = torch.empty((N, Do, Ho, Wo, C), dtype=dtype)
    EMU.upsample_bwd(du, dz_ref, f)
    dz = torch.empty((N, Do, Ho, Wo, C), dtype=dtype, device=DEV)
    B.upsample_bwd(du.to(DEV), dz, f)
    close(dz, dz_ref, rtol=1e-4 if dtype == torch.float32 else 2e-2, atol=1e-5 if dtype == torch.float32 else 3e-2)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("f,shape", [((2, 2, 2), (2, 5, 9, 11)), ((1, 2, 2), (1, 3, 18, 9)), ((2, 2, 2), (1, 1, 1, 1))])
@pytest.mark.parametrize("C", [16, 40])
def test_fused_norm_backward_in_pool_and_upsample_backward(B, dtype, f, shape, C):
    """The decoder block's first norm backward (d = c0 * g + c1 * x + c2 per (n, c)) applied on the fly by the tiled up-sampling
    backward and by the max-pool backward, on channel SLICES of a 2C-wide concat buffer, vs the materialised form.  Odd tile
    counts and a single-voxel volume exercise the clamped tile borders."""
    N, D, H, W = shape
    Dh, Hh, Wh = D * f[0], H * f[1], W * f[2]
    g_cat = act((N, Dh, Hh, Wh, 2 * C), dtype, 21)
    cat = act((N, Dh, Hh, Wh, 2 * C), dtype, 22, relu=True)
    coef = torch.stack([1 + 0.2 * act((N, 2 * C), torch.float32, 23), 0.3 * act((N, 2 * C), torch.float32, 24),
                        0.1 * act((N, 2 * C), torch.float32, 25)], -1).contiguous()
    t = tol(dtype) if dtype == torch.float32 else dict(rtol=2e-2, atol=6e-2)
    # up-sampling backward of the first C channels; `up` = the up-sampled low-resolution tensor (what the concat buffer holds)
    zlow = act((N, D, H, W, C), dtype, 27)
    dz_ref = torch.empty((N, D, H, W, C), dtype=dtype)
    EMU.upsample_bwd(g_cat[..., :C], dz_ref, f, zlow=zlow, coef=coef[:, :C])
    gd, cd, kd = g_cat.to(DEV), cat.to(DEV), coef.to(DEV)
    assert B.fused_up_bwd_ok(gd[..., :C], f) == (256 % (C // (8 if dtype == torch.bfloat16 else 4)) == 0)
    dz = torch.empty((N, D, H, W, C), dtype=dtype, device=DEV)
    if B.fused_up_bwd_ok(gd[..., :C], f):
        B.upsample_bwd(gd[..., :C], dz, f, zlow=zlow.to(DEV), coef=kd[:, :C])
        close(dz, dz_ref, **t)
    dz2_ref = torch.empty((N, D, H, W, C), dtype=dtype)
    EMU.upsample_bwd(g_cat[..., :C], dz2_ref, f)
    B.upsample_bwd(gd[..., :C], dz, f)
    close(dz, dz2_ref, **t)
    # max-pool backward of the skip half: pooled gradient + (c0 * g + c1 * skip + c2), ReLU mask of the skip
    dp = act((N, D, H, W, C), dtype, 26)
    o_ref = torch.empty((N, Dh, Hh, Wh, C), dtype=dtype)
    EMU.maxpool_bwd(cat[..., C:], dp, g_cat[..., C:], o_ref, f, 1, coef=coef[:, C:])
    o = torch.empty((N, Dh, Hh, Wh, C), dtype=dtype, device=DEV)
    B.maxpool_bwd(cd[..., C:], dp.to(DEV), gd[..., C:], o, f, 1, coef=kd[:, C:])
    close(o, o_ref, **t)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("act_name", [None, "Sigmoid", "ReLU", "Tanh"])
@pytest.mark.parametrize("Cin,Cout", [(16, 2), (5, 3), (32, 12), (32, 2), (64, 2), (128, 1), (64, 1), (32, 3), (64, 12), (128, 8), (32, 6), (64, 4)])
def test_head(B, dtype, act_name, Cin, Cout):
    N, D, H, W = (2, 3, 5, 9) if Cin < 64 else (2, 7, 13, 19)       # the wide heads: several warps per channel slice
    x = act((N, D, H, W, Cin), dtype, 1, relu=True)
    w = act((Cout, Cin, 1, 1, 1), torch.float32, 2, scale=0.3)
    b = act((Cout,), torch.float32, 3)
    o_ref = torch.empty((N, Cout, D, H, W)); EMU.head_fwd(x, w, b, o_ref, act_name)
    o = torch.empty((N, Cout, D, H, W), device=DEV)
    B.head_fwd(x.to(DEV), w.to(DEV), b.to(DEV), o, act_name)
    close(o, o_ref, rtol=1e-4, atol=1e-5)
    go = act((N, Cout, D, H, W), torch.float32, 4)
    dx_ref = torch.empty((N, D, H, W, Cin), dtype=dtype); dw_ref = torch.zeros_like(w); db_ref = torch.zeros(Cout)
    EMU.head_bwd(go, o_ref, x, w, dx_ref, dw_ref, db_ref, act_name, 1)
    dx = torch.empty((N, D, H, W, Cin), dtype=dtype, device=DEV); dw = torch.zeros_like(w, device=DEV); db = torch.zeros(Cout, device=DEV)
    B.head_bwd(go.to(DEV), o, x.to(DEV), w.to(DEV), dx, dw, db, act_name, 1)
    close(dx, dx_ref, **tol(dtype))
    close(dw, dw_ref, rtol=1e-3, atol=1e-3)
    close(db, db_ref, rtol=1e-3, atol=1e-3)


def test_layout_and_memset(B):
    x = act((2, 3, 4, 5, 6), torch.float32, 1)
    for dtype in (torch.float32, torch.bfloat16):
        y = torch.empty((2, 4, 5, 6, 3), dtype=dtype, device=DEV)
        B.to_ndhwc(x.to(DEV), y)
        close(y, x.permute(0, 2, 3, 4, 1).to(dtype), rtol=0, atol=0)
    import ctypes
    from torch_em_b200 import _lib
    buf = torch.ones(1000003, dtype=torch.uint8, device=DEV)
    _lib.call("b200em_memset_zero", ctypes.c_void_p(buf.data_ptr()), buf.numel(), None)
    torch.cuda.synchronize()
    assert int(buf.sum()) == 0
    assert _lib.launch_count() > 0
