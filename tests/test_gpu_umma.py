"""GPU parity of the tcgen05 implicit-GEMM convolution (csrc/conv_umma.cu) against the plain-PyTorch contract
(tests/emu_backend.py) with the SAME bf16-rounded operands, so the only differences are fp32 accumulation order and
the final bf16 rounding of the stored output (one bf16 ulp = 2^-8 relative)."""
import numpy as np
import pytest
import torch

from tests.emu_backend import TorchEmuBackend

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
EMU = TorchEmuBackend()


class P:
    def __init__(self, w):
        self.w = w


@pytest.fixture(scope="module", params=["dstacked", "plain"])
def B(request):
    """The tensor-core conv variants: depth-stacked (conv_umma_ds.cu, Cout <= 80 with a resident filter) and the plain one."""
    from torch_em_b200 import _lib
    from torch_em_b200.backend import CudaBackend
    _lib.load()
    return CudaBackend(use_ds=request.param == "dstacked", use_cs=request.param == "dstacked")


CASES = [
    (1, 4, 16, 8, 32, 32, (1, 1, 1)),
    (1, 4, 16, 8, 32, 32, (3, 3, 3)),
    (2, 5, 20, 13, 32, 64, (3, 3, 3)),
    (1, 6, 16, 16, 16, 16, (3, 3, 3)),
    (1, 3, 16, 8, 64, 32, (1, 3, 3)),
    (1, 9, 33, 17, 48, 80, (3, 3, 3)),
    (1, 4, 8, 8, 128, 256, (3, 3, 3)),
    (1, 2, 8, 8, 256, 512, (3, 3, 3)),
    (3, 8, 32, 32, 32, 32, (3, 3, 3)),
    (1, 5, 9, 30, 64, 64, (3, 3, 3)),
    (2, 3, 17, 14, 32, 16, (1, 3, 3)),
    (1, 4, 24, 43, 16, 80, (3, 3, 3)),
    (2, 37, 16, 16, 32, 32, (3, 3, 3)),
    (1, 70, 17, 9, 64, 32, (3, 3, 3)),
    (1, 21, 16, 8, 32, 64, (3, 3, 1)),
    (1, 19, 20, 11, 16, 48, (3, 1, 3)),
    (1, 15, 16, 8, 16, 80, (3, 3, 3)),
    (1, 11, 18, 12, 32, 48, (3, 3, 3)),
    (2, 9, 16, 24, 48, 16, (3, 3, 3)),
]


def rnd(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g) * scale


@pytest.mark.parametrize("case", CASES)
def test_umma_conv_forward_and_dgrad(B, case):
    from torch_em_b200 import _lib
    N, D, H, W, Cin, Cout, k = case
    assert _lib.load().b200em_conv3d_umma_supported(Cin, Cout, *k)
    x = rnd((N, D, H, W, Cin), 1).bfloat16()
    w = rnd((Cout, Cin) + k, 2, scale=(Cin * k[0] * k[1] * k[2]) ** -0.5)
    wq = w.bfloat16().float()                       # the operand the tensor cores see
    b = rnd((Cout,), 3)
    ss = torch.stack([1 + 0.1 * rnd((N, Cin), 4), 0.1 * rnd((N, Cin), 5)], -1).contiguous()
    pk = B.pack(("umma-test", case), w.to(DEV))
    assert pk.umma_fwd is not None or pk.ds_fwd is not None
    assert pk.umma_dgrad is not None or pk.ds_dgrad is not None
    if B.use_ds and _lib.load().b200em_conv3d_umma_ds_supported(Cin, Cout, *k):
        assert pk.ds_fwd is not None
    for in_ss, relu, bias in ((None, False, None), (ss, True, b)):
        y_ref = torch.empty((N, D, H, W, Cout), dtype=torch.bfloat16)
        s_ref = torch.zeros((N, Cout, 2))
        xin = x
        if in_ss is not None:                       # the loader rounds the normalised operand to bf16
            xin = (x.float() * in_ss[:, None, None, None, :, 0] + in_ss[:, None, None, None, :, 1]).bfloat16()
        EMU.conv(xin, None, P(wq), bias, y_ref, s_ref, k, relu, False)
        ybuf = torch.zeros((N, D, H, W, Cout + 16), dtype=torch.bfloat16, device=DEV)
        y = ybuf[..., 16:]
        s = torch.zeros((N, Cout, 2), device=DEV)
        B.conv(x.to(DEV), None if in_ss is None else in_ss.to(DEV), pk, None if bias is None else bias.to(DEV), y, s, k, relu, False)
        torch.cuda.synchronize()
        np.testing.assert_allclose(y.float().cpu().numpy(), y_ref.float().numpy(), rtol=1e-2, atol=1e-2)
        np.testing.assert_allclose(s.cpu().numpy(), s_ref.numpy(), rtol=5e-3, atol=0.5)
        assert float(ybuf[..., :16].abs().max()) == 0.0
    dz = rnd((N, D, H, W, Cout), 6).bfloat16()
    g_ref = torch.empty((N, D, H, W, Cin), dtype=torch.bfloat16)
    EMU.conv(dz, None, P(wq), None, g_ref, None, k, False, True)
    g = torch.empty((N, D, H, W, Cin), dtype=torch.bfloat16, device=DEV)
    B.conv(dz.to(DEV), None, pk, None, g, None, k, False, True)
    torch.cuda.synchronize()
    np.testing.assert_allclose(g.float().cpu().numpy(), g_ref.float().numpy(), rtol=1e-2, atol=1e-2)
    # data gradient with the fused norm-backward reductions: sums = (sum g, sum g * x) over the stored g
    d_ref = torch.zeros((N, Cin, 2))
    EMU.channel_dot_sums(g_ref, x, d_ref)
    d = torch.zeros((N, Cin, 2), device=DEV)
    B.conv(dz.to(DEV), None, pk, None, g, d, k, False, True, dot_x=x.to(DEV))
    torch.cuda.synchronize()
    np.testing.assert_allclose(d.cpu().numpy(), d_ref.numpy(), rtol=1e-2, atol=2e-2 * float(d_ref.abs().max()))


@pytest.mark.parametrize("tma", ["1", "0"])
@pytest.mark.parametrize("ks", ["1", "2", "3", "7"])
@pytest.mark.parametrize("case", [(1, 4, 8, 8, 128, 256, (3, 3, 3)), (2, 6, 13, 9, 96, 48, (3, 3, 3)), (1, 2, 4, 4, 512, 128, (1, 3, 3)),
                                  (1, 8, 8, 8, 256, 128, (1, 1, 1))])
def test_plain_conv_split_k(case, ks, tma, monkeypatch):
    """Split-K of the plain kernel (layers with fewer output tiles than SMs): ks CTAs reduce disjoint Cin-chunk ranges of a tile
    through the registered workspace, the last one runs the epilogue (bias, ReLU, statistics / norm-backward reductions).  Forced
    factors incl. ones that do not divide the chunk count; the workspace must be all-zero again afterwards."""
    from torch_em_b200.backend import CudaBackend
    monkeypatch.setenv("B200EM_KSPLIT", ks)
    monkeypatch.setenv("B200EM_UMMA_TMA", tma)       # "0": the cp.async fallback of the plain kernel's tile loader
    B = CudaBackend(use_ds=False, use_cs=False)
    N, D, H, W, Cin, Cout, k = case
    x = rnd((N, D, H, W, Cin), 1).bfloat16()
    w = rnd((Cout, Cin) + k, 2, scale=(Cin * k[0] * k[1] * k[2]) ** -0.5)
    wq = w.bfloat16().float()
    b = rnd((Cout,), 3)
    ss = torch.stack([1 + 0.1 * rnd((N, Cin), 4), 0.1 * rnd((N, Cin), 5)], -1).contiguous()
    pk = B.pack(("splitk-test", case), w.to(DEV))
    xin = (x.float() * ss[:, None, None, None, :, 0] + ss[:, None, None, None, :, 1]).bfloat16()
    y_ref = torch.empty((N, D, H, W, Cout), dtype=torch.bfloat16)
    s_ref = torch.zeros((N, Cout, 2))
    EMU.conv(xin, None, P(wq), b, y_ref, s_ref, k, True, False)
    for _ in range(2):                               # twice: the second run sees the workspace the first one left behind
        y = torch.empty((N, D, H, W, Cout), dtype=torch.bfloat16, device=DEV)
        s = torch.zeros((N, Cout, 2), device=DEV)
        B.conv(x.to(DEV), ss.to(DEV), pk, b.to(DEV), y, s, k, True, False)
        torch.cuda.synchronize()
        np.testing.assert_allclose(y.float().cpu().numpy(), y_ref.float().numpy(), rtol=1e-2, atol=1e-2)
        np.testing.assert_allclose(s.cpu().numpy(), s_ref.numpy(), rtol=5e-3, atol=0.5)
        assert not bool(B._workspaces[0].any()), "split-K must leave the workspace all-zero"
    dz = rnd((N, D, H, W, Cout), 6).bfloat16()
    g_ref = torch.empty((N, D, H, W, Cin), dtype=torch.bfloat16)
    EMU.conv(dz, None, P(wq), None, g_ref, None, k, False, True)
    d_ref = torch.zeros((N, Cin, 2))
    EMU.channel_dot_sums(g_ref, x, d_ref)
    g = torch.empty((N, D, H, W, Cin), dtype=torch.bfloat16, device=DEV)
    d = torch.zeros((N, Cin, 2), device=DEV)
    B.conv(dz.to(DEV), None, pk, None, g, d, k, False, True, dot_x=x.to(DEV))
    torch.cuda.synchronize()
    np.testing.assert_allclose(g.float().cpu().numpy(), g_ref.float().numpy(), rtol=1e-2, atol=1e-2)
    np.testing.assert_allclose(d.cpu().numpy(), d_ref.numpy(), rtol=1e-2, atol=2e-2 * float(d_ref.abs().max()))
    assert not bool(B._workspaces[0].any())


@pytest.mark.parametrize("tma", ["1", "0"])
@pytest.mark.parametrize("case", [(2, 24, 48, 40, 64, 32, (3, 3, 3)), (2, 24, 48, 40, 32, 32, (3, 3, 3))])
def test_dstacked_many_items_per_cta(case, tma, monkeypatch):
    """Depth-stacked kernel with more work items than SMs (short depth segments forced): the operand ring, the accumulator
    ring and the dot_x prefetch ring all wrap many times per CTA, and the dgrad runs with fewer ring stages than loader warps."""
    from torch_em_b200 import _lib
    from torch_em_b200.backend import CudaBackend
    monkeypatch.setenv("B200EM_DS_DR", "3")
    monkeypatch.setenv("B200EM_DS_TMA", tma)         # "0": the cp.async fallback of the tile loader
    N, D, H, W, Cin, Cout, k = case
    Bd = CudaBackend(use_ds=True)
    assert _lib.load().b200em_conv3d_umma_ds_supported(Cin, Cout, *k) and _lib.load().b200em_conv3d_umma_ds_supported(Cout, Cin, *k)
    x = rnd((N, D, H, W, Cin), 31).bfloat16()
    w = rnd((Cout, Cin) + k, 32, scale=(Cin * 27) ** -0.5)
    wq = w.bfloat16().float()
    b = rnd((Cout,), 33)
    ss = torch.stack([1 + 0.1 * rnd((N, Cin), 34), 0.1 * rnd((N, Cin), 35)], -1).contiguous()
    pk = Bd.pack(("ds-many", case), w.to(DEV))
    xin = (x.float() * ss[:, None, None, None, :, 0] + ss[:, None, None, None, :, 1]).bfloat16()
    y_ref = torch.empty((N, D, H, W, Cout), dtype=torch.bfloat16); s_ref = torch.zeros((N, Cout, 2))
    EMU.conv(xin, None, P(wq), b, y_ref, s_ref, k, True, False)
    y = torch.empty((N, D, H, W, Cout), dtype=torch.bfloat16, device=DEV); s = torch.zeros((N, Cout, 2), device=DEV)
    for _ in range(3):                              # repeated launches: a ring race shows up as a mismatch or a trapped kernel
        s.zero_()
        Bd.conv(x.to(DEV), ss.to(DEV), pk, b.to(DEV), y, s, k, True, False)
    torch.cuda.synchronize()
    np.testing.assert_allclose(y.float().cpu().numpy(), y_ref.float().numpy(), rtol=1e-2, atol=1e-2)
    np.testing.assert_allclose(s.cpu().numpy(), s_ref.numpy(), rtol=5e-3, atol=0.5)
    dz = rnd((N, D, H, W, Cout), 36).bfloat16()
    g_ref = torch.empty((N, D, H, W, Cin), dtype=torch.bfloat16)
    EMU.conv(dz, None, P(wq), None, g_ref, None, k, False, True)
    d_ref = torch.zeros((N, Cin, 2))
    EMU.channel_dot_sums(g_ref, x, d_ref)
    g = torch.empty((N, D, H, W, Cin), dtype=torch.bfloat16, device=DEV); d = torch.zeros((N, Cin, 2), device=DEV)
    for _ in range(3):
        d.zero_()
        Bd.conv(dz.to(DEV), None, pk, None, g, d, k, False, True, dot_x=x.to(DEV))
    torch.cuda.synchronize()
    np.testing.assert_allclose(g.float().cpu().numpy(), g_ref.float().numpy(), rtol=1e-2, atol=1e-2)
    np.testing.assert_allclose(d.cpu().numpy(), d_ref.numpy(), rtol=1e-2, atol=2e-2 * float(d_ref.abs().max()))


WG_CASES = [
    (1, 2, 16, 8, 32, 32, (1, 1, 1)),
    (1, 2, 16, 8, 32, 32, (3, 3, 3)),
    (2, 5, 20, 13, 32, 64, (3, 3, 3)),
    (1, 3, 16, 8, 64, 32, (1, 3, 3)),
    (1, 7, 33, 17, 96, 48, (3, 3, 3)),
    (1, 4, 8, 8, 128, 256, (3, 3, 3)),
    (3, 8, 32, 32, 32, 32, (3, 3, 3)),
    (2, 37, 16, 16, 32, 32, (3, 3, 3)),
    (1, 21, 20, 11, 64, 64, (3, 1, 3)),
    (2, 24, 48, 40, 64, 32, (3, 3, 3)),
    (1, 9, 33, 17, 96, 64, (3, 3, 3)),
    # 1x1x1 filters (the up-sampler convs): 64 / 128 input channels per CTA instead of 32 (M rows = channels of one slice)
    (2, 5, 20, 13, 64, 32, (1, 1, 1)),
    (1, 7, 17, 9, 128, 64, (1, 1, 1)),
    (2, 3, 8, 8, 512, 256, (1, 1, 1)),
    (1, 9, 33, 17, 192, 48, (1, 1, 1)),
]


@pytest.mark.parametrize("case", WG_CASES)
def test_umma_wgrad(B, case):
    from torch_em_b200 import _lib
    N, D, H, W, Cin, Cout, k = case
    assert _lib.load().b200em_conv3d_wgrad_umma_supported(Cin, Cout, *k)
    x = rnd((N, D, H, W, Cin), 11).bfloat16()
    dz = rnd((N, D, H, W, Cout), 12).bfloat16()
    ss = torch.stack([1 + 0.1 * rnd((N, Cin), 13), 0.1 * rnd((N, Cin), 14)], -1).contiguous()
    for in_ss in (None, ss):
        xin = x if in_ss is None else (x.float() * in_ss[:, None, None, None, :, 0] + in_ss[:, None, None, None, :, 1]).bfloat16()
        dw_ref, db_ref = torch.zeros((Cout, Cin) + k), torch.zeros(Cout)
        EMU.wgrad(xin, None, dz, dw_ref, db_ref, k)
        dw = torch.full((Cout, Cin) + k, 0.5, device=DEV)          # accumulate semantics: += on top of existing values
        db = torch.full((Cout,), -1.0, device=DEV)
        B.wgrad(x.to(DEV), None if in_ss is None else in_ss.to(DEV), dz.to(DEV), dw, db, k)
        torch.cuda.synchronize()
        np.testing.assert_allclose(dw.cpu().numpy() - 0.5, dw_ref.numpy(), rtol=2e-3, atol=2e-3 * float(dw_ref.abs().max()))
        np.testing.assert_allclose(db.cpu().numpy() + 1.0, db_ref.numpy(), rtol=2e-3, atol=2e-3 * float(db_ref.abs().max()))


@pytest.mark.parametrize("tma", ["1", "0"])
@pytest.mark.parametrize("dr", ["2", "5"])
def test_wgrad_cs_many_items_per_cta(dr, tma, monkeypatch):
    """w-stacked weight gradient with short depth segments forced: several work items per CTA, the stage ring and its
    mirrored slots wrap many times; repeated launches must give the same answer."""
    from torch_em_b200.backend import CudaBackend
    monkeypatch.setenv("B200EM_CS_DR", dr)
    monkeypatch.setenv("B200EM_CS_TMA", tma)         # "0": the cp.async fallback of the tile loader
    N, D, H, W, Cin, Cout, k = 2, 24, 48, 40, 64, 32, (3, 3, 3)
    Bc = CudaBackend(use_cs=True)
    x = rnd((N, D, H, W, Cin), 41).bfloat16()
    dz = rnd((N, D, H, W, Cout), 42).bfloat16()
    ss = torch.stack([1 + 0.1 * rnd((N, Cin), 43), 0.1 * rnd((N, Cin), 44)], -1).contiguous()
    xin = (x.float() * ss[:, None, None, None, :, 0] + ss[:, None, None, None, :, 1]).bfloat16()
    dw_ref, db_ref = torch.zeros((Cout, Cin) + k), torch.zeros(Cout)
    EMU.wgrad(xin, None, dz, dw_ref, db_ref, k)
    for _ in range(3):
        dw = torch.zeros((Cout, Cin) + k, device=DEV)
        db = torch.zeros((Cout,), device=DEV)
        Bc.wgrad(x.to(DEV), ss.to(DEV), dz.to(DEV), dw, db, k)
        torch.cuda.synchronize()
        np.testing.assert_allclose(dw.cpu().numpy(), dw_ref.numpy(), rtol=2e-3, atol=2e-3 * float(dw_ref.abs().max()))
        np.testing.assert_allclose(db.cpu().numpy(), db_ref.numpy(), rtol=2e-3, atol=2e-3 * float(db_ref.abs().max()))


@pytest.mark.parametrize("case", [(2, 6, 20, 13, 1, 32, (3, 3, 3)), (1, 4, 16, 16, 2, 16, (3, 3, 3)), (1, 3, 18, 9, 1, 64, (1, 3, 3)),
                                  (1, 40, 33, 17, 1, 16, (3, 3, 3)), (2, 24, 48, 40, 1, 64, (3, 3, 3)), (1, 5, 16, 8, 1, 48, (3, 3, 3))])
def test_thin_k_first_conv(B, case):
    """First conv (Cin <= 4): im2col + 1x1x1 tcgen05 conv, forward and weight gradient."""
    N, D, H, W, Cin, Cout, k = case
    x = rnd((N, D, H, W, Cin), 21).bfloat16()
    w = rnd((Cout, Cin) + k, 22, scale=0.2)
    b = rnd((Cout,), 23)
    ss = torch.stack([1 + 0.1 * rnd((N, Cin), 24), 0.1 * rnd((N, Cin), 25)], -1).contiguous()
    pk = B.pack(("thin-test", case), w.to(DEV))
    from torch_em_b200 import _lib
    first = B.use_ds and _lib.load().b200em_conv3d_first_supported(Cin, Cout, *k)   # on-the-fly im2col kernel (Cin = 1, 3x3x3)
    assert pk.thin is not None or first
    xin = (x.float() * ss[:, None, None, None, :, 0] + ss[:, None, None, None, :, 1]).bfloat16()
    y_ref = torch.empty((N, D, H, W, Cout), dtype=torch.bfloat16); s_ref = torch.zeros((N, Cout, 2))
    EMU.conv(xin, None, P(w.bfloat16().float()), b, y_ref, s_ref, k, True, False)
    y = torch.empty((N, D, H, W, Cout), dtype=torch.bfloat16, device=DEV); s = torch.zeros((N, Cout, 2), device=DEV)
    B.conv(x.to(DEV), ss.to(DEV), pk, b.to(DEV), y, s, k, True, False)
    torch.cuda.synchronize()
    np.testing.assert_allclose(y.float().cpu().numpy(), y_ref.float().numpy(), rtol=1e-2, atol=1e-2)
    np.testing.assert_allclose(s.cpu().numpy(), s_ref.numpy(), rtol=5e-3, atol=0.5)
    dz = rnd((N, D, H, W, Cout), 26).bfloat16()
    dw_ref, db_ref = torch.zeros((Cout, Cin) + k), torch.zeros(Cout)
    EMU.wgrad(xin, None, dz, dw_ref, db_ref, k)
    dw, db = torch.zeros((Cout, Cin) + k, device=DEV), torch.zeros(Cout, device=DEV)
    B.wgrad(x.to(DEV), ss.to(DEV), dz.to(DEV), dw, db, k)
    torch.cuda.synchronize()
    np.testing.assert_allclose(dw.cpu().numpy(), dw_ref.numpy(), rtol=2e-3, atol=2e-3 * float(dw_ref.abs().max()))
    np.testing.assert_allclose(db.cpu().numpy(), db_ref.numpy(), rtol=2e-3, atol=2e-3 * float(db_ref.abs().max()))
