"""TF32 tensor-core path for fp32 activations (csrc/conv_umma.cu, kind::tf32): what the reference's ``mixed_precision=False``
training uses on a GPU, because torch runs fp32 convolutions in TF32 by default (torch.backends.cudnn.allow_tf32;
default_trainer.py:132-142, BASELINE.json configs[3]).

TF32 keeps 10 mantissa bits of every operand, so there is no exact answer: the yardstick is float64 / exact fp32, and the bar is
the error of the reference's OWN TF32 run (the same functional graph through cuDNN with allow_tf32=True on this GPU)."""
import numpy as np
import pytest
import torch

import torch_em_b200 as tb
from oracle import dice as odice
from oracle import unet as ounet
from tests.emu_backend import TorchEmuBackend

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
EMU = TorchEmuBackend()


class P:
    def __init__(self, w):
        self.w = w


@pytest.fixture()
def tf32():
    torch.backends.cudnn.allow_tf32 = True           # (the autouse fixture of conftest.py restores the previous value)
    yield


def rnd(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g) * scale


CASES = [
    (1, 4, 16, 8, 32, 32, (3, 3, 3)),
    (2, 5, 20, 13, 16, 64, (3, 3, 3)),
    (1, 3, 16, 8, 64, 32, (1, 3, 3)),
    (1, 4, 8, 8, 128, 256, (3, 3, 3)),
    (1, 6, 17, 9, 48, 80, (3, 3, 3)),
    (2, 3, 9, 16, 64, 128, (1, 1, 1)),
    (1, 9, 33, 17, 32, 16, (3, 3, 3)),
]


@pytest.mark.parametrize("case", CASES)
def test_tf32_conv_forward_dgrad_wgrad(tf32, case):
    from torch_em_b200 import _lib
    from torch_em_b200.backend import CudaBackend
    N, D, H, W, Cin, Cout, k = case
    B = CudaBackend()
    assert B.tf32_enabled() and _lib.load().b200em_conv3d_umma_tf32_supported(Cin, Cout, *k)
    x = rnd((N, D, H, W, Cin), 1)
    w = rnd((Cout, Cin) + k, 2, scale=(Cin * k[0] * k[1] * k[2]) ** -0.5)
    b = rnd((Cout,), 3)
    ss = torch.stack([1 + 0.1 * rnd((N, Cin), 4), 0.1 * rnd((N, Cin), 5)], -1).contiguous()
    pk = B.pack(("tf32", case), w.to(DEV))
    assert pk.tf32_fwd is not None and pk.tf32_dgrad is not None
    tol = 3e-3                                          # TF32: 2^-11 per operand, random accumulation
    for in_ss, relu, bias in ((None, False, None), (ss, True, b)):
        y_ref = torch.empty((N, D, H, W, Cout))
        s_ref = torch.zeros((N, Cout, 2))
        EMU.conv(x, in_ss, P(w), bias, y_ref, s_ref, k, relu, False)
        ybuf = torch.zeros((N, D, H, W, Cout + 8), device=DEV)
        y = ybuf[..., 8:]
        s = torch.zeros((N, Cout, 2), device=DEV)
        B.calls.clear()
        B.conv(x.to(DEV), None if in_ss is None else in_ss.to(DEV), pk, None if bias is None else bias.to(DEV), y, s, k, relu, False)
        torch.cuda.synchronize()
        assert dict(B.calls) == {"tf32:fwd": 1}
        np.testing.assert_allclose(y.cpu().numpy(), y_ref.numpy(), rtol=tol, atol=tol * float(y_ref.abs().max()))
        np.testing.assert_allclose(s.cpu().numpy(), s_ref.numpy(), rtol=5e-3, atol=5e-3 * float(s_ref.abs().max()))
        assert float(ybuf[..., :8].abs().max()) == 0.0
    dz = rnd((N, D, H, W, Cout), 6)
    g_ref = torch.empty((N, D, H, W, Cin))
    EMU.conv(dz, None, P(w), None, g_ref, None, k, False, True)
    d_ref = torch.zeros((N, Cin, 2))
    EMU.channel_dot_sums(g_ref, x, d_ref)
    g = torch.empty((N, D, H, W, Cin), device=DEV)
    d = torch.zeros((N, Cin, 2), device=DEV)
    B.conv(dz.to(DEV), None, pk, None, g, d, k, False, True, dot_x=x.to(DEV))
    torch.cuda.synchronize()
    assert B.calls["tf32:dgrad"] == 1
    np.testing.assert_allclose(g.cpu().numpy(), g_ref.numpy(), rtol=tol, atol=tol * float(g_ref.abs().max()))
    np.testing.assert_allclose(d.cpu().numpy(), d_ref.numpy(), rtol=1e-2, atol=1e-2 * float(d_ref.abs().max()))
    if Cin % 32 == 0:
        dw_ref, db_ref = torch.zeros_like(w), torch.zeros(Cout)
        EMU.wgrad(x, ss, dz, dw_ref, db_ref, k)
        dw, db = torch.zeros_like(w).to(DEV), torch.zeros(Cout, device=DEV)
        B.wgrad(x.to(DEV), ss.to(DEV), dz.to(DEV), dw, db, k)
        torch.cuda.synchronize()
        assert B.calls["split3:wgrad"] == 1
        np.testing.assert_allclose(dw.cpu().numpy(), dw_ref.numpy(), rtol=tol, atol=tol * float(dw_ref.abs().max()))
        np.testing.assert_allclose(db.cpu().numpy(), db_ref.numpy(), rtol=1e-4, atol=1e-4 * float(db_ref.abs().max()))


def test_exact_fp32_when_tf32_is_disallowed():
    """torch.backends.cudnn.allow_tf32 = False (the test default, conftest.py) selects the exact CUDA-core kernels."""
    from torch_em_b200.backend import CudaBackend
    B = CudaBackend()
    assert not B.tf32_enabled()
    x, w = rnd((1, 4, 16, 8, 32), 1), rnd((32, 32, 3, 3, 3), 2, 0.03)
    pk = B.pack("k", w.to(DEV))
    y = torch.empty((1, 4, 16, 8, 32), device=DEV)
    B.conv(x.to(DEV), None, pk, None, y, None, (3, 3, 3), False, False)
    assert dict(B.calls) == {"direct:fwd": 1}
    y_ref = torch.empty((1, 4, 16, 8, 32))
    EMU.conv(x, None, P(w), None, y_ref, None, (3, 3, 3), False, False)
    np.testing.assert_allclose(y.cpu().numpy(), y_ref.numpy(), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("kw,shape", [
    (dict(in_channels=1, out_channels=2, depth=3, initial_features=32, final_activation="Sigmoid"), (2, 1, 32, 32, 32)),
    # configs[3] topology (depth 5, boundary-style 2-channel output) at quarter width
    (dict(in_channels=1, out_channels=2, depth=5, initial_features=16, final_activation="Sigmoid"), (1, 1, 64, 64, 64)),
])
def test_model_fp32_tf32_no_worse_than_reference_tf32(tf32, kw, shape):
    from torch_em_b200.backend import default_backend
    torch.manual_seed(0)
    net = tb.UNet3d(**kw).to(DEV)
    depth = kw["depth"]
    x = torch.randn(*shape)
    t = (torch.nn.functional.avg_pool3d(torch.randn(shape[0], 2, *shape[2:]), 5, 1, 2) > 0).float()     # smooth, learnable targets

    def oracle(dtype, dev):
        sd = {k: v.detach().to(dev).to(dtype).clone().requires_grad_(True) for k, v in net.state_dict().items()}
        y_ = ounet.unet3d_forward(x.to(dev).to(dtype), sd, [2] * depth, final_activation="Sigmoid")
        l_ = odice.dice_loss(y_, t.to(dev).to(dtype))
        l_.backward()
        return y_.detach().cpu().double(), l_.item(), {k: v.grad.cpu().double() for k, v in sd.items()}

    y64, l64, g64 = oracle(torch.float64, "cpu")
    y_tf, l_tf, g_tf = oracle(torch.float32, DEV)              # the reference's arithmetic: cuDNN with TF32 allowed
    B = default_backend()
    B.calls.clear()
    y = net(x.to(DEV))
    loss = tb.DiceLoss()(y, t.to(DEV))
    loss.backward()
    torch.cuda.synchronize()
    assert B.calls.get("tf32:fwd", 0) > 0 and B.calls.get("tf32:dgrad", 0) > 0 and B.calls.get("split3:wgrad", 0) > 0, dict(B.calls)
    assert y.dtype == torch.float32

    def rel(a, b):
        return float((a - b).norm() / (b.norm() + 1e-30))

    e_y, e_ref = rel(y.detach().cpu().double(), y64), rel(y_tf, y64)
    assert e_y < 5e-3 and e_y <= 2.0 * e_ref + 1e-4, (e_y, e_ref)
    assert abs(loss.item() - l64) < 2e-3 * abs(l64)
    gmax = max(float(v.norm()) for v in g64.values())
    bad = []
    for k, p in net.named_parameters():
        e_o = float((p.grad.cpu().double() - g64[k]).norm()) / gmax
        e_r = float((g_tf[k] - g64[k]).norm()) / gmax
        if e_o > 2.0 * e_r + 2e-3:
            bad.append((k, e_o, e_r))
    assert not bad, bad
