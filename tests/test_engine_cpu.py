"""Host logic of the product on CPU: the kernel schedule (torch-em_b200/engine.py) driven through a plain-PyTorch
emulation of the C ABI (tests/emu_backend.py) must reproduce the reference's forward, loss and every parameter
gradient stored in tests/golden/*.npz (generated from the reference itself by tests/golden/make_golden.py)."""
import copy
import os

import numpy as np
import pytest
import torch

import torch_em_b200 as tb
from oracle import dice as odice
from tests.emu_backend import TorchEmuBackend

CASES = {
    "unet3d_d2_f4_instnorm": ("UNet3d", dict(in_channels=1, out_channels=2, depth=2, initial_features=4,
                                             final_activation="Sigmoid")),
    "unet3d_d2_f8_groupnorm": ("UNet3d", dict(in_channels=2, out_channels=3, depth=2, initial_features=8,
                                              final_activation="Sigmoid", norm="GroupNorm")),
    "unet3d_d1_f32_groupnorm": ("UNet3d", dict(in_channels=1, out_channels=2, depth=1, initial_features=32,
                                               final_activation="Sigmoid", norm="GroupNorm")),
    "unet3d_d1_f4_nonorm": ("UNet3d", dict(in_channels=1, out_channels=1, depth=1, initial_features=4,
                                           final_activation=None, norm=None)),
    "aniso_f4_anisokernel": ("AnisotropicUNet", dict(in_channels=1, out_channels=3,
                                                     scale_factors=[[1, 2, 2], [2, 2, 2]], initial_features=4,
                                                     final_activation="Sigmoid", anisotropic_kernel=True)),
    "aniso_f4_isokernel": ("AnisotropicUNet", dict(in_channels=1, out_channels=2,
                                                   scale_factors=[[1, 2, 2], [2, 2, 2]], initial_features=4,
                                                   final_activation="Sigmoid", anisotropic_kernel=False)),
}


def build(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    ctor, kw = CASES[name]
    net = getattr(tb, ctor)(**kw)
    sd = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w:")}
    missing, unexpected = net.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    return z, net


@pytest.mark.parametrize("name", sorted(CASES))
def test_schedule_matches_reference(golden_dir, name):
    z, net = build(golden_dir, name)
    net._backend_override = TorchEmuBackend()
    y = net(torch.from_numpy(z["x"]))
    assert y.dtype == torch.float32 and y.requires_grad
    np.testing.assert_allclose(y.detach().numpy(), z["y"], rtol=1e-4, atol=1e-5)
    loss = odice.dice_loss(y, torch.from_numpy(z["t"]))
    np.testing.assert_allclose(loss.item(), z["loss"], rtol=1e-5)
    loss.backward()
    for k, p in net.named_parameters():
        g = z["g:" + k]
        assert p.grad is not None, k
        np.testing.assert_allclose(p.grad.numpy(), g, rtol=2e-3, atol=2e-6 + 1e-4 * np.abs(g).max(), err_msg=k)


def test_state_dict_keys_match_reference(golden_dir):
    for name in CASES:
        z, net = build(golden_dir, name)
        ref_keys = [k[2:] for k in z.files if k.startswith("w:")]
        assert list(net.state_dict().keys()) == ref_keys
        for k, v in net.state_dict().items():
            assert tuple(v.shape) == z["w:" + k].shape


def test_constructor_surface():
    net = tb.UNet3d(1, 2, depth=3, initial_features=4, final_activation="Sigmoid")
    assert net.in_channels == 1 and net.out_channels == 2 and net.depth == 3
    assert set(net.init_kwargs) == {"in_channels", "out_channels", "depth", "initial_features", "gain", "final_activation",
                                    "return_side_outputs", "conv_block_impl", "postprocessing"}
    an = tb.AnisotropicUNet(1, 2, scale_factors=[[1, 2, 2], [2, 2, 2]], initial_features=4, anisotropic_kernel=True)
    assert "scale_factors" in an.init_kwargs and an.init_kwargs["anisotropic_kernel"] is True
    # re-creatable from init_kwargs, deep-copyable (predict_with_halo, prediction.py:188-192), picklable class path
    again = tb.UNet3d(**net.init_kwargs)
    assert [k for k in again.state_dict()] == [k for k in net.state_dict()]
    cp = copy.deepcopy(net)
    assert all(torch.equal(a, b) for a, b in zip(cp.state_dict().values(), net.state_dict().values()))
    assert f"{type(net).__module__}.{type(net).__name__}" == "torch_em_b200.model.unet.UNet3d"
    with pytest.raises(ValueError, match="Invalid shape for U-Net"):          # test/model/test_unet.py:19-23
        net._backend_override = TorchEmuBackend()
        net(torch.zeros(1, 1, 12, 16, 16))
    with pytest.raises(ValueError, match="Invalid activation"):
        tb.UNet3d(1, 1, depth=1, final_activation="NoSuchActivation")
    # what cannot be fused is refused loudly instead of falling back silently
    class MyBlock(torch.nn.Module):
        pass
    for kw in (dict(conv_block_impl=MyBlock), dict(out_channels=None), dict(final_activation="Softmax"), dict(kernel_size=5, padding=2)):
        with pytest.raises(NotImplementedError):
            tb.UNet3d(**{**dict(in_channels=1, out_channels=1, depth=1), **kw})


def test_no_cpu_fallback():
    """The product path must fail loudly on CPU tensors instead of silently computing elsewhere."""
    net = tb.UNet3d(1, 1, depth=1, initial_features=2)
    with pytest.raises(RuntimeError, match="CUDA"):
        net(torch.zeros(1, 1, 8, 8, 8))
    with pytest.raises(RuntimeError, match="CUDA"):
        tb.DiceLoss()(torch.rand(1, 1, 4, 4), torch.rand(1, 1, 4, 4))
    with pytest.raises(ValueError):                                           # test/loss/test_dice.py:40-49
        tb.DiceLoss()(torch.rand(1, 2, 4, 4), torch.rand(1, 3, 4, 4))
    with pytest.raises(RuntimeError, match="CUDA"):
        tb.AffinityTransform([[-1, 0, 0]])(torch.zeros(4, 4, 4, dtype=torch.int64))


def test_eval_no_grad_keeps_nothing(golden_dir):
    z, net = build(golden_dir, "unet3d_d2_f4_instnorm")
    net._backend_override = TorchEmuBackend()
    net.eval()
    with torch.no_grad():
        y = net(torch.from_numpy(z["x"]))
    assert not y.requires_grad
    np.testing.assert_allclose(y.numpy(), z["y"], rtol=1e-4, atol=1e-5)


def test_autograd_contract(golden_dir):
    """(1) no silent None gradient for an input that requires grad, (2) an accurate error on a second backward,
    (3) an in-place edit of the returned prediction is caught by autograd's version check (the head's backward reads it)."""
    z, net = build(golden_dir, "unet3d_d2_f4_instnorm")
    net._backend_override = TorchEmuBackend()
    x = torch.from_numpy(z["x"])
    with pytest.raises(NotImplementedError, match="gradient w.r.t. its input"):
        net(x.clone().requires_grad_(True))
    y = net(x)
    loss = y.sum()
    loss.backward(retain_graph=True)
    with pytest.raises(RuntimeError, match="second backward"):
        loss.backward()
    y = net(x)
    y.clamp_(0.1, 0.9)
    with pytest.raises(RuntimeError, match="modified by an inplace operation"):
        y.sum().backward()


def test_zero_arena_views_are_zero_disjoint_and_aligned():
    """engine._ZeroArena: the per-pass reduction scratch -- zero-initialised, 16-byte aligned, non-overlapping views; a request
    larger than what is left of (or than) a chunk opens a new buffer."""
    import torch
    from torch_em_b200.engine import _ZeroArena
    ar = _ZeroArena(torch.device("cpu"))
    views = [ar.zeros((4, 32, 2)), ar.zeros((3, 5, 2)), ar.zeros((1, 1, 2)), ar.zeros((2, _ZeroArena.CHUNK)), ar.zeros((4, 512, 2))]
    for i, v in enumerate(views):
        assert v.dtype == torch.float32 and v.is_contiguous() and float(v.abs().sum()) == 0.0
        assert v.data_ptr() % 16 == 0
        v.fill_(float(i + 1))
    for i, v in enumerate(views):                       # nobody wrote into somebody else's view
        assert bool((v == float(i + 1)).all())
