"""The rest of the reference's U-Net constructor surface (torch_em/model/unet.py) on the fused kernel schedule, checked
DIRECTLY against the reference modules (imported through tests/ref_harness.py from /root/reference or baseline/_ref; skipped
when neither exists): BatchNorm / InstanceNormTrackStats incl. running statistics and eval mode (unet.py:391-406), side outputs
(:211-227), model-internal post-processing (:15-95), Decoder._crop for shapes the shape check would refuse (:363-370) and UNet2d
(:481-563).  Host logic on CPU through the emulation backend; tests/test_gpu_model_surface.py repeats them on the kernels."""
import numpy as np
import pytest
import torch

import torch_em_b200 as tb
from tests import ref_harness
from tests.emu_backend import TorchEmuBackend

torch_em = ref_harness.import_torch_em()
pytestmark = pytest.mark.skipif(torch_em is None, reason="reference package not available (/root/reference or baseline/_ref)")


def make_pair(cls_name, device="cpu", backend=None, **kw):
    """Reference model and ours with identical weights (and non-trivial affine norm parameters)."""
    import torch_em.model.unet as R
    torch.manual_seed(0)
    ref = getattr(R, cls_name)(**kw)
    with torch.no_grad():
        for k, p in ref.named_parameters():
            if p.dim() == 1 and (".block.0." in k or ".block.3." in k):
                p.add_(0.1 * torch.randn_like(p))
    ours = getattr(tb, cls_name)(**kw)
    missing, unexpected = ours.load_state_dict(ref.state_dict(), strict=True)
    assert not missing and not unexpected
    assert list(ours.state_dict().keys()) == list(ref.state_dict().keys())
    for k, v in ref.state_dict().items():
        assert tuple(ours.state_dict()[k].shape) == tuple(v.shape), k
    if backend is not None:
        ours._backend_override = backend
    return ref.to(device), ours.to(device)


def compare_step(ref, ours, x, rtol=1e-4, atol=1e-4, grtol=2e-3, weights=None):
    """One forward + backward through both; outputs (tensor or list), every parameter gradient and every buffer must agree."""
    yr, yo = ref(x), ours(x)
    lr_, lo = (yr, yo) if isinstance(yr, (list, tuple)) else ([yr], [yo])
    assert isinstance(yr, (list, tuple)) == isinstance(yo, (list, tuple)) and len(lr_) == len(lo)
    for a, b in zip(lr_, lo):
        assert tuple(a.shape) == tuple(b.shape)
        np.testing.assert_allclose(b.detach().cpu().numpy(), a.detach().cpu().numpy(), rtol=rtol, atol=atol)
    g = torch.Generator().manual_seed(1)
    ws = [torch.rand(a.shape, generator=g).to(a.device) for a in lr_]
    if weights is not None:
        ws = [None if k is None else w * k for w, k in zip(ws, weights)]
    sum((a * w).sum() for a, w in zip(lr_, ws) if w is not None).backward()
    sum((b * w).sum() for b, w in zip(lo, ws) if w is not None).backward()
    # absolute floor: 1e-5 of the largest gradient entry of the whole model (a conv bias in front of a norm has an exactly
    # zero gradient: both sides then hold rounding noise only)
    gmax = max(float(p.grad.abs().max()) for p in ref.parameters() if p.grad is not None)
    for (k, pr), (_, po) in zip(ref.named_parameters(), ours.named_parameters()):
        gr = pr.grad
        if gr is None:
            assert po.grad is None or float(po.grad.abs().max()) == 0.0, k
            continue
        gr = gr.cpu().numpy()
        np.testing.assert_allclose(po.grad.cpu().numpy(), gr, rtol=grtol, atol=1e-5 * gmax + 1e-4 * np.abs(gr).max(), err_msg=k)
    for (k, br), (_, bo) in zip(ref.named_buffers(), ours.named_buffers()):
        np.testing.assert_allclose(bo.cpu().numpy(), br.cpu().numpy(), rtol=1e-4, atol=1e-6, err_msg=k)


KW3 = dict(in_channels=1, out_channels=2, depth=2, initial_features=4, final_activation="Sigmoid")


@pytest.mark.parametrize("norm", ["BatchNorm", "InstanceNormTrackStats"])
def test_running_stat_norms_train_and_eval(norm):
    ref, ours = make_pair("UNet3d", backend=TorchEmuBackend(), norm=norm, **KW3)
    x = torch.randn(2, 1, 8, 16, 16)
    ref.train(); ours.train()
    compare_step(ref, ours, x)                               # batch / instance statistics, running statistics updated
    ref.zero_grad(); ours.zero_grad()
    compare_step(ref, ours, x * 1.5 + 0.3)
    ref.eval(); ours.eval()                                  # running statistics used, not updated
    ref.zero_grad(); ours.zero_grad()
    compare_step(ref, ours, torch.randn(1, 1, 8, 16, 16))


def test_invalid_norm_is_a_value_error():
    with pytest.raises(ValueError, match="Invalid norm"):
        tb.UNet3d(1, 1, depth=1, norm="LayerNorm")


def test_side_outputs():
    ref, ours = make_pair("UNet3d", backend=TorchEmuBackend(), return_side_outputs=True, **KW3)
    assert ours.out_channels == ref.out_channels == [2, 2]
    assert ours.init_kwargs["out_channels"] == ref.init_kwargs["out_channels"]
    x = torch.randn(1, 1, 8, 16, 16)
    compare_step(ref, ours, x)
    # a loss that uses only the LOW-resolution side output: the full-resolution head gets no gradient
    ref.zero_grad(); ours.zero_grad()
    compare_step(ref, ours, x, weights=[None, 1.0])
    ref2, ours2 = make_pair("AnisotropicUNet", backend=TorchEmuBackend(), in_channels=1, out_channels=[3, 1],
                            scale_factors=[[1, 2, 2], [2, 2, 2]], initial_features=4, return_side_outputs=True)
    compare_step(ref2, ours2, torch.randn(1, 1, 4, 16, 16))


@pytest.mark.parametrize("post", ["affinities_to_boundaries3d", "affinities_with_foreground_to_boundaries3d",
                                  "affinities_to_boundaries_anisotropic", "affinities_to_boundaries2d",
                                  "affinities_with_foreground_to_boundaries2d"])
def test_postprocessing(post):
    kw = dict(KW3, out_channels=4, postprocessing=post)
    ref, ours = make_pair("UNet3d", backend=TorchEmuBackend(), **kw)
    compare_step(ref, ours, torch.randn(1, 1, 8, 8, 8))
    with pytest.raises(ValueError, match="Invalid postprocessing"):
        tb.UNet3d(1, 1, depth=1, postprocessing="no_such_postprocessing")


def test_decoder_crop_for_unchecked_shapes():
    """check_shape=False with a spatial shape that is not divisible by the pooling factors: the skip connections are centre-
    cropped to the up-sampled shape (Decoder._crop, unet.py:363-373) and the output is smaller than the input.  The reference's
    crop only produces matching shapes for EVEN size differences (slice(sd, size - sd) with sd = diff // 2), which 2x pooling can
    never leave: the case exists for larger pooling factors.  Odd differences fail in both implementations."""
    kw = dict(in_channels=1, out_channels=2, scale_factors=[[3, 3, 3], [2, 2, 2]], initial_features=4, final_activation="Sigmoid",
              check_shape=False)
    ref, ours = make_pair("AnisotropicUNet", backend=TorchEmuBackend(), **kw)
    x = torch.randn(2, 1, 14, 12, 20)                       # 14 -> 4 -> 2 -> 4 -> 12 (crop 14 to 12); 20 -> 6 -> 3 -> 6 -> 18
    assert tuple(ref(x).shape) == (2, 2, 12, 12, 18)
    compare_step(ref, ours, x)
    with pytest.raises(ValueError, match="Invalid shape for U-Net"):
        ours._check_shape(x)
    bad = torch.randn(1, 1, 13, 12, 12)                     # 13 -> 4 -> 2 -> 4 -> 12: odd difference
    with pytest.raises(RuntimeError, match="Sizes of tensors must match"):
        ref(bad)
    with pytest.raises(RuntimeError, match="Sizes of tensors must match"):
        ours(bad)


def test_unet2d():
    kw = dict(in_channels=2, out_channels=3, depth=2, initial_features=4, final_activation="Sigmoid")
    ref, ours = make_pair("UNet2d", backend=TorchEmuBackend(), **kw)
    assert set(ours.init_kwargs) == set(ref.init_kwargs)
    assert ours.in_channels == 2 and ours.out_channels == 3 and ours.depth == 2
    x = torch.randn(2, 2, 16, 24)
    compare_step(ref, ours, x)
    with pytest.raises(ValueError, match="Invalid shape for U-Net"):
        ours(torch.zeros(1, 2, 18, 24))
    refg, oursg = make_pair("UNet2d", backend=TorchEmuBackend(), norm="GroupNorm", return_side_outputs=True, **kw)
    compare_step(refg, oursg, x)
    refn, oursn = make_pair("UNet2d", backend=TorchEmuBackend(), norm=None, check_shape=False, **dict(kw, final_activation=None))
    compare_step(refn, oursn, torch.randn(1, 2, 20, 24))
