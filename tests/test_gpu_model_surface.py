"""tests/test_model_surface_cpu.py repeated on the CUDA kernels (fp32: the exact CUDA-core path; the reference modules run on
the same GPU with TF32 off), plus UNet2d under bf16 autocast through the tensor-core kernels."""
import pytest
import torch

import torch_em_b200 as tb
from tests import ref_harness
from tests.test_model_surface_cpu import KW3, compare_step, make_pair

torch_em = ref_harness.import_torch_em()
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(torch_em is None, reason="reference package not available")]
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _no_tf32():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("norm", ["BatchNorm", "InstanceNormTrackStats"])
def test_running_stat_norms_train_and_eval(norm):
    ref, ours = make_pair("UNet3d", device=DEV, norm=norm, **KW3)
    x = torch.randn(2, 1, 8, 16, 16, device=DEV)
    ref.train(); ours.train()
    compare_step(ref, ours, x)
    ref.zero_grad(); ours.zero_grad()
    compare_step(ref, ours, x * 1.5 + 0.3)
    ref.eval(); ours.eval()
    ref.zero_grad(); ours.zero_grad()
    compare_step(ref, ours, torch.randn(1, 1, 8, 16, 16, device=DEV))


def test_side_outputs_and_postprocessing():
    ref, ours = make_pair("UNet3d", device=DEV, return_side_outputs=True, **KW3)
    x = torch.randn(1, 1, 8, 16, 16, device=DEV)
    compare_step(ref, ours, x)
    ref.zero_grad(); ours.zero_grad()
    compare_step(ref, ours, x, weights=[None, 1.0])
    refp, oursp = make_pair("UNet3d", device=DEV, **dict(KW3, out_channels=4, postprocessing="affinities_with_foreground_to_boundaries3d"))
    compare_step(refp, oursp, torch.randn(1, 1, 8, 8, 8, device=DEV))


def test_decoder_crop_for_unchecked_shapes():
    kw = dict(in_channels=1, out_channels=2, scale_factors=[[3, 3, 3], [2, 2, 2]], initial_features=4, final_activation="Sigmoid",
              check_shape=False)
    ref, ours = make_pair("AnisotropicUNet", device=DEV, **kw)
    x = torch.randn(2, 1, 14, 12, 20, device=DEV)
    assert tuple(ours(x).shape) == (2, 2, 12, 12, 18)
    ours.zero_grad()
    compare_step(ref, ours, x)
    with pytest.raises(RuntimeError, match="Sizes of tensors must match"):
        ours(torch.randn(1, 1, 13, 12, 12, device=DEV))


def test_unet2d_fp32_and_bf16():
    kw = dict(in_channels=2, out_channels=3, depth=2, initial_features=4, final_activation="Sigmoid")
    ref, ours = make_pair("UNet2d", device=DEV, **kw)
    x = torch.randn(2, 2, 16, 24, device=DEV)
    compare_step(ref, ours, x)
    refg, oursg = make_pair("UNet2d", device=DEV, norm="GroupNorm", return_side_outputs=True, **kw)
    compare_step(refg, oursg, x)
    # real widths under bf16 autocast: (1,3,3) kernels through the tcgen05 kernels; yardstick = the fp32 reference
    kw32 = dict(in_channels=1, out_channels=2, depth=3, initial_features=32, final_activation="Sigmoid")
    ref32, ours32 = make_pair("UNet2d", device=DEV, **kw32)
    xb = torch.randn(2, 1, 128, 128, device=DEV)
    from torch_em_b200.backend import default_backend
    default_backend().calls.clear()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = ours32(xb)
    y.sum().backward()
    calls = dict(default_backend().calls)
    assert any(k.startswith("plain:") for k in calls) and any(k.endswith(":wgrad") and not k.startswith("direct") for k in calls), calls
    with torch.no_grad():
        y_ref = ref32(xb)
    assert float((y.detach() - y_ref).norm() / y_ref.norm()) < 2e-2
