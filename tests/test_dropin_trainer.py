"""The drop-in claim, executed: the REAL, unmodified ``torch_em.default_segmentation_trainer`` / ``DefaultTrainer``
(segmentation.py:466-577, default_trainer.py) trains our model with our loss, writes checkpoints, restores them through
``DefaultTrainer.from_checkpoint`` and exchanges ``state_dict``s with the reference ``UNet3d`` in both directions.
Template: the reference's own test/trainer/test_default_trainer.py:69-157.

The reference package is imported from /root/reference (build container) or baseline/_ref (GPU box) through the stub
finder of tests/ref_harness.py; the tests skip cleanly when neither exists.
  * CPU flavour: the model's kernel schedule runs on the plain-PyTorch emulation of the C ABI (tests/emu_backend.py);
    the loss is the reference's own DiceLoss (ours has no CPU path by design).
  * GPU flavour (-m gpu): our model AND our loss on cuda:0 under the trainer's bf16 autocast.
"""
import os

import numpy as np
import pytest
import torch

import torch_em_b200 as tb
from tests import ref_harness
from tests.emu_backend import TorchEmuBackend

torch_em = ref_harness.import_torch_em()
pytestmark = pytest.mark.skipif(torch_em is None, reason="reference package not available (/root/reference or baseline/_ref)")

MODEL_KW = dict(in_channels=1, out_channels=2, depth=2, initial_features=4, final_activation="Sigmoid")


class SyntheticPatches(torch.utils.data.Dataset):
    """In-memory (raw, binary target) patches; module-level so the trainer can pickle it into the checkpoint."""

    def __init__(self, n=4, shape=(16, 16, 16), seed=0):
        g = torch.Generator().manual_seed(seed)
        self.raw = torch.randn((n, 1) + tuple(shape), generator=g)
        # learnable target: thresholded smooth function of the raw data
        sm = torch.nn.functional.avg_pool3d(self.raw, 3, 1, 1)
        self.labels = torch.cat([(sm > 0).float(), (sm < 0).float()], 1)

    def __len__(self):
        return self.raw.shape[0]

    def __getitem__(self, i):
        return self.raw[i], self.labels[i]


LOSS_LOG = []


class RecordingDiceLoss(tb.DiceLoss):
    """Module-level (the trainer pickles the loss by dotted class path into the checkpoint)."""

    def forward(self, p, t):
        v = super().forward(p, t)
        LOSS_LOG.append(float(v.detach()))
        return v


if torch_em is not None:
    class RecordingRefDiceLoss(torch_em.loss.DiceLoss):
        def forward(self, p, t):
            v = super().forward(p, t)
            LOSS_LOG.append(float(v.detach()))
            return v


def _loaders():
    from torch_em.segmentation import get_data_loader
    ds = SyntheticPatches()
    return get_data_loader(ds, batch_size=2, shuffle=True, pin_memory=False), get_data_loader(ds, batch_size=2, pin_memory=False)


def _run_trainer(tmp_path, device, loss, metric, compile_model, **extra):
    train, val = _loaders()
    torch.manual_seed(0)
    model = tb.UNet3d(**MODEL_KW)
    trainer = torch_em.default_segmentation_trainer(
        name="dropin", model=model, train_loader=train, val_loader=val, loss=loss, metric=metric, device=device,
        logger=None, compile_model=compile_model, save_root=str(tmp_path), **extra)
    w0 = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    trainer.fit(iterations=10)
    assert trainer.iteration == 10
    folder = os.path.join(str(tmp_path), "checkpoints", "dropin")
    assert os.path.exists(os.path.join(folder, "best.pt")) and os.path.exists(os.path.join(folder, "latest.pt"))
    w1 = trainer.model.state_dict()
    assert any(not torch.equal(w0[k.replace("_orig_mod.", "")], w1[k].cpu()) for k in w1), "weights did not move"
    return trainer, folder


def _check_checkpoint_roundtrip(trainer, folder, device):
    from torch_em.trainer import DefaultTrainer
    ck = torch.load(os.path.join(folder, "latest.pt"), map_location="cpu", weights_only=False)
    assert ck["init"]["model_class"] == "torch_em_b200.model.unet.UNet3d"
    assert ck["init"]["model_kwargs"]["depth"] == 2
    t2 = DefaultTrainer.from_checkpoint(folder, name="latest", device=device)
    assert t2.iteration == trainer.iteration
    assert type(t2.model).__name__ == "UNet3d" and type(t2.model).__module__ == "torch_em_b200.model.unet"
    assert torch_em.util.model_is_equal(trainer.model, t2.model)
    lr1 = [pg["lr"] for pg in trainer.optimizer.param_groups][0]
    lr2 = [pg["lr"] for pg in t2.optimizer.param_groups][0]
    assert lr1 == lr2
    t2.fit(4)
    assert t2.iteration == 14
    return ck


def _check_state_dict_exchange(ck, device, backend=None):
    """Reference UNet3d <-> ours, both directions, and identical predictions from identical weights."""
    from torch_em.model import UNet3d as RefUNet3d
    ref = RefUNet3d(**MODEL_KW)
    missing, unexpected = ref.load_state_dict(ck["model_state"], strict=True)      # ours -> reference
    assert not missing and not unexpected
    ours = tb.UNet3d(**MODEL_KW)
    torch.manual_seed(3)
    ref2 = RefUNet3d(**MODEL_KW)
    missing, unexpected = ours.load_state_dict(ref2.state_dict(), strict=True)      # reference -> ours
    assert not missing and not unexpected
    if backend is not None:
        ours._backend_override = backend
    x = torch.randn(1, 1, 16, 16, 16)
    with torch.no_grad():
        y_ref = ref2(x)
        y = ours.to(device)(x.to(device)).cpu()
    np.testing.assert_allclose(y.numpy(), y_ref.numpy(), rtol=1e-4, atol=1e-4)


def test_real_trainer_cpu_emulated_backend(tmp_path, monkeypatch):
    import torch_em_b200.model.unet as our_unet
    emu = TorchEmuBackend()
    monkeypatch.setattr(our_unet, "default_backend", lambda: emu)
    loss = torch_em.loss.DiceLoss()
    trainer, folder = _run_trainer(tmp_path, "cpu", loss, torch_em.loss.DiceLoss(), compile_model=False, mixed_precision=False)
    ck = _check_checkpoint_roundtrip(trainer, folder, "cpu")
    _check_state_dict_exchange(ck, "cpu", backend=emu)


def test_real_trainer_default_compile_setting_cpu(tmp_path, monkeypatch):
    """``compile_model=None`` is the trainer's default and means torch.compile (default_trainer.py:541, util.py:38-74): the
    model's forward opts out of Dynamo tracing (it is already one fused kernel schedule), so the default works unchanged."""
    import torch_em_b200.model.unet as our_unet
    emu = TorchEmuBackend()
    monkeypatch.setattr(our_unet, "default_backend", lambda: emu)
    trainer, folder = _run_trainer(tmp_path, "cpu", torch_em.loss.DiceLoss(), torch_em.loss.DiceLoss(), compile_model=None,
                                   mixed_precision=False)
    assert torch_em.util.util.is_compiled(trainer.model)
    ck = torch.load(os.path.join(folder, "latest.pt"), map_location="cpu", weights_only=False)
    assert ck["init"]["model_class"] == "torch_em_b200.model.unet.UNet3d"


@pytest.mark.gpu
@pytest.mark.parametrize("compile_model", [False, None])
def test_real_trainer_gpu_bf16(tmp_path, compile_model):
    dev = "cuda:0"
    LOSS_LOG.clear()
    tb.reset_launch_count()
    trainer, folder = _run_trainer(tmp_path, dev, RecordingDiceLoss(), tb.DiceLoss(), compile_model=compile_model,
                                   mixed_precision=True, mixed_precision_dtype="bfloat16")
    loss_vals = list(LOSS_LOG)
    assert tb.launch_count() > 100, "the CUDA kernels of libb200em did not run"
    assert not trainer.scaler.is_enabled()
    assert loss_vals[-1] < loss_vals[0], loss_vals
    if compile_model is False:
        ck = _check_checkpoint_roundtrip(trainer, folder, dev)
        _check_state_dict_exchange(ck, dev)


@pytest.mark.gpu
def test_real_trainer_gpu_default_mixed_precision_fp16(tmp_path):
    """``mixed_precision=True`` with the trainer's DEFAULT dtype (float16 + GradScaler, default_trainer.py:132-142): the model
    serves fp16 autocast on the h16 path (fp32 activations, fp16 tensor-core operands), the scaler stays enabled and works."""
    dev = "cuda:0"
    LOSS_LOG.clear()
    tb.reset_launch_count()
    trainer, folder = _run_trainer(tmp_path, dev, RecordingDiceLoss(), tb.DiceLoss(), compile_model=False, mixed_precision=True)
    loss_vals = list(LOSS_LOG)
    assert tb.launch_count() > 100, "the CUDA kernels of libb200em did not run"
    assert trainer.scaler.is_enabled()
    assert all(v == v for v in loss_vals) and loss_vals[-1] < loss_vals[0], loss_vals


@pytest.mark.gpu
def test_real_trainer_gpu_fp32_matches_reference_model_trajectory(tmp_path):
    """Same seed, same data order, fp32: the reference UNet3d + DiceLoss and ours must produce the same loss curve
    through the same real trainer (rtol 2e-3 over 10 AdamW steps, the bound of tests/test_gpu_unet.py)."""
    from torch_em.model import UNet3d as RefUNet3d
    dev = "cuda:0"
    curves = []
    for flavour in ("reference", "ours"):
        torch.manual_seed(0)
        ours = tb.UNet3d(**MODEL_KW)
        if flavour == "reference":
            model = RefUNet3d(**MODEL_KW)
            model.load_state_dict(ours.state_dict())
            base_loss, rec_loss = torch_em.loss.DiceLoss, RecordingRefDiceLoss
        else:
            model, base_loss, rec_loss = ours, tb.DiceLoss, RecordingDiceLoss
        LOSS_LOG.clear()

        torch.manual_seed(1)
        ds = SyntheticPatches()
        from torch_em.segmentation import get_data_loader
        train = get_data_loader(ds, batch_size=2, shuffle=False, pin_memory=False)
        old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
        try:
            trainer = torch_em.default_segmentation_trainer(
                name=f"traj_{flavour}", model=model, train_loader=train, val_loader=train, loss=rec_loss(), metric=base_loss(),
                device=dev, logger=None, compile_model=False, mixed_precision=False, save_root=str(tmp_path))
            trainer.fit(iterations=10)
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
        curves.append(list(LOSS_LOG))
    np.testing.assert_allclose(curves[1], curves[0], rtol=2e-3)
