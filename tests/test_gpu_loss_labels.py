"""GPU parity of the fused Dice / masked-Dice / affinity / boundary kernels against the golden vectors generated from
the reference (tests/golden/dice.npz, labels.npz) and against the CPU oracle on larger seeded inputs."""
import os

import numpy as np
import pytest
import torch

import torch_em_b200 as tb
from oracle import dice as odice
from oracle import labels as olabels

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_dice_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "dice.npz"))
    t, m = torch.from_numpy(z["t"]).to(DEV), torch.from_numpy(z["m"]).to(DEV)
    for red in ("sum", "mean", "max", "min"):
        p = torch.from_numpy(z["p"]).to(DEV).requires_grad_(True)
        l = tb.DiceLoss(reduce_channel=red)(p, t)
        l.backward()
        np.testing.assert_allclose(l.item(), z[f"loss_{red}"], rtol=1e-5)
        np.testing.assert_allclose(p.grad.cpu().numpy(), z[f"grad_{red}"], rtol=1e-4, atol=1e-8)
    p = torch.from_numpy(z["p"]).to(DEV).requires_grad_(True)
    l = tb.DiceLoss(channelwise=False)(p, t)
    l.backward()
    np.testing.assert_allclose(l.item(), z["loss_pooled"], rtol=1e-5)
    np.testing.assert_allclose(p.grad.cpu().numpy(), z["grad_pooled"], rtol=1e-4, atol=1e-8)
    # the reference's affinity-loss idiom: LossWrapper(DiceLoss(), ApplyAndRemoveMask("multiply"))
    for loss in (tb.LossWrapper(tb.DiceLoss(), tb.ApplyAndRemoveMask("multiply")), tb.AffinityLoss()):
        p = torch.from_numpy(z["p"]).to(DEV).requires_grad_(True)
        l = loss(p, torch.cat([t, m], 1))
        l.backward()
        np.testing.assert_allclose(l.item(), z["loss_masked"], rtol=1e-5)
        np.testing.assert_allclose(p.grad.cpu().numpy(), z["grad_masked"], rtol=1e-4, atol=1e-8)
    # per-channel scores and the un-inverted score
    p = torch.from_numpy(z["p"]).to(DEV)
    ref = odice.dice_score(torch.from_numpy(z["p"]), torch.from_numpy(z["t"]), reduce_channel=None)
    np.testing.assert_allclose(tb.dice_score(p, t, reduce_channel=None).cpu().numpy(), ref.numpy(), rtol=1e-5)
    for red in ("sum", "mean", "max", "min"):
        ref = odice.dice_score(torch.from_numpy(z["p"]), torch.from_numpy(z["t"]), reduce_channel=red)
        np.testing.assert_allclose(tb.dice_score(p, t, reduce_channel=red).item(), ref.item(), rtol=1e-5)


def test_dice_known_answers_and_errors():
    # test/loss/test_dice.py:25-38
    ones, zeros = torch.ones(1, 1, 8, 8, device=DEV), torch.zeros(1, 1, 8, 8, device=DEV)
    assert tb.DiceLoss()(ones, ones).item() == pytest.approx(0.0, abs=1e-7)
    assert tb.DiceLoss()(ones, zeros).item() == pytest.approx(1.0, abs=1e-7)
    # Dice(0, 0): denominator below the clamp -> loss = C with zero gradient
    p = torch.zeros(1, 2, 4, 4, device=DEV, requires_grad=True)
    l = tb.DiceLoss()(p, torch.zeros(1, 2, 4, 4, device=DEV))
    l.backward()
    assert l.item() == pytest.approx(2.0) and float(p.grad.abs().max()) == 0.0
    with pytest.raises(ValueError):                                       # test_dice.py:40-49
        tb.DiceLoss()(torch.rand(1, 2, 4, 4, device=DEV), torch.rand(1, 3, 4, 4, device=DEV))
    for red, shape in ((None, (3,)), ("sum", ()), ("mean", ()), ("max", ()), ("min", ())):   # test_dice.py:51-65
        out = tb.DiceLoss(reduce_channel=red)(torch.rand(2, 3, 8, 8, device=DEV), torch.rand(2, 3, 8, 8, device=DEV))
        assert tuple(out.shape) == shape


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_dice_large_ragged(dtype):
    # odd sizes (no 16-byte alignment), 3-D, bf16 predictions like the autocast train step
    g = torch.Generator().manual_seed(0)
    p = torch.rand((2, 3, 5, 33, 31), generator=g).to(dtype)
    t = (torch.rand((2, 3, 5, 33, 31), generator=g) > 0.5).float()
    m = (torch.rand((2, 3, 5, 33, 31), generator=g) > 0.3).float()
    pr = p.float().clone().requires_grad_(True)
    ref = odice.masked_dice_loss(pr, torch.cat([t, m], 1))
    ref.backward()
    pg = p.to(DEV).requires_grad_(True)
    l = tb.LossWrapper(tb.DiceLoss(), tb.ApplyAndRemoveMask("multiply"))(pg, torch.cat([t, m], 1).to(DEV))
    l.backward()
    np.testing.assert_allclose(l.item(), ref.item(), rtol=1e-5)
    tol = dict(rtol=1e-4, atol=1e-9) if dtype == torch.float32 else dict(rtol=1e-2, atol=1e-7)
    np.testing.assert_allclose(pg.grad.float().cpu().numpy(), pr.grad.numpy(), **tol)


@pytest.mark.parametrize("method", ["multiply", "crop"])
def test_masking_gradient_property(method):
    # test/loss/test_loss_wrapper.py:6-34,64-87: 0 < loss < 1, gradient exactly 0 outside the mask, non-zero inside
    g = torch.Generator().manual_seed(1)
    shape = (1, 1, 64, 64)
    p = torch.rand(shape, generator=g).to(DEV).requires_grad_(True)
    t = (torch.rand(shape, generator=g) > 0.5).float()
    m = (torch.rand(shape, generator=g) > 0.5).float()
    loss = tb.LossWrapper(tb.DiceLoss(), tb.ApplyAndRemoveMask(masking_method=method))
    l = loss(p, torch.cat([t, m], 1).to(DEV))
    l.backward()
    assert 0.0 < l.item() < 1.0
    grad, mask = p.grad.cpu().numpy(), m.numpy().astype(bool)
    assert (grad[~mask] == 0).all() and (grad[mask] != 0).all()
    # MaskIgnoreLabel, wrapper.py:155-183
    p2 = p.detach().clone().requires_grad_(True)
    t2 = t.clone(); t2[m == 0] = -1
    l2 = tb.LossWrapper(tb.DiceLoss(), tb.MaskIgnoreLabel(-1, masking_method=method))(p2, t2.to(DEV))
    l2.backward()
    np.testing.assert_allclose(l2.item(), l.item(), rtol=1e-5)
    np.testing.assert_allclose(p2.grad.cpu().numpy(), grad, rtol=1e-4, atol=1e-9)
    with pytest.raises(ValueError):                                       # multi-channel mask with crop
        tb.LossWrapper(tb.DiceLoss(), tb.ApplyAndRemoveMask("crop"))(torch.rand(1, 2, 8, 8, device=DEV), torch.rand(1, 4, 8, 8, device=DEV))


def test_label_targets_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "labels.npz"))
    o2, o3 = z["offs2"].tolist(), z["offs3"].tolist()
    n3 = len(o3)
    s2, s2z, s3 = (torch.from_numpy(z[k]).to(DEV) for k in ("seg2", "seg2z", "seg3"))
    eq = lambda a, b: np.array_equal(a.cpu().numpy(), b)
    assert eq(tb.AffinityTransform(o2)(s2), z["affs2"])
    a = tb.AffinityTransform(o2, ignore_label=0, add_mask=True)(s2z)
    assert eq(a[:6], z["affs2z"]) and eq(a[6:], z["mask2z"])
    a = tb.AffinityTransform(o2, ignore_label=0, add_mask=True, include_ignore_transitions=True)(s2z)
    assert eq(a[:6], z["affs2z_it"]) and eq(a[6:], z["mask2z_it"])
    assert eq(tb.AffinityTransform(o3)(s3), z["affs3"])
    a = tb.AffinityTransform(o3, ignore_label=0, add_mask=True)(s3)
    assert eq(a[:n3], z["affs3z"]) and eq(a[n3:], z["mask3z"])
    a = tb.AffinityTransform(o3, ignore_label=0, add_mask=True, include_ignore_transitions=True)(s3)
    assert eq(a[:n3], z["affs3z_it"]) and eq(a[n3:], z["mask3z_it"])
    assert eq(tb.BoundaryTransform()(s3), z["bound3"])
    b = tb.BoundaryTransform(add_binary_target=True)(s3)
    assert eq(b, olabels.boundary_targets(z["seg3"], add_binary_target=True))


CREMI_OFFSETS = [[-1, 0, 0], [0, -1, 0], [0, 0, -1], [-2, 0, 0], [0, -3, 0], [0, 0, -3],
                 [-3, 0, 0], [0, -9, 0], [0, 0, -9], [-4, 0, 0], [0, -27, 0], [0, 0, -27]]


def test_label_targets_batched_vs_oracle():
    labs = np.stack([olabels.synthetic_labels((12, 40, 36), n_seeds=30, seed=s) for s in range(3)])
    lt = torch.from_numpy(labs).to(DEV)
    for kw in (dict(), dict(ignore_label=0, add_mask=True), dict(ignore_label=0, add_mask=True, add_binary_target=True),
               dict(add_binary_target=True, add_mask=True), dict(ignore_label=0, add_mask=True, include_ignore_transitions=True)):
        got = tb.AffinityTransform(CREMI_OFFSETS, **kw)(lt).cpu().numpy()
        for i in range(3):
            assert np.array_equal(got[i], olabels.affinity_targets(labs[i], CREMI_OFFSETS, **kw)), kw
    got = tb.BoundaryTransform(add_binary_target=True)(lt[:, None]).cpu().numpy()
    for i in range(3):
        assert np.array_equal(got[i], olabels.boundary_targets(labs[i], add_binary_target=True))
    # int32 labels and a degenerate volume (one label: no boundaries, only out-of-bounds disaffinities)
    one = torch.ones((4, 4, 4), dtype=torch.int32, device=DEV)
    assert float(tb.BoundaryTransform()(one).sum()) == 0.0
    a = tb.AffinityTransform([[-1, 0, 0]])(one)
    assert float(a[0, 0].sum()) == 16.0 and float(a[0, 1:].sum()) == 0.0


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("iit", [False, True])
def test_fused_affinity_loss_vs_oracle(dtype, iit):
    labs = np.stack([olabels.synthetic_labels((10, 37, 29), n_seeds=25, seed=10 + s) for s in range(2)])
    g = torch.Generator().manual_seed(3)
    p = torch.rand((2, len(CREMI_OFFSETS)) + labs.shape[1:], generator=g).to(dtype)
    tgt = np.stack([olabels.affinity_targets(l, CREMI_OFFSETS, ignore_label=0, add_mask=True, include_ignore_transitions=iit)
                    for l in labs])
    pr = p.float().clone().requires_grad_(True)
    ref = odice.masked_dice_loss(pr, torch.from_numpy(tgt))
    ref.backward()
    pg = p.to(DEV).requires_grad_(True)
    loss = tb.AffinityLoss(CREMI_OFFSETS, ignore_label=0, include_ignore_transitions=iit)
    l = loss(pg, torch.from_numpy(labs).to(DEV))
    l.backward()
    np.testing.assert_allclose(l.item(), ref.item(), rtol=1e-5)
    tol = dict(rtol=1e-4, atol=1e-9) if dtype == torch.float32 else dict(rtol=1e-2, atol=1e-7)
    np.testing.assert_allclose(pg.grad.float().cpu().numpy(), pr.grad.numpy(), **tol)
