"""The oracle (oracle/) against the golden vectors generated from the reference (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import dice as odice
from oracle import labels as olabels
from oracle import unet as ounet

UNET_CASES = {
    "unet3d_d2_f4_instnorm": dict(scale_factors=[2, 2], norm="InstanceNorm", final_activation="Sigmoid"),
    "unet3d_d2_f8_groupnorm": dict(scale_factors=[2, 2], norm="GroupNorm", final_activation="Sigmoid"),
    "unet3d_d1_f32_groupnorm": dict(scale_factors=[2], norm="GroupNorm", final_activation="Sigmoid"),
    "unet3d_d1_f4_nonorm": dict(scale_factors=[2], norm=None, final_activation=None),
    "aniso_f4_anisokernel": dict(scale_factors=[[1, 2, 2], [2, 2, 2]], norm="InstanceNorm",
                                 final_activation="Sigmoid", anisotropic_kernel=True),
    "aniso_f4_isokernel": dict(scale_factors=[[1, 2, 2], [2, 2, 2]], norm="InstanceNorm",
                               final_activation="Sigmoid", anisotropic_kernel=False),
}


def load_case(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    sd = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w:")}
    grads = {k[2:]: z[k] for k in z.files if k.startswith("g:")}
    return z, sd, grads


@pytest.mark.parametrize("name", sorted(UNET_CASES))
def test_unet_oracle_matches_reference(golden_dir, name):
    z, sd, grads = load_case(golden_dir, name)
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    y = ounet.unet3d_forward(torch.from_numpy(z["x"]), sd, **UNET_CASES[name])
    # fp32 CPU vs fp32 CPU, same ATen kernels: tolerance is thread-count reduction-order noise only
    np.testing.assert_allclose(y.detach().numpy(), z["y"], rtol=1e-5, atol=1e-6)
    loss = odice.dice_loss(y, torch.from_numpy(z["t"]))
    np.testing.assert_allclose(loss.item(), z["loss"], rtol=1e-5)
    loss.backward()
    for k, g in grads.items():
        np.testing.assert_allclose(sd[k].grad.numpy(), g, rtol=1e-3, atol=1e-6, err_msg=k)


def test_unet_oracle_shape_check():
    sd = ounet.init_state_dict(1, 1, [2, 2, 2], initial_features=4)
    with pytest.raises(ValueError, match="Invalid shape for U-Net"):
        ounet.unet3d_forward(torch.zeros(1, 1, 12, 16, 16), sd, [2, 2, 2])


def test_flops_match_survey():
    # cfg2 3805.3 / 11401.5 GFLOP, cfg1 28.1 / 84.1 GFLOP
    f, _ = ounet.conv_flops_fwd(1, 2, [2] * 4, (128,) * 3, 4, 32)
    assert abs(f / 1e9 - 3805.3) < 0.1
    assert abs(ounet.conv_flops_train(1, 2, [2] * 4, (128,) * 3, 4, 32) / 1e9 - 11401.5) < 0.1
    assert abs(ounet.conv_flops_train(1, 2, [2] * 3, (64,) * 3, 1, 16) / 1e9 - 84.1) < 0.1


def test_dice_oracle_matches_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "dice.npz"))
    t, m = torch.from_numpy(z["t"]), torch.from_numpy(z["m"])
    for red in ("sum", "mean", "max", "min"):
        p = torch.from_numpy(z["p"]).requires_grad_(True)
        l = odice.dice_loss(p, t, reduce_channel=red)
        l.backward()
        np.testing.assert_allclose(l.item(), z[f"loss_{red}"], rtol=1e-6)
        np.testing.assert_allclose(p.grad.numpy(), z[f"grad_{red}"], rtol=1e-5, atol=1e-9)
    p = torch.from_numpy(z["p"]).requires_grad_(True)
    l = odice.dice_loss(p, t, channelwise=False)
    l.backward()
    np.testing.assert_allclose(l.item(), z["loss_pooled"], rtol=1e-6)
    np.testing.assert_allclose(p.grad.numpy(), z["grad_pooled"], rtol=1e-5, atol=1e-9)
    p = torch.from_numpy(z["p"]).requires_grad_(True)
    l = odice.masked_dice_loss(p, torch.cat([t, m], 1))
    l.backward()
    np.testing.assert_allclose(l.item(), z["loss_masked"], rtol=1e-6)
    np.testing.assert_allclose(p.grad.numpy(), z["grad_masked"], rtol=1e-5, atol=1e-9)
    # closed form used by the CUDA backward
    np.testing.assert_allclose(odice.dice_grad(p.detach(), t, m).numpy(), z["grad_masked"], rtol=1e-4, atol=1e-9)
    np.testing.assert_allclose(odice.dice_grad(p.detach(), t).numpy(), z["grad_sum"], rtol=1e-4, atol=1e-9)
    # known answers, test/loss/test_dice.py:25-38
    assert z["loss_ones_ones"] == 0.0 and z["loss_ones_zeros"] == 1.0
    assert odice.dice_loss(torch.ones(1, 1, 8, 8), torch.ones(1, 1, 8, 8)).item() == 0.0
    assert odice.dice_loss(torch.ones(1, 1, 8, 8), torch.zeros(1, 1, 8, 8)).item() == 1.0


def test_dice_shape_mismatch():
    with pytest.raises(ValueError):
        odice.dice_loss(torch.rand(1, 2, 4, 4), torch.rand(1, 3, 4, 4))


def test_labels_oracle_matches_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "labels.npz"))
    o2, o3 = z["offs2"].tolist(), z["offs3"].tolist()
    assert np.array_equal(olabels.affinity_targets(z["seg2"], o2), z["affs2"])
    a = olabels.affinity_targets(z["seg2z"], o2, ignore_label=0, add_mask=True)
    assert np.array_equal(a[:6], z["affs2z"]) and np.array_equal(a[6:], z["mask2z"])
    a = olabels.affinity_targets(z["seg2z"], o2, ignore_label=0, add_mask=True, include_ignore_transitions=True)
    assert np.array_equal(a[:6], z["affs2z_it"]) and np.array_equal(a[6:], z["mask2z_it"])
    n = len(o3)
    assert np.array_equal(olabels.affinity_targets(z["seg3"], o3), z["affs3"])
    a = olabels.affinity_targets(z["seg3"], o3, ignore_label=0, add_mask=True)
    assert np.array_equal(a[:n], z["affs3z"]) and np.array_equal(a[n:], z["mask3z"])
    a = olabels.affinity_targets(z["seg3"], o3, ignore_label=0, add_mask=True, include_ignore_transitions=True)
    assert np.array_equal(a[:n], z["affs3z_it"]) and np.array_equal(a[n:], z["mask3z_it"])
    assert np.array_equal(olabels.boundary_targets(z["seg3"]), z["bound3"])
    b = olabels.boundary_targets(z["seg3"], add_binary_target=True)
    assert b.shape[0] == 2 and np.array_equal(b[0], (z["seg3"] != 0).astype("float32"))


def test_labels_channel_layout():
    seg = olabels.synthetic_labels((4, 8, 8), n_seeds=6, seed=1)
    offs = [[-1, 0, 0], [0, -1, 0], [0, 0, -1]]
    a = olabels.affinity_targets(seg, offs, ignore_label=0, add_binary_target=True, add_mask=True)
    assert a.shape == (8, 4, 8, 8) and a.dtype == np.float32        # [fg, 3 affs, fg-mask, 3 masks]
    assert np.array_equal(a[0], (seg != 0).astype("float32"))
    assert np.array_equal(a[4], (seg != 0).astype("float32"))


def test_label_family_oracle_matches_reference_classes(golden_dir):
    """oracle/labels.py restatements vs the goldens produced by the reference's own classes (make_golden.py::labels2_case;
    find_boundaries itself substituted, see there)."""
    z = np.load(os.path.join(golden_dir, "labels2.npz"))
    segi, seg2, sem = z["segi"], z["seg2"], z["sem"]
    for b in (False, True):
        np.testing.assert_array_equal(olabels.no_to_background_boundary_targets(segi, add_binary_target=b), z[f"ntb_{b}"])
        np.testing.assert_array_equal(olabels.no_to_background_boundary_targets(segi, bg_label=2, add_binary_target=b), z[f"ntb_bg2_{b}"])
        np.testing.assert_array_equal(olabels.boundary_targets_with_ignore_label(segi, add_binary_target=b), z[f"bwi_{b}"])
        np.testing.assert_array_equal(olabels.boundary_targets_with_ignore_label(seg2, ignore_label=0, add_binary_target=b), z[f"bwi2d_{b}"])
    np.testing.assert_array_equal(olabels.one_hot_targets(sem), z["onehot_none"])
    np.testing.assert_array_equal(olabels.one_hot_targets(sem, 4), z["onehot_4"])
    np.testing.assert_array_equal(olabels.one_hot_targets(sem, [3, 1, 7]), z["onehot_list"])
    np.testing.assert_array_equal(olabels.segmentation_to_affinities(z["segb"], z["offs3"].tolist()), z["segaffs3"])
    np.testing.assert_array_equal(olabels.segmentation_to_affinities(seg2[None, None], z["offs2"].tolist()), z["segaffs2"])
