"""GPU label-target family (csrc/labels.cu) -- bit-exact against goldens produced by the reference's own classes
(tests/golden/make_golden.py::labels2_case: transform/label.py:133-244, 330-353, loss/affinity_side_loss.py:70-89) and against
the oracle on larger random volumes."""
import os

import numpy as np
import pytest
import torch

import torch_em_b200 as tb
from oracle import labels as olabels

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def z(golden_dir):
    return np.load(os.path.join(golden_dir, "labels2.npz"))


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.mark.parametrize("b", [False, True])
def test_masked_boundary_transforms_match_reference(z, b):
    segi, seg2 = _t(z["segi"]), _t(z["seg2"])
    np.testing.assert_array_equal(tb.NoToBackgroundBoundaryTransform(add_binary_target=b)(segi).cpu().numpy(), z[f"ntb_{b}"])
    np.testing.assert_array_equal(tb.NoToBackgroundBoundaryTransform(bg_label=2, add_binary_target=b)(segi).cpu().numpy(), z[f"ntb_bg2_{b}"])
    np.testing.assert_array_equal(tb.BoundaryTransformWithIgnoreLabel(add_binary_target=b)(segi).cpu().numpy(), z[f"bwi_{b}"])
    np.testing.assert_array_equal(tb.BoundaryTransformWithIgnoreLabel(ignore_label=0, add_binary_target=b, ndim=2)(seg2).cpu().numpy(),
                                  z[f"bwi2d_{b}"])
    # batched (N, 1, D, H, W) input -> (N, channels, D, H, W)
    out = tb.BoundaryTransformWithIgnoreLabel(add_binary_target=b)(torch.stack([segi, segi])[:, None])
    assert tuple(out.shape) == (2, 2 if b else 1) + tuple(segi.shape)
    np.testing.assert_array_equal(out[1].cpu().numpy(), z[f"bwi_{b}"])


def test_one_hot_matches_reference(z):
    sem = _t(z["sem"])
    np.testing.assert_array_equal(tb.OneHotTransform()(sem).cpu().numpy(), z["onehot_none"])
    np.testing.assert_array_equal(tb.OneHotTransform(class_ids=4)(sem).cpu().numpy(), z["onehot_4"])
    np.testing.assert_array_equal(tb.OneHotTransform(class_ids=[3, 1, 7])(sem).cpu().numpy(), z["onehot_list"])
    assert tb.OneHotTransform(4)(sem.to(torch.int32)).dtype == torch.float32


def test_segmentation_to_affinities_matches_reference(z):
    np.testing.assert_array_equal(tb.segmentation_to_affinities(_t(z["segb"]), z["offs3"].tolist()).cpu().numpy(), z["segaffs3"])
    np.testing.assert_array_equal(tb.segmentation_to_affinities(_t(z["seg2"][None, None]), z["offs2"].tolist()).cpu().numpy(), z["segaffs2"])
    np.testing.assert_array_equal(tb.segmentation_to_affinities(_t(z["segb"]).float(), z["offs3"].tolist()).cpu().numpy(), z["segaffs3_float"])
    with pytest.raises(AssertionError):
        tb.segmentation_to_affinities(_t(z["segb"]).repeat(1, 2, 1, 1, 1), z["offs3"].tolist())


def test_label_family_vs_oracle_on_larger_volumes():
    lab = olabels.synthetic_labels((24, 40, 36), n_seeds=60, zero_fraction=0.15, seed=9)
    rng = np.random.default_rng(3)
    lab[rng.random(lab.shape) < 0.05] = -1
    t = _t(lab)
    for b in (False, True):
        np.testing.assert_array_equal(tb.NoToBackgroundBoundaryTransform(add_binary_target=b)(t).cpu().numpy(),
                                      olabels.no_to_background_boundary_targets(lab, add_binary_target=b))
        np.testing.assert_array_equal(tb.BoundaryTransformWithIgnoreLabel(add_binary_target=b)(t).cpu().numpy(),
                                      olabels.boundary_targets_with_ignore_label(lab, add_binary_target=b))
    offs = [[-1, 0, 0], [0, -9, 0], [0, 0, 27], [4, -3, 2]]
    np.testing.assert_array_equal(tb.segmentation_to_affinities(t[None, None], offs).cpu().numpy(),
                                  olabels.segmentation_to_affinities(lab[None, None], offs))
    with pytest.raises(RuntimeError, match="CUDA"):
        tb.OneHotTransform(3)(torch.zeros(4, 4, dtype=torch.int64))
