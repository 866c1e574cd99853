"""Generate the golden vectors under tests/golden/ from the REFERENCE itself (run in the build container only).

    python tests/golden/make_golden.py

Imports the reference's torch-only modules by file path from /root/reference (``import torch_em`` fails in
this image: imageio/skimage/bioimage_cpp are absent), runs them on seeded inputs in fp32 on CPU and stores
inputs, weights, outputs, losses and gradients as small .npz fixtures.  /root/reference does not exist on the
GPU box, so the tests only ever read the .npz files.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

REF = "/root/reference/torch_em"
HERE = os.path.dirname(os.path.abspath(__file__))


def _load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def _np(sd):
    return {k: v.detach().cpu().numpy() for k, v in sd.items()}


def unet_case(ref_unet, ref_dice, name, ctor, kwargs, shape, seed):
    torch.manual_seed(seed)
    net = getattr(ref_unet, ctor)(**kwargs)
    if kwargs.get("norm") == "GroupNorm":  # make the affine params non-trivial
        with torch.no_grad():
            for k, p in net.named_parameters():
                if p.dim() == 1 and (".block.0." in k or ".block.3." in k):
                    p.add_(0.1 * torch.randn_like(p))
    x = torch.randn(*shape)
    t = (torch.rand(shape[0], kwargs["out_channels"], *shape[2:]) > 0.5).float()
    y = net(x)
    loss = ref_dice.DiceLoss()(y, t)
    loss.backward()
    out = {"x": x.numpy(), "t": t.numpy(), "y": y.detach().numpy(), "loss": loss.detach().numpy()}
    for k, v in net.state_dict().items():
        out["w:" + k] = v.numpy()
    for k, p in net.named_parameters():
        out["g:" + k] = p.grad.numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "loss", float(loss), "bytes", os.path.getsize(os.path.join(HERE, name + ".npz")))


def losses_case():
    """Dice-family losses beyond DiceLoss (loss/dice.py:136-256, combined_loss.py, distance_based.py): values and gradients
    from the reference classes (package imported through the stub finder of tests/ref_harness.py)."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from tests import ref_harness
    te = ref_harness.import_torch_em()
    L = te.loss
    from torch_em.loss.dice import BCEDiceLoss, BCEDiceLossWithLogits, DiceLossWithLogits
    from torch_em.loss.distance_based import DiceBasedDistanceLoss, DistanceLoss
    torch.manual_seed(11)
    shape = (2, 3, 4, 8, 8)
    x = torch.randn(*shape) * 2.0                       # logits
    p = torch.rand(*shape) * 0.98 + 0.01                # probabilities
    t = (torch.rand(*shape) > 0.5).float()
    out = {"x": x.numpy(), "p": p.numpy(), "t": t.numpy()}

    def record(name, loss_fn, inp, tgt):
        inp = inp.clone().requires_grad_(True)
        l = loss_fn(inp, tgt)
        if l.dim() > 0:
            l.sum().backward()
        else:
            l.backward()
        out["loss_" + name] = l.detach().numpy()
        out["grad_" + name] = inp.grad.numpy().copy()

    for red in ("sum", "mean", "max", "min", None):
        record(f"dice_logits_{red}", DiceLossWithLogits(reduce_channel=red), x, t)
    record("dice_logits_pooled", DiceLossWithLogits(channelwise=False), x, t)
    record("bce_dice", BCEDiceLoss(alpha=0.7, beta=1.3), p, t)
    record("bce_dice_pooled", BCEDiceLoss(alpha=1.0, beta=0.5, channelwise=False), p, t)
    record("bce_dice_logits", BCEDiceLossWithLogits(alpha=0.7, beta=1.3), x, t)
    record("combined", L.CombinedLoss(L.DiceLoss(), BCEDiceLoss(), loss_weights=[0.25, 0.75]), p, t)
    # distance losses: channel 0 = foreground, channels 1, 2 = distances in [0, 1]
    td = torch.cat([t[:, :1], torch.rand(2, 2, 4, 8, 8)], 1)
    out["td"] = td.numpy()
    for m in (True, False):
        record(f"distance_{m}", DistanceLoss(mask_distances_in_bg=m), p, td)
        record(f"dice_distance_{m}", DiceBasedDistanceLoss(mask_distances_in_bg=m), p, td)
    np.savez_compressed(os.path.join(HERE, "losses.npz"), **out)
    print("losses", {k: np.round(v, 5).tolist() for k, v in out.items() if k.startswith("loss")})


def labels2_case():
    """Label-target family (transform/label.py:133-244, 330-353; loss/affinity_side_loss.py:70-89) from the reference's own
    classes, imported through the stub finder.  OneHotTransform and segmentation_to_affinities are pure numpy / torch.  The two
    masked boundary transforms call skimage.segmentation.find_boundaries (absent here): it is substituted by the restatement of
    oracle/labels.py, so the class logic (which boundaries are masked, channel order, dtype handling) is the reference's while
    find_boundaries itself stays parity-unpinned."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from tests import ref_harness
    ref_harness.import_torch_em()
    import skimage.segmentation
    from oracle import labels as L
    import torch_em.transform.label as RL
    from torch_em.loss.affinity_side_loss import segmentation_to_affinities
    skimage.segmentation.find_boundaries = lambda img, mode="thick": L._find_boundaries_thick(img)
    rng = np.random.default_rng(21)
    seg = L.synthetic_labels((6, 12, 14), n_seeds=12, zero_fraction=0.2, seed=5)
    segi = seg.copy()
    segi[rng.random(seg.shape) < 0.1] = -1                      # ignore / mask label
    seg2 = L.synthetic_labels((16, 18), n_seeds=8, zero_fraction=0.2, seed=6)
    out = {"seg": seg, "segi": segi, "seg2": seg2}
    for b in (False, True):
        out[f"ntb_{b}"] = RL.NoToBackgroundBoundaryTransform(add_binary_target=b)(segi).astype("float32")
        out[f"ntb_bg2_{b}"] = RL.NoToBackgroundBoundaryTransform(bg_label=2, mask_label=-1, add_binary_target=b)(segi).astype("float32")
        out[f"bwi_{b}"] = RL.BoundaryTransformWithIgnoreLabel(add_binary_target=b)(segi).astype("float32")
        out[f"bwi2d_{b}"] = RL.BoundaryTransformWithIgnoreLabel(ignore_label=0, add_binary_target=b)(seg2).astype("float32")
    sem = rng.integers(0, 5, size=(5, 9, 11)).astype("int64")
    out["sem"] = sem
    out["onehot_none"] = RL.OneHotTransform()(sem)
    out["onehot_4"] = RL.OneHotTransform(class_ids=4)(sem)
    out["onehot_list"] = RL.OneHotTransform(class_ids=[3, 1, 7])(sem)
    offs3 = [[-1, 0, 0], [0, -1, 0], [0, 0, -1], [0, -3, 0], [2, 0, 0], [1, 2, -3], [0, 0, 20]]
    offs2 = [[-1, 0], [0, -1], [3, -2]]
    segb = np.stack([seg, np.roll(seg, 3, 1)])[:, None]
    out["segb"], out["offs3"], out["offs2"] = segb, np.array(offs3), np.array(offs2)
    out["segaffs3"] = segmentation_to_affinities(torch.from_numpy(segb), offs3).numpy()
    out["segaffs2"] = segmentation_to_affinities(torch.from_numpy(seg2[None, None]), offs2).numpy()
    out["segaffs3_float"] = segmentation_to_affinities(torch.from_numpy(segb).float(), offs3).numpy()
    np.savez_compressed(os.path.join(HERE, "labels2.npz"), **out)
    print("labels2 ok", {k: v.shape for k, v in out.items()})


def main():
    only = sys.argv[1:]                      # optional: names of the U-Net cases to (re)generate
    global unet_case
    if only:
        _all = unet_case
        unet_case = lambda a, b, name, *r: _all(a, b, name, *r) if name in only else None
    torch.set_num_threads(4)
    ref_unet = _load("ref_unet", "model/unet.py")
    ref_dice = _load("ref_dice", "loss/dice.py")
    ref_wrap = _load("ref_wrap", "loss/wrapper.py")

    unet_case(ref_unet, ref_dice, "unet3d_d2_f4_instnorm", "UNet3d",
              dict(in_channels=1, out_channels=2, depth=2, initial_features=4, final_activation="Sigmoid"),
              (2, 1, 16, 16, 16), 0)
    unet_case(ref_unet, ref_dice, "unet3d_d2_f8_groupnorm", "UNet3d",
              dict(in_channels=2, out_channels=3, depth=2, initial_features=8, final_activation="Sigmoid",
                   norm="GroupNorm"),
              (1, 2, 8, 16, 16), 1)
    # GroupNorm with MORE than one channel per group: GroupNorm(min(32, C), C) has 2 channels per group at 64 channels
    # (base block and the decoder block's first norm); batch 2 pins the per-sample statistics
    unet_case(ref_unet, ref_dice, "unet3d_d1_f32_groupnorm", "UNet3d",
              dict(in_channels=1, out_channels=2, depth=1, initial_features=32, final_activation="Sigmoid",
                   norm="GroupNorm"),
              (2, 1, 8, 16, 16), 6)
    unet_case(ref_unet, ref_dice, "unet3d_d1_f4_nonorm", "UNet3d",
              dict(in_channels=1, out_channels=1, depth=1, initial_features=4, final_activation=None, norm=None),
              (1, 1, 8, 8, 8), 2)
    unet_case(ref_unet, ref_dice, "aniso_f4_anisokernel", "AnisotropicUNet",
              dict(in_channels=1, out_channels=3, scale_factors=[[1, 2, 2], [2, 2, 2]], initial_features=4,
                   final_activation="Sigmoid", anisotropic_kernel=True),
              (1, 1, 4, 16, 16), 3)
    unet_case(ref_unet, ref_dice, "aniso_f4_isokernel", "AnisotropicUNet",
              dict(in_channels=1, out_channels=2, scale_factors=[[1, 2, 2], [2, 2, 2]], initial_features=4,
                   final_activation="Sigmoid", anisotropic_kernel=False),
              (1, 1, 4, 16, 16), 4)

    if only == ["losses"]:
        losses_case()
    if only == ["labels2"]:
        labels2_case()
    if only:
        return
    # Dice + masked Dice (LossWrapper(DiceLoss(), ApplyAndRemoveMask("multiply"))) with gradients
    torch.manual_seed(5)
    p = torch.rand(2, 3, 4, 8, 8, requires_grad=True)
    t = (torch.rand(2, 3, 4, 8, 8) > 0.5).float()
    m = (torch.rand(2, 3, 4, 8, 8) > 0.3).float()
    out = {"p": p.detach().numpy(), "t": t.numpy(), "m": m.numpy()}
    for red in ("sum", "mean", "max", "min"):
        p.grad = None
        l = ref_dice.DiceLoss(reduce_channel=red)(p, t)
        l.backward()
        out[f"loss_{red}"] = l.detach().numpy()
        out[f"grad_{red}"] = p.grad.numpy().copy()
    p.grad = None
    l = ref_dice.DiceLoss(channelwise=False)(p, t)
    l.backward()
    out["loss_pooled"], out["grad_pooled"] = l.detach().numpy(), p.grad.numpy().copy()
    p.grad = None
    wrapped = ref_wrap.LossWrapper(ref_dice.DiceLoss(), transform=ref_wrap.ApplyAndRemoveMask("multiply"))
    l = wrapped(p, torch.cat([t, m], 1))
    l.backward()
    out["loss_masked"], out["grad_masked"] = l.detach().numpy(), p.grad.numpy().copy()
    out["loss_ones_ones"] = ref_dice.DiceLoss()(torch.ones(1, 1, 8, 8), torch.ones(1, 1, 8, 8)).numpy()
    out["loss_ones_zeros"] = ref_dice.DiceLoss()(torch.ones(1, 1, 8, 8), torch.zeros(1, 1, 8, 8)).numpy()
    np.savez_compressed(os.path.join(HERE, "dice.npz"), **out)
    print("dice", {k: float(v) for k, v in out.items() if k.startswith("loss")})

    losses_case()
    labels2_case()

    # Affinity / boundary targets: the reference's arithmetic lives in absent third-party code; the fixture is
    # generated with the brute-force functions restated from the reference's own test (oracle/labels.py).
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import labels as L
    rng = np.random.default_rng(7)
    seg2 = rng.integers(1, 6, size=(64, 64)).astype("int64")
    seg2z = seg2.copy(); seg2z[rng.random(seg2.shape) < 0.25] = 0
    offs2 = [[-1, 0], [0, -1], [-3, 0], [0, -3], [4, 5], [-3, 2]]          # test_label_transforms.py:70-72
    seg3 = rng.integers(1, 5, size=(6, 10, 12)).astype("int64"); seg3[rng.random(seg3.shape) < 0.2] = 0
    offs3 = [[-1, 0, 0], [0, -1, 0], [0, 0, -1], [-2, 0, 0], [0, -3, 0], [0, 0, -3],
             [-3, 0, 0], [0, -9, 0], [0, 0, -9], [-4, 0, 0], [0, -27, 0], [0, 0, -27], [1, 2, -3]]
    out = {"seg2": seg2, "seg2z": seg2z, "offs2": np.array(offs2), "seg3": seg3, "offs3": np.array(offs3)}
    out["affs2"] = L.affs_brute_force(seg2, offs2)
    a, m_ = L.affs_brute_force_with_mask(seg2z, offs2, True); out["affs2z"], out["mask2z"] = a, m_
    a, m_ = L.affs_brute_force_with_mask(seg2z, offs2, False); out["affs2z_it"], out["mask2z_it"] = a, m_
    out["affs3"] = L.affs_brute_force(seg3, offs3)
    a, m_ = L.affs_brute_force_with_mask(seg3, offs3, True); out["affs3z"], out["mask3z"] = a, m_
    a, m_ = L.affs_brute_force_with_mask(seg3, offs3, False); out["affs3z_it"], out["mask3z_it"] = a, m_
    out["bound3"] = L.boundary_targets_scipy(seg3)
    np.savez_compressed(os.path.join(HERE, "labels.npz"), **out)
    print("labels ok")


if __name__ == "__main__":
    main()
