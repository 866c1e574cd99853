"""Oracle: instance labels -> affinity / boundary training targets (TEST INFRASTRUCTURE ONLY).

The reference delegates the arithmetic to un-vendored third-party code
(``bioimage_cpp.affinities.compute_affinities``, version unpinned -- setup.py:8; and
``skimage.segmentation.find_boundaries``, unpinned -- setup.py:17), neither of which is in this image.
What pins the results is the reference's own test: test/transform/test_label_transforms.py:5-55 holds
pure-python brute-force functions that ``AffinityTransform`` must equal.  ``affs_brute_force*`` below restate
those two functions generalised from 2-D to n-D (same rules, same float32 outputs); ``affinity_targets`` is
the vectorised numpy form of the whole ``AffinityTransform.__call__`` (transform/label.py:290-327) including
``add_binary_target``, ``add_mask`` and ``include_ignore_transitions`` (label.py:277-288).

``boundary_targets`` restates ``BoundaryTransform`` (label.py:100-129):  find_boundaries(mode="thick") is
``grey_dilation(lab, cross) != grey_erosion(lab, cross)`` with reflect borders, i.e. "some in-bounds
6-neighbour carries a different label".  No reference test pins it: PARITY UNPINNED; pinned against scipy.
"""
import itertools

import numpy as np


def affs_brute_force(seg, offsets):
    """n-D restatement of test_label_transforms.py:5-20 (no ignore label)."""
    shape = seg.shape
    affs = np.zeros((len(offsets),) + shape, dtype="float32")
    for p in itertools.product(*[range(s) for s in shape]):
        val = seg[p]
        for c, off in enumerate(offsets):
            q = tuple(pi + oi for pi, oi in zip(p, off))
            if any(qi < 0 or qi >= s for qi, s in zip(q, shape)):
                affs[(c,) + p] = 1.0
                continue
            affs[(c,) + p] = 0.0 if val == seg[q] else 1.0
    return affs


def affs_brute_force_with_mask(seg, offsets, mask_bg_transition=True, ignore_label=0):
    """n-D restatement of test_label_transforms.py:23-55."""
    shape = seg.shape
    affs = np.zeros((len(offsets),) + shape, dtype="float32")
    mask = np.zeros((len(offsets),) + shape, dtype="float32")
    for p in itertools.product(*[range(s) for s in shape]):
        val = seg[p]
        for c, off in enumerate(offsets):
            q = tuple(pi + oi for pi, oi in zip(p, off))
            idx = (c,) + p
            if any(qi < 0 or qi >= s for qi, s in zip(q, shape)):
                affs[idx], mask[idx] = 1.0, 0.0
                continue
            oval = seg[q]
            n_ignore = int(val == ignore_label) + int(oval == ignore_label)
            if n_ignore == 2 or (n_ignore == 1 and mask_bg_transition):
                affs[idx], mask[idx] = 1.0, 0.0
                continue
            affs[idx] = 0.0 if val == oval else 1.0
            mask[idx] = 1.0
    return affs, mask


def _shifted_views(shape, off):
    """Slices (src p, dst q=p+off) of the in-bounds region for one offset."""
    p_sl, q_sl = [], []
    for s, o in zip(shape, off):
        lo, hi = max(0, -o), min(s, s - o)
        if hi <= lo:
            return None, None
        p_sl.append(slice(lo, hi))
        q_sl.append(slice(lo + o, hi + o))
    return tuple(p_sl), tuple(q_sl)


def affinity_targets(labels, offsets, ignore_label=None, add_binary_target=False, add_mask=False,
                     include_ignore_transitions=False):
    """Vectorised AffinityTransform.__call__ (label.py:290-327): float32 (channels, *spatial)."""
    labels = np.asarray(labels)
    shape = labels.shape
    n = len(offsets)
    disaff = np.ones((n,) + shape, dtype="float32")     # OOB -> 1
    mask = np.zeros((n,) + shape, dtype="float32")       # OOB -> 0
    for c, off in enumerate(offsets):
        p, q = _shifted_views(shape, off)
        if p is None:
            continue
        a, b = labels[p], labels[q]
        d = (a != b).astype("float32")
        m = np.ones_like(d)
        if ignore_label is not None:
            ia, ib = a == ignore_label, b == ignore_label
            invalid = ia | ib
            d[invalid] = 1.0
            m[invalid] = 0.0
            if include_ignore_transitions:
                trans = ia ^ ib
                d[trans] = 1.0
                m[trans] = 1.0
        disaff[(c,) + p] = d
        mask[(c,) + p] = m
    out = disaff
    if add_binary_target:
        out = np.concatenate([(labels != 0)[None].astype("float32"), out], axis=0)
    if add_mask:
        if add_binary_target:
            mb = np.ones((1,) + shape, "float32") if ignore_label is None else \
                (labels != ignore_label)[None].astype("float32")
            mask = np.concatenate([mb, mask], axis=0)
        out = np.concatenate([out, mask], axis=0)
    return out


def boundary_targets(labels, add_binary_target=False):
    """BoundaryTransform(mode='thick') (label.py:113-129): float32 (1 or 2, *spatial), [foreground, boundary]."""
    labels = np.asarray(labels)
    b = np.zeros(labels.shape, dtype=bool)
    for ax in range(labels.ndim):
        sl_a = [slice(None)] * labels.ndim
        sl_b = [slice(None)] * labels.ndim
        sl_a[ax], sl_b[ax] = slice(0, -1), slice(1, None)
        diff = labels[tuple(sl_a)] != labels[tuple(sl_b)]
        b[tuple(sl_a)] |= diff
        b[tuple(sl_b)] |= diff
    out = b[None].astype("float32")
    if add_binary_target:
        out = np.concatenate([(labels != 0)[None].astype("float32"), out], axis=0)
    return out


def boundary_targets_scipy(labels):
    """The scipy pin: what skimage.find_boundaries(mode='thick') computes (morphology on a cross footprint)."""
    from scipy import ndimage as ndi
    fp = ndi.generate_binary_structure(labels.ndim, 1)
    lab = labels.astype("int64")
    return (ndi.grey_dilation(lab, footprint=fp) != ndi.grey_erosion(lab, footprint=fp)).astype("float32")[None]


def synthetic_labels(shape, n_seeds=40, zero_fraction=0.1, seed=0):
    """Deterministic Voronoi-like instance segmentation, int64."""
    rng = np.random.default_rng(seed)
    from scipy.spatial import cKDTree
    pts = rng.random((n_seeds, len(shape))) * np.array(shape)
    grid = np.stack(np.meshgrid(*[np.arange(s) for s in shape], indexing="ij"), -1).reshape(-1, len(shape))
    _, idx = cKDTree(pts).query(grid)
    lab = (idx + 1).reshape(shape).astype("int64")
    if zero_fraction > 0:
        ids = rng.choice(np.arange(1, n_seeds + 1), size=max(1, int(n_seeds * zero_fraction)), replace=False)
        lab[np.isin(lab, ids)] = 0
    return lab


def _find_boundaries_thick(img):
    """skimage.segmentation.find_boundaries(img, mode='thick') restated (parity unpinned, see the module docstring): bool."""
    return boundary_targets(np.asarray(img).astype("int64"))[0].astype(bool)


def no_to_background_boundary_targets(labels, bg_label=0, mask_label=-1, add_binary_target=False):
    """NoToBackgroundBoundaryTransform.__call__ (label.py:160-189): float32 restatement of the int8 output."""
    labels = np.asarray(labels)
    boundaries = _find_boundaries_thick(labels).astype("float32")
    boundaries[_find_boundaries_thick(labels != bg_label)] = mask_label
    if add_binary_target:
        binary = (labels != bg_label).astype("float32")
        binary[labels == mask_label] = mask_label
        return np.stack([binary, boundaries])
    return boundaries[None]


def boundary_targets_with_ignore_label(labels, ignore_label=-1, add_binary_target=False):
    """BoundaryTransformWithIgnoreLabel.__call__ (label.py:217-244)."""
    labels = np.asarray(labels)
    boundaries = _find_boundaries_thick(labels).astype("float32")
    boundaries[_find_boundaries_thick(labels == ignore_label)] = ignore_label
    if add_binary_target:
        binary = (labels != 0).astype("float32")
        binary[labels == ignore_label] = ignore_label
        return np.stack([binary, boundaries])
    return boundaries[None]


def one_hot_targets(labels, class_ids=None):
    """OneHotTransform.__call__ (label.py:339-353)."""
    labels = np.asarray(labels)
    ids = list(range(class_ids)) if isinstance(class_ids, int) else class_ids
    ids = np.unique(labels).tolist() if ids is None else ids
    return np.stack([(labels == c).astype("float32") for c in ids])


def segmentation_to_affinities(segmentation, offsets):
    """segmentation_to_affinities (loss/affinity_side_loss.py:70-89): (N, 1, *spatial) -> (N, C, *spatial) float32 affinities
    [seg[p] == seg[clamp(p + offset)]] (replication padding of the shifted copy)."""
    seg = np.asarray(segmentation)
    assert seg.shape[1] == 1
    sp = seg.shape[2:]
    grids = np.meshgrid(*[np.arange(s) for s in sp], indexing="ij")
    out = []
    for off in offsets:
        idx = tuple(np.clip(g + o, 0, s - 1) for g, o, s in zip(grids, off, sp))
        out.append((seg[:, 0][(slice(None),) + idx] == seg[:, 0]).astype("float32"))
    return np.stack(out, 1)
