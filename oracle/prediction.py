"""TEST INFRASTRUCTURE ONLY -- CPU restatement of torch-em's block-wise prediction (util/prediction.py).

numpy, plain loops; follows ``_load_block`` (prediction.py:98-142) and the per-block body of ``predict_with_halo``
(prediction.py:249-309) with ``bioimage_cpp.utils.Blocking`` restated as a C-order regular grid with truncated last blocks
(the library is absent here and unpinned in the reference's setup.py:8; the reference's own tests pin coverage / shape only,
test/util/test_prediction.py:20-31 -- "parity unpinned" for the block order, which does not change the result because the
blocks are disjoint).
"""
import itertools

import numpy as np


def standardize(raw, eps=1e-7):
    """transform/raw.py:40-65."""
    raw = raw.astype("float32")
    raw = raw - raw.mean()
    return raw / (raw.std() + eps)


def load_block(input_, offset, block_shape, halo, with_channels=False):
    """prediction.py:98-142: clip the haloed bounding box to the volume, np.pad(mode="reflect") what is missing."""
    shape = input_.shape[1:] if with_channels else input_.shape
    starts = [off - ha for off, ha in zip(offset, halo)]
    stops = [off + bs + ha for off, bs, ha in zip(offset, block_shape, halo)]
    pad_left = [max(0, -s) for s in starts]
    pad_right = [max(0, st - sh) for st, sh in zip(stops, shape)]
    bb = tuple(slice(max(0, s), min(sh, st)) for s, st, sh in zip(starts, stops, shape))
    data = input_[(slice(None),) + bb] if with_channels else input_[bb]
    if any(pad_left) or any(pad_right):
        width = tuple(zip(pad_left, pad_right))
        if with_channels:
            width = ((0, 0),) + width
        data = np.pad(data, width, mode="reflect")
    return data


def predict_with_halo(input_, net, block_shape, halo, n_out, preprocess=standardize, with_channels=False):
    """net: callable (1, C, *spatial) float32 numpy -> (1, n_out, *spatial) numpy."""
    shape = input_.shape[1:] if with_channels else input_.shape
    ndim = len(shape)
    output = np.zeros((n_out,) + tuple(shape), dtype="float32")
    grid = [range(0, sh, bs) for sh, bs in zip(shape, block_shape)]
    for begin in itertools.product(*grid):                       # C order
        end = [min(b + bs, sh) for b, bs, sh in zip(begin, block_shape, shape)]
        bshape = [e - b for b, e in zip(begin, end)]
        inp = load_block(input_, begin, block_shape, halo, with_channels)
        if preprocess is not None:
            inp = preprocess(inp)
        inp = inp[None] if with_channels else inp[None, None]
        pred = net(inp.astype("float32"))[0]
        inner = (slice(None),) + tuple(slice(ha, ha + bs) for ha, bs in zip(halo, bshape))
        output[(slice(None),) + tuple(slice(b, e) for b, e in zip(begin, end))] = pred[inner]
    return output
