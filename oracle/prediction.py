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


def block_grid(begin, end, block_shape):
    """bioimage_cpp.utils.Blocking(begin, end, block_shape) restated: C-order list of (block begin, block end)."""
    grid = [range(b, e, bs) for b, e, bs in zip(begin, end, block_shape)]
    return [(list(bg), [min(b + bs, e) for b, bs, e in zip(bg, block_shape, end)]) for bg in itertools.product(*grid)]


def predict_with_halo(input_, net, block_shape, halo, n_out, preprocess=standardize, with_channels=False, output=None,
                      postprocess=None, skip_block=None, mask=None, roi=None, iter_list=None, grid_shift=None):
    """prediction.py:190-324.  net: callable (1, C, *spatial) float32 numpy -> (1, n_out, *spatial) numpy."""
    shape0 = input_.shape[1:] if with_channels else input_.shape
    ndim = len(shape0)
    pad_left = (0,) * ndim
    if grid_shift is not None:                                   # prediction.py:205-222: zero padding to the left
        pad_left = tuple(int(np.rint(abs(gs) * bs)) for gs, bs in zip(grid_shift, block_shape))
        width = tuple((p, 0) for p in pad_left)
        input_ = np.pad(input_, (((0, 0),) + width) if with_channels else width, mode="constant", constant_values=0)
        if mask is not None:
            mask = np.pad(mask, width, mode="constant", constant_values=0)
    shape = input_.shape[1:] if with_channels else input_.shape
    if roi is None:
        blocks = block_grid([0] * ndim, list(shape), block_shape)
    else:
        blocks = block_grid([0 if r.start is None else r.start for r in roi],
                            [sh if r.stop is None else r.stop for r, sh in zip(roi, shape)], block_shape)
    if output is None:
        output = np.zeros((n_out,) + tuple(shape), dtype="float32")
    for block_id in (range(len(blocks)) if iter_list is None else iter_list):
        begin, end = blocks[block_id]
        bshape = [e - b for b, e in zip(begin, end)]
        inner = tuple(slice(ha, ha + bs) for ha, bs in zip(halo, bshape))
        mask_block = None
        if mask is not None:
            mask_block = load_block(mask, begin, block_shape, halo)[inner].astype("bool")
            if mask_block.sum() == 0:
                continue
        inp = load_block(input_, begin, block_shape, halo, with_channels)
        if skip_block is not None and skip_block(inp):
            continue
        if preprocess is not None:
            inp = preprocess(inp)
        inp = inp[None] if with_channels else inp[None, None]
        pred = net(inp.astype("float32"))[0]
        if postprocess is not None:
            pred = postprocess(pred)
        pred = pred[((slice(None),) + inner) if pred.ndim == ndim + 1 else inner]
        if mask_block is not None:
            pred[~(np.broadcast_to(mask_block[None], pred.shape) if pred.ndim == ndim + 1 else mask_block)] = 0
        bb = tuple(slice(b, e) for b, e in zip(begin, end))
        if isinstance(output, list):
            for out, channel_slice in output:
                out[bb if out.ndim == ndim else (slice(None),) + bb] = pred[channel_slice]
        else:
            output[((slice(None),) + bb) if output.ndim == ndim + 1 else bb] = pred
    if grid_shift is not None:
        crop = tuple(slice(p, p + s) for p, s in zip(pad_left, shape0))
        output = output[((slice(None),) + crop) if output.ndim == ndim + 1 else crop]
    return output
