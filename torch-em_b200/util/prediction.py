"""Tiled prediction with a halo: ``predict_with_halo`` / ``predict_with_halo_pipelined`` with torch-em's signatures
(util/prediction.py:145-164, 487-510), device-resident.

The reference moves every haloed block host -> device and every prediction device -> host (prediction.py:273,279) and does the
reflect padding, the standardisation, the inner crop and the masking with numpy on the host.  Here, when the volume fits the
device (the common case: 180 GB of HBM), it is copied to the device ONCE in its raw dtype and every per-block step is a kernel
of ``csrc/tiling.cu``:

  gather_blocks       ``_load_block`` (prediction.py:98-142): clipped haloed box + np.pad(mode="reflect") of the clipped data,
                      raw dtype -> fp32, block statistics on the fly
  standardize_blocks  ``standardize`` (transform/raw.py:40-65) with the statistics of the whole haloed block
  model forward       ``batch_size`` blocks per forward pass (the caller's autocast context is honoured, also in worker threads)
  scatter_blocks      inner crop ``[halo, halo + block.shape)``, zero outside ``mask``, write to the device-resident output

and the result comes back in one device -> host copy through pinned memory.  Every argument of the reference is honoured:
``mask`` (blocks with an empty inner mask are skipped), ``roi``, ``iter_list``, list-of-(array, channel slice) outputs,
``grid_shift`` (zero padding + final crop, prediction.py:205-222, 319-322), ``with_channels``, ``prediction_function``, several
``gpu_ids`` (block i runs on device i % n, prediction.py:249-250, one replica and one thread per device).

Host callbacks that are defined on numpy arrays -- ``skip_block``, ``postprocess``, a ``preprocess`` other than ``standardize`` --
and inputs that do not fit the device (lazy hdf5 / zarr datasets larger than the memory budget) take the streaming path: the
reference's own per-block host loop, restated, around the same device model.
"""
import ctypes
import os
from concurrent import futures
from copy import deepcopy
from typing import Any, Callable, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from .._lib import call

__all__ = ["Blocking", "predict_with_halo", "predict_with_halo_pipelined", "standardize"]

MAX_BLOCKS_PER_LAUNCH = 16          # B200EM_MAX_TILE_BLOCKS
# CUDA-event times (ms) of the last device-path call of worker 0: host -> device copy of the volume, the block loop (gather,
# standardize, forward, scatter) and the device -> host copy of the result.  Read by bench.py; diagnostic only.
last_timing = {}
_RAW_CODES = {"uint8": 0, "int8": 1, "uint16": 2, "int16": 3, "int32": 4, "uint32": 5, "float16": 6, "float32": 7, "float64": 8}
_SAME_BITS = {"uint16": "int16", "uint32": "int32"}      # dtypes torch cannot hold are shipped as their signed twin


class Blocking:
    """C-order regular grid of blocks over [begin, end) with truncated last blocks (bioimage_cpp.utils.Blocking semantics
    as used at util/prediction.py:229-234,253-255)."""

    class Block:
        __slots__ = ("begin", "end", "shape")

        def __init__(self, begin, end):
            self.begin, self.end = list(begin), list(end)
            self.shape = [e - b for b, e in zip(begin, end)]

    def __init__(self, begin: Sequence[int], end: Sequence[int], block_shape: Sequence[int]):
        if not (len(begin) == len(end) == len(block_shape)):
            raise ValueError("begin, end and block_shape must have the same length")
        if any(bs <= 0 for bs in block_shape) or any(e <= b for b, e in zip(begin, end)):
            raise ValueError("empty blocking")
        self.begin, self.end, self.block_shape = list(begin), list(end), list(block_shape)
        self.blocks_per_axis = [-(-(e - b) // bs) for b, e, bs in zip(begin, end, block_shape)]

    @property
    def number_of_blocks(self) -> int:
        n = 1
        for k in self.blocks_per_axis:
            n *= k
        return n

    def get_block(self, block_id: int) -> "Blocking.Block":
        if not 0 <= block_id < self.number_of_blocks:
            raise IndexError(block_id)
        pos = []
        for k in reversed(self.blocks_per_axis):          # C order: the last axis runs fastest
            pos.append(block_id % k)
            block_id //= k
        pos = pos[::-1]
        begin = [b + p * bs for b, p, bs in zip(self.begin, pos, self.block_shape)]
        end = [min(b + bs, e) for b, bs, e in zip(begin, self.block_shape, self.end)]
        return Blocking.Block(begin, end)


def standardize(raw, eps: float = 1e-7):
    """transform/raw.py:40-65 (default arguments): float32, population statistics of the whole array.  Accepts a numpy array
    (host path) or a tensor."""
    if torch.is_tensor(raw):
        raw = raw.to(torch.float32)
        raw = raw - raw.mean()
        return raw / (raw.std(unbiased=False) + eps)
    raw = np.asarray(raw).astype("float32")
    raw -= raw.mean()
    raw /= (raw.std() + eps)
    return raw


def _is_standardize(fn) -> bool:
    if fn is standardize:
        return True
    return getattr(fn, "__name__", None) == "standardize" and str(getattr(fn, "__module__", "")).endswith("transform.raw")


# ---- reference-literal host helpers (streaming path) -----------------------------------------------------------------------
def _load_block_host(input_, offset, block_shape, halo, with_channels=False):
    """prediction.py:98-142 on an array-like (numpy, hdf5, zarr): slice the clipped box, np.pad(mode="reflect") the rest."""
    shape = input_.shape[1:] if with_channels else input_.shape
    starts = [off - ha for off, ha in zip(offset, halo)]
    stops = [off + bs + ha for off, bs, ha in zip(offset, block_shape, halo)]
    pad_left = [max(0, -s) for s in starts]
    pad_right = [max(0, st - sh) for st, sh in zip(stops, shape)]
    bb = tuple(slice(max(0, s), min(sh, st)) for s, st, sh in zip(starts, stops, shape))
    data = np.asarray(input_[(slice(None),) + bb] if with_channels else input_[bb])
    if any(pad_left) or any(pad_right):
        width = tuple(zip(pad_left, pad_right))
        data = np.pad(data, (((0, 0),) + width) if with_channels else width, mode="reflect")
    return data


def _write_block_host(prediction, block, output, ndim, mask_block, inner_bb, postprocess):
    """prediction.py:281-309: postprocess, inner crop, zero outside the mask, write."""
    if postprocess is not None:
        prediction = postprocess(prediction)
    prediction = prediction[((slice(None),) + inner_bb) if prediction.ndim == ndim + 1 else inner_bb]
    if mask_block is not None:
        mb = np.broadcast_to(mask_block[None], prediction.shape) if prediction.ndim == ndim + 1 else mask_block
        prediction[~mb] = 0
    bb = tuple(slice(beg, end) for beg, end in zip(block.begin, block.end))
    if isinstance(output, list):
        for out, channel_slice in output:
            out[bb if out.ndim == ndim else (slice(None),) + bb] = prediction[channel_slice]
    else:
        output[((slice(None),) + bb) if output.ndim == ndim + 1 else bb] = prediction


def _first_tensor(pred):
    return pred[0] if isinstance(pred, (list, tuple)) else pred


class _Autocast:
    """The caller's autocast state, re-entered inside worker threads (autocast is thread-local)."""

    def __init__(self):
        self.enabled = torch.is_autocast_enabled("cuda")
        self.dtype = torch.get_autocast_dtype("cuda") if self.enabled else None

    def __call__(self, device):
        return torch.autocast("cuda", dtype=self.dtype, enabled=self.enabled and device.type == "cuda")


# ---- device-resident path ----------------------------------------------------------------------------------------------------
def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _ints(rows):
    flat = [int(v) for r in rows for v in r]
    return (ctypes.c_int * len(flat))(*flat)


def _to3(v, ndim, fill):
    return [fill] * (3 - ndim) + [int(x) for x in v]


def _copy_rows(dst, src, lo, hi):
    """Rows [lo, hi) of the first spatial axis of a (C, rows, ...) array, one contiguous piece per channel (plain memcpys: a
    strided tensor copy between host and device would go through temporaries)."""
    for c in range(dst.shape[0]):
        dst[c, lo:hi].copy_(src[c, lo:hi], non_blocking=True)


def _device_worker(net, dev, worker_id, n_workers, vol_np, mask_np, blocks, block_ids, block_shape, halo, ndim, with_channels,
                   standardize_blocks, prediction_function, batch_size, autocast, n_out_hint):
    """All blocks of one device: returns (host result (C_out, *spatial) float32 pinned tensor, processed block ids).

    The transfers ride on two side streams.  The volume goes up in row ranges of its first spatial axis, always ahead of the batch
    that needs them (a haloed block reads rows [begin - halo, begin + block_shape + halo) clipped to the volume: the reflect
    padding of prediction.py:98-142 mirrors the CLIPPED data), so that only the rows of the first batch are waited for; finished
    output rows -- those below the first row of every block still to come -- go down while later blocks are computed, so that
    only the last rows' copy is exposed.  ``B200EM_PREDICT_OVERLAP=0`` restores one copy up, the loop, one copy down."""
    mine = [b for b in block_ids if b % n_workers == worker_id]
    if not mine:
        return None, []
    code = _RAW_CODES[str(vol_np.dtype)]
    ship = vol_np.view(_SAME_BITS[str(vol_np.dtype)]) if str(vol_np.dtype) in _SAME_BITS else vol_np
    overlap = os.environ.get("B200EM_PREDICT_OVERLAP", "1") != "0"
    with torch.cuda.device(dev), torch.no_grad(), autocast(dev):
        main = torch.cuda.current_stream(dev)
        up, down = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        ship_t = torch.from_numpy(ship)
        vol = torch.empty(ship_t.shape, dtype=ship_t.dtype, device=dev)
        vol.record_stream(up)
        up.wait_stream(main)                                                     # the allocation may recycle memory still in use on main
        rows = vol.shape[1]
        uploaded = 0

        def upload(hi):
            """Rows [uploaded, hi) of the volume, host -> device on the side stream; the compute stream waits for them."""
            nonlocal uploaded
            hi = min(rows, hi)
            if hi <= uploaded:
                return
            with torch.cuda.stream(up):
                _copy_rows(vol, ship_t, uploaded, hi)
                done = torch.cuda.Event()
                done.record(up)
            main.wait_event(done)
            uploaded = hi

        C = vol.shape[0]
        D, H, W = _to3(vol.shape[1:], ndim, 1)
        spatial = tuple(vol.shape[1:])
        mask_d = None
        if mask_np is not None:
            mask_d = torch.from_numpy(np.ascontiguousarray(mask_np != 0).view(np.uint8)).to(dev)
            # blocks whose inner mask is empty are skipped (prediction.py:258-262): one reduction per block, ONE synchronisation
            flags = torch.stack([mask_d[tuple(slice(b, e) for b, e in zip(blocks[i].begin, blocks[i].end))].any() for i in mine])
            keep = flags.cpu().tolist()
            mine = [i for i, k in zip(mine, keep) if k]
            if not mine:
                return None, []
        bd, bh, bw = _to3([bs + 2 * ha for bs, ha in zip(block_shape, halo)], ndim, 1)
        hd, hh, hw = _to3(halo, ndim, 0)
        batches = [mine[s:s + batch_size] for s in range(0, len(mine), batch_size)]
        need = [max(blocks[i].begin[0] + block_shape[0] + halo[0] for i in ids) for ids in batches]    # volume rows a batch reads
        # output rows no later batch writes: the smallest first row of the blocks still to come
        tail, ready = rows, []
        for ids in reversed(batches):
            ready.append(tail)
            tail = min(tail, min(blocks[i].begin[0] for i in ids))
        ready.reverse()
        step_rows = -(-rows // len(batches))                                     # spread the upload evenly over the batches
        upload(need[0] if overlap else rows)
        ev[1].record()
        out_d = host = None
        copied = 0
        for s, ids in enumerate(batches):
            nb = len(ids)
            upload(need[s])
            begins = [_to3([b - ha for b, ha in zip(blocks[i].begin, halo)], ndim, 0) for i in ids]
            inp = torch.empty((nb, C, bd, bh, bw), dtype=torch.float32, device=dev)
            stats = torch.zeros((nb, 2), dtype=torch.float64, device=dev) if standardize_blocks else None
            for k in range(0, nb, MAX_BLOCKS_PER_LAUNCH):
                n_ = min(MAX_BLOCKS_PER_LAUNCH, nb - k)
                call("b200em_gather_blocks", _ptr(vol), code, C, D, H, W, _ints(begins[k:k + n_]), n_, bd, bh, bw, _ptr(inp[k:]),
                     _ptr(stats[k:]) if stats is not None else None, _stream(dev))
            if standardize_blocks:
                call("b200em_standardize_blocks", _ptr(inp), nb, C * bd * bh * bw, _ptr(stats), 1e-7, _stream(dev))
            x = inp if ndim == 3 else inp[:, :, 0]
            pred = _first_tensor(net(x) if prediction_function is None else prediction_function(net, x))
            pred = pred.to(torch.float32)
            if pred.dim() == ndim + 1:                                          # a model without a channel axis in its output
                pred = pred[:, None]
            pred = pred.contiguous()
            Cp = pred.shape[1]
            if tuple(pred.shape[2:]) != tuple(x.shape[2:]):
                raise ValueError(f"predict_with_halo: the model changed the spatial shape {tuple(x.shape[2:])} -> {tuple(pred.shape[2:])}")
            if out_d is None:
                out_d = torch.zeros((Cp,) + spatial, dtype=torch.float32, device=dev)
                out_d.record_stream(down)
                host = torch.empty(out_d.shape, dtype=torch.float32, pin_memory=True)
            obeg = [_to3(blocks[i].begin, ndim, 0) for i in ids]
            oshp = [_to3(blocks[i].shape, ndim, 1) for i in ids]
            for k in range(0, nb, MAX_BLOCKS_PER_LAUNCH):
                n_ = min(MAX_BLOCKS_PER_LAUNCH, nb - k)
                call("b200em_scatter_blocks", _ptr(pred[k:]), Cp, bd, bh, bw, hd, hh, hw, _ints(obeg[k:k + n_]), _ints(oshp[k:k + n_]), n_,
                     _ptr(out_d), D, H, W, 0, Cp, _ptr(mask_d) if mask_d is not None else None, _stream(dev))
            last = s + 1 == len(batches)
            if last:
                ev[2].record()
            if overlap or last:
                hi = rows if last else ready[s]
                if hi > copied:                                                  # finished output rows go down behind this batch
                    down.wait_stream(main)
                    with torch.cuda.stream(down):
                        _copy_rows(host, out_d, copied, hi)
                    copied = hi
            if not last:                                                         # the next batches' rows go up while this one computes
                upload(max(need[s + 1], uploaded + step_rows))
        main.wait_stream(down)
        ev[3].record()
        main.synchronize()
        if worker_id == 0:
            last_timing.update(h2d_ms=ev[0].elapsed_time(ev[1]), loop_ms=ev[1].elapsed_time(ev[2]), d2h_ms=ev[2].elapsed_time(ev[3]),
                               blocks=len(mine), batch_size=batch_size, overlap=overlap)
    return host, mine


def _device_budget_ok(devices, in_bytes, out_bytes, block_bytes):
    need = in_bytes + out_bytes + 64 * block_bytes          # volume + output + a generous bound on the activations of one batch
    for d in devices:
        free, _ = torch.cuda.mem_get_info(d)
        if need > 0.85 * free:
            return False
    return True


def _run(input_, model, gpu_ids, block_shape, halo, output, preprocess, postprocess, with_channels, skip_block, mask, prediction_function,
         roi, iter_list, grid_shift, batch_size):
    devices = [torch.device("cuda", g) if isinstance(g, int) else torch.device(g) for g in gpu_ids]
    if not devices:
        raise ValueError("gpu_ids must name at least one device")
    models = [(model if next(model.parameters()).device == d else deepcopy(model).to(d), d) for d in devices]
    n_workers = len(devices)
    shape0 = tuple(input_.shape)
    spatial0 = shape0[1:] if with_channels else shape0
    ndim = len(spatial0)
    if not (len(block_shape) == len(halo) == ndim):
        raise ValueError("block_shape and halo must have one entry per spatial axis")

    # grid_shift: zero padding to the left + final crop (prediction.py:205-222, 241-247, 319-322)
    input_eff, mask_eff = input_, mask
    pad_left = (0,) * ndim
    if grid_shift is not None:
        assert len(grid_shift) == ndim, "grid_shift must match number of spatial dims"
        if output is not None:
            raise ValueError(
                "grid_shift is not supported together with a user-provided `output`, because grid_shift requires internal "
                "zero-padding and a final cropping step. Pass `output=None` (let this function allocate the output) or disable "
                "`grid_shift`. Or pad the input manually beforehand.")
        if not isinstance(input_eff, np.ndarray):
            raise TypeError("grid_shift padding currently requires input_ to be a numpy array")
        pad_left = tuple(int(np.rint(abs(gs) * bs)) for gs, bs in zip(grid_shift, block_shape))
        width = tuple((p, 0) for p in pad_left)
        input_eff = np.pad(input_eff, (((0, 0),) + width) if with_channels else width, mode="constant", constant_values=0)
        if mask_eff is not None:
            if not isinstance(mask_eff, np.ndarray):
                raise TypeError("grid_shift padding currently requires mask to be a numpy array")
            mask_eff = np.pad(mask_eff, width, mode="constant", constant_values=0)
    spatial = tuple(input_eff.shape[1:] if with_channels else input_eff.shape)

    if roi is None:
        blocking = Blocking([0] * ndim, list(spatial), list(block_shape))
    else:
        assert len(roi) == ndim
        blocking = Blocking([0 if r.start is None else r.start for r in roi],
                            [sh if r.stop is None else r.stop for r, sh in zip(roi, spatial)], list(block_shape))
    n_blocks = blocking.number_of_blocks
    block_ids = list(range(n_blocks)) if iter_list is None else [int(i) for i in iter_list]
    blocks = {i: blocking.get_block(i) for i in block_ids}

    if batch_size is None:
        # predict_with_halo has no batch argument (the reference feeds one block per forward pass).  On the device path several
        # haloed blocks share one forward pass -- the per-block results are identical (the norm layers are per sample) and the
        # host-side launch overhead of a forward pass is amortised -- unless a user prediction_function could tell the difference.
        big_vox = int(np.prod([bs + 2 * ha for bs, ha in zip(block_shape, halo)]))
        batch_size = 1 if prediction_function is not None else max(1, min(4, (32 << 20) // max(big_vox, 1)))
    own_output = output is None
    n_out_hint = getattr(models[0][0], "out_channels", None)
    autocast = _Autocast()

    # ---- which path -------------------------------------------------------------------------------------------------------
    on_device = (all(d.type == "cuda" for d in devices) and ndim in (2, 3) and skip_block is None and postprocess is None
                 and (preprocess is None or _is_standardize(preprocess)))
    vol_np = None
    if on_device:
        n_out_est = n_out_hint if isinstance(n_out_hint, int) else (max(n_out_hint) if n_out_hint else 4)
        n_vox = int(np.prod(spatial))
        itemsize = np.dtype(input_eff.dtype).itemsize
        block_bytes = 4 * batch_size * int(np.prod([bs + 2 * ha for bs, ha in zip(block_shape, halo)]))
        in_ch = shape0[0] if with_channels else 1
        on_device = _device_budget_ok(devices, n_vox * in_ch * itemsize, 4 * n_out_est * n_vox, block_bytes)
    if on_device:
        vol_np = np.asarray(input_eff if isinstance(input_eff, np.ndarray) else input_eff[...])
        if str(vol_np.dtype) not in _RAW_CODES:
            vol_np = vol_np.astype("uint8" if vol_np.dtype == bool else "float32")
        vol_np = np.ascontiguousarray(vol_np if with_channels else vol_np[None])
        mask_np = None if mask_eff is None else np.asarray(mask_eff if isinstance(mask_eff, np.ndarray) else mask_eff[...])

        def work(w):
            net, dev = models[w]
            return _device_worker(net, dev, w, n_workers, vol_np, mask_np, blocks, block_ids, block_shape, halo, ndim, with_channels,
                                  preprocess is not None, prediction_function, batch_size, autocast, n_out_hint)

        if n_workers == 1:
            results = [work(0)]
        else:
            with futures.ThreadPoolExecutor(n_workers) as tp:
                results = list(tp.map(work, range(n_workers)))
        results = [(h, ids) for h, ids in results if h is not None]
        if own_output and len(results) == 1 and grid_shift is None:
            return results[0][0].numpy()                     # backed by the pinned buffer of the single device -> host copy
        if own_output:
            n_out = results[0][0].shape[0] if results else (n_out_hint if isinstance(n_out_hint, int) else 1)
            output = np.zeros((n_out,) + spatial, dtype="float32")
        for host, ids in results:
            res = host.numpy()
            for i in ids:                                    # only processed blocks are written, like the reference
                b = blocks[i]
                bb = tuple(slice(beg, end) for beg, end in zip(b.begin, b.end))
                pred = res[(slice(None),) + bb]
                if isinstance(output, list):
                    for out, channel_slice in output:
                        out[bb if out.ndim == ndim else (slice(None),) + bb] = pred[channel_slice]
                else:
                    output[((slice(None),) + bb) if output.ndim == ndim + 1 else bb] = pred
    else:
        # ---- streaming path: the reference's per-block host loop (prediction.py:249-317) around the device model ----------------
        if own_output:
            n_out = n_out_hint if isinstance(n_out_hint, int) else n_out_hint[0]
            output = np.zeros((n_out,) + spatial, dtype="float32")

        def predict_block(block_id):
            net, dev = models[block_id % n_workers]
            block = blocks[block_id]
            inner_bb = tuple(slice(ha, ha + bs) for ha, bs in zip(halo, block.shape))
            mask_block = None
            if mask_eff is not None:
                mask_block = _load_block_host(mask_eff, block.begin, block_shape, halo)[inner_bb].astype("bool")
                if mask_block.sum() == 0:
                    return
            inp = _load_block_host(input_eff, block.begin, block_shape, halo, with_channels)
            if skip_block is not None and skip_block(inp):
                return
            if preprocess is not None:
                inp = preprocess(inp)
            inp = np.ascontiguousarray(inp[None] if with_channels else inp[None, None])
            if inp.dtype not in (np.float32, np.float64, np.float16):
                inp = inp.astype("float32")
            with torch.no_grad(), autocast(dev):
                x = torch.from_numpy(inp).to(dev)
                pred = _first_tensor(net(x) if prediction_function is None else prediction_function(net, x))
                pred = pred.float().cpu().numpy().squeeze(0)
            _write_block_host(pred, block, output, ndim, mask_block, inner_bb, postprocess)

        if n_workers == 1:
            for i in block_ids:
                predict_block(i)
        else:
            with futures.ThreadPoolExecutor(n_workers) as tp:
                list(tp.map(predict_block, block_ids))

    if grid_shift is not None:
        crop = tuple(slice(p, p + s) for p, s in zip(pad_left, spatial0))
        output = output[((slice(None),) + crop) if output.ndim == ndim + 1 else crop]
    return output


def predict_with_halo(
    input_,
    model: torch.nn.Module,
    gpu_ids: List[Union[str, int]],
    block_shape: Tuple[int, ...],
    halo: Tuple[int, ...],
    output=None,
    preprocess: Optional[Callable] = standardize,
    postprocess: Optional[Callable] = None,
    with_channels: bool = False,
    skip_block: Optional[Callable[[Any], bool]] = None,
    mask=None,
    disable_tqdm: bool = False,
    tqdm_desc: str = "predict with halo",
    prediction_function: Optional[Callable] = None,
    roi: Optional[Tuple[slice]] = None,
    iter_list: Optional[List[int]] = None,
    grid_shift: Optional[Tuple[float, ...]] = None,
):
    """Block-wise network prediction with a halo; same signature and results as
    ``torch_em.util.prediction.predict_with_halo`` (prediction.py:145-324).  See the module docstring for what runs where.
    ``disable_tqdm`` / ``tqdm_desc`` are accepted for compatibility (there is no per-block host loop to report on)."""
    return _run(input_, model, gpu_ids, block_shape, halo, output, preprocess, postprocess, with_channels, skip_block, mask,
                prediction_function, roi, iter_list, grid_shift, batch_size=None)


def predict_with_halo_pipelined(
    input_,
    model: torch.nn.Module,
    gpu_ids: List[Union[str, int]],
    block_shape: Tuple[int, ...],
    halo: Tuple[int, ...],
    output=None,
    preprocess: Optional[Callable] = standardize,
    postprocess: Optional[Callable] = None,
    with_channels: bool = False,
    skip_block: Optional[Callable[[Any], bool]] = None,
    mask=None,
    disable_tqdm: bool = False,
    tqdm_desc: str = "predict with halo (pipelined)",
    prediction_function: Optional[Callable] = None,
    roi: Optional[Tuple[slice]] = None,
    iter_list: Optional[List[int]] = None,
    batch_size: int = 1,
    num_prefetch_workers: int = 4,
    queue_size: Optional[int] = None,
    num_write_workers: int = 1,
    write_queue_size: Optional[int] = None,
    grid_shift: Optional[Tuple[float, ...]] = None,
):
    """Same signature and results as ``torch_em.util.prediction.predict_with_halo_pipelined`` (prediction.py:487-759).  The
    reference pipelines host loading, device prediction and host writing through queues; with the volume and the output
    resident on the device there is nothing left to pipeline, so the queue / worker-count arguments are accepted and unused and
    ``batch_size`` blocks share one forward pass.  ``prediction_function`` must operate on the leading batch axis."""
    if grid_shift is not None:
        raise NotImplementedError(
            "grid_shift is not supported by predict_with_halo_pipelined. "
            "Use predict_with_halo for grid_shift, or pre-pad the input and use roi.")
    return _run(input_, model, gpu_ids, block_shape, halo, output, preprocess, postprocess, with_channels, skip_block, mask,
                prediction_function, roi, iter_list, None, batch_size=max(1, int(batch_size)))
