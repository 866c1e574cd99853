"""Data-parallel training of the fused U-Net: one process per GPU, the gradient all-reduce of ONE flat buffer per step.

Mirror of ``torch_em/multi_gpu_training.py`` (DDP wrap at :79) for this path.  The reference wraps the model in
``DistributedDataParallel`` (25 MB buckets, hook per parameter, ``find_unused_parameters=True`` graph walk).  Here the
backward pass already writes every parameter gradient into one contiguous fp32 buffer (``engine.FlatGrads``), so the
exchange is an ``all_reduce`` (NCCL over NVLink / NVSwitch; gloo in the CPU tests) over that buffer and no other collective,
no per-parameter hooks.  The buffer is exchanged in a few large contiguous pieces launched as soon as they are final (see
``engine.backward_pass``): everything from the base block to the end of the buffer -- ~95 % of the parameters -- is done while
the most expensive part of the backward pass (the shallow encoder levels) is still ahead, so the collective hides behind it
(85 MB at cfg2, 1.37 GB at cfg4).  Samples are independent under InstanceNorm / GroupNorm, so replicas stay bit-identical:
identical averaged gradients feed identical AdamW updates.
"""
import torch
import torch.distributed as dist

__all__ = ["sync_gradients", "broadcast_parameters", "parameter_checksum", "FlatGradSync"]


def broadcast_parameters(model, src=0, group=None):
    """One-time setup (what DDP's constructor does implicitly): every rank starts from rank ``src``'s weights (and buffers)."""
    with torch.no_grad():
        tensors = list(model.parameters()) + [b for b in model.buffers() if b.is_floating_point()]
        flat = torch.cat([p.detach().reshape(-1).float() for p in tensors])
        dist.broadcast(flat, src=src, group=group)
        off = 0
        for p in tensors:
            p.copy_(flat[off:off + p.numel()].view_as(p))
            off += p.numel()


class FlatGradSync:
    """Averages the flat gradient buffer across ``group`` in contiguous pieces that are launched (asynchronously, on the
    collective's own stream) as the backward pass reports them final, and waited for when the backward pass ends.

    Adjacent ready ranges are merged until a piece has at least ``min_chunk_bytes``: collectives over NVSwitch are latency-,
    not link-bound, so few large pieces are right."""

    def __init__(self, group=None, average=True, min_chunk_bytes=32 << 20):
        self.group, self.average, self.min_chunk = group, average, int(min_chunk_bytes)
        self.world = dist.get_world_size(group)
        self.launched = 0                      # collectives issued in the last backward pass (tests / bench evidence)

    def begin(self, flat):
        self.flat, self.works, self.pending, self.launched = flat, [], None, 0

    def _launch(self, lo, hi):
        piece = self.flat[lo:hi]
        if self.average:
            piece.mul_(1.0 / self.world)       # pre-scale: sum of pre-scaled == mean, no overflow concern in fp32
        self.works.append(dist.all_reduce(piece, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        self.launched += 1

    def ready(self, lo, hi):
        if self.world == 1 or hi <= lo:
            return
        if self.pending is not None and self.pending[0] == hi:
            self.pending = (lo, self.pending[1])              # ranges arrive in descending address order
        else:
            if self.pending is not None:
                self._launch(*self.pending)
            self.pending = (lo, hi)
        if (self.pending[1] - self.pending[0]) * 4 >= self.min_chunk:
            self._launch(*self.pending)
            self.pending = None

    def finish(self):
        if self.pending is not None:
            self._launch(*self.pending)
            self.pending = None
        for w in self.works:
            w.wait()                            # the current stream waits for the collective's stream
        self.works = []

    def __call__(self, flat):                   # plain-callable form: the whole buffer at once
        self.begin(flat)
        self.ready(0, flat.numel())
        self.finish()


def sync_gradients(model, group=None, average=True, min_chunk_bytes=32 << 20):
    """Make ``model`` average its gradients across ``group`` during every backward pass (overlapped flat-buffer all-reduce)."""
    model.grad_sync = FlatGradSync(group, average, min_chunk_bytes)
    return model


def parameter_checksum(model):
    """Order-dependent fp64 checksum of all parameters (replica-divergence check)."""
    with torch.no_grad():
        flat = torch.cat([p.detach().reshape(-1).double() for p in model.parameters()])
        w = torch.arange(1, flat.numel() + 1, dtype=torch.float64, device=flat.device)
        return float((flat * (w % 1009)).sum())
