"""Data-parallel training of the fused U-Net: one process per GPU, ONE gradient all-reduce per step.

Mirror of ``torch_em/multi_gpu_training.py`` (DDP wrap at :79) for this path.  The reference wraps the model in
``DistributedDataParallel`` (25 MB buckets, hook per parameter, ``find_unused_parameters=True`` graph walk).  Here the
backward pass already writes every parameter gradient into one contiguous fp32 buffer (``engine.FlatGrads``), so the
exchange is a single ``all_reduce`` (NCCL over NVLink / NVSwitch; gloo in the CPU tests) over that buffer, issued at
the end of the network's backward -- no other collective, no per-parameter hooks.  Samples are independent under
InstanceNorm / GroupNorm, so replicas stay bit-identical: identical averaged gradients feed identical AdamW updates.
"""
import torch
import torch.distributed as dist

__all__ = ["sync_gradients", "broadcast_parameters", "parameter_checksum"]


def broadcast_parameters(model, src=0, group=None):
    """One-time setup (what DDP's constructor does implicitly): every rank starts from rank ``src``'s weights."""
    with torch.no_grad():
        flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
        dist.broadcast(flat, src=src, group=group)
        off = 0
        for p in model.parameters():
            p.copy_(flat[off:off + p.numel()].view_as(p))
            off += p.numel()


def sync_gradients(model, group=None, average=True):
    """Make ``model`` average its gradients across ``group`` with one all-reduce per backward pass."""
    world = dist.get_world_size(group)

    def _sync(flat):
        if world == 1:
            return
        if average:
            flat.mul_(1.0 / world)      # pre-scale: sum of pre-scaled == mean, no overflow concern in fp32
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)

    model.grad_sync = _sync
    return model


def parameter_checksum(model):
    """Order-dependent fp64 checksum of all parameters (replica-divergence check)."""
    with torch.no_grad():
        flat = torch.cat([p.detach().reshape(-1).double() for p in model.parameters()])
        w = torch.arange(1, flat.numel() + 1, dtype=torch.float64, device=flat.device)
        return float((flat * (w % 1009)).sum())
