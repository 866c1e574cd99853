"""``train_multi_gpu`` with torch-em's signature (torch_em/multi_gpu_training.py:107-190): data-parallel training on all local
GPUs, one process per GPU, for the fused U-Net.

Differences from the reference, by design:
  * no ``DistributedDataParallel`` wrapper (multi_gpu_training.py:79): the model's backward already produces ONE flat gradient
    buffer, which ``distributed.sync_gradients`` all-reduces in a few large pieces overlapped with the backward pass; the
    trainer therefore sees the plain model (attributes such as ``init_kwargs`` need no ``__getattr__`` forwarding, :41-52).
    ``find_unused_parameters`` is accepted and ignored (there is no autograd graph walk to configure);
  * the rendezvous address is 127.0.0.1 with a free port picked per launch instead of the fixed localhost:12355 (:13-18);
  * two extra optional arguments, ``world_size`` and ``backend`` (defaults: all visible GPUs, "nccl"), so that the host logic can
    be tested with gloo on CPUs.
The trainer is ``torch_em.default_segmentation_trainer`` unless ``trainer_callable`` is given, exactly like the reference.
"""
import os
import socket
from functools import partial
from typing import Any, Callable, Dict, Optional

import torch
import torch.distributed as dist
import torch.utils.data

from . import distributed as _dist

__all__ = ["train_multi_gpu"]


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def setup(rank, world_size, port, backend):
    """@private"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group(backend, rank=rank, world_size=world_size)


def cleanup():
    """@private"""
    dist.destroy_process_group()


def _create_data_loader(ds_callable, ds_kwargs, loader_kwargs, world_size, rank):
    """multi_gpu_training.py:27-40: the shuffle flag moves from the loader to a DistributedSampler that shards the dataset."""
    ds = ds_callable(**ds_kwargs)
    loader_kwargs = dict(loader_kwargs)
    shuffle = loader_kwargs.pop("shuffle", False)
    sampler = torch.utils.data.distributed.DistributedSampler(ds, num_replicas=world_size, rank=rank, shuffle=shuffle)
    loader = torch.utils.data.DataLoader(ds, sampler=sampler, **loader_kwargs)
    loader.shuffle = shuffle
    return loader


def _train_impl(rank, world_size, port, backend, model_callable, model_kwargs, train_dataset_callable, train_dataset_kwargs,
                val_dataset_callable, val_dataset_kwargs, loader_kwargs, iterations, find_unused_parameters=True,
                optimizer_callable=None, optimizer_kwargs=None, lr_scheduler_callable=None, lr_scheduler_kwargs=None,
                trainer_callable=None, **kwargs):
    assert "device" not in kwargs
    print(f"Running data-parallel training on rank {rank}.")
    setup(rank, world_size, port, backend)
    device = torch.device("cuda", rank) if backend == "nccl" else torch.device("cpu")
    if device.type == "cuda":
        torch.cuda.set_device(device)
    model = model_callable(**model_kwargs).to(device)
    _dist.broadcast_parameters(model)              # what DDP's constructor does implicitly
    _dist.sync_gradients(model)                    # flat-buffer all-reduce inside every backward pass

    if optimizer_callable is not None:
        optimizer = optimizer_callable(model.parameters(), **(optimizer_kwargs or {}))
        kwargs["optimizer"] = optimizer
        if lr_scheduler_callable is not None:
            kwargs["lr_scheduler"] = lr_scheduler_callable(optimizer, **(lr_scheduler_kwargs or {}))

    train_loader = _create_data_loader(train_dataset_callable, train_dataset_kwargs, loader_kwargs, world_size, rank)
    val_loader = _create_data_loader(val_dataset_callable, val_dataset_kwargs, loader_kwargs, world_size, rank)

    if trainer_callable is None:
        import torch_em
        trainer_callable = torch_em.default_segmentation_trainer

    trainer = trainer_callable(model=model, train_loader=train_loader, val_loader=val_loader, device=device, rank=rank, **kwargs)
    trainer.fit(iterations=iterations)
    cleanup()


def train_multi_gpu(
    model_callable: Callable[[Any], torch.nn.Module],
    model_kwargs: Dict[str, Any],
    train_dataset_callable: Callable[[Any], torch.utils.data.Dataset],
    train_dataset_kwargs: Dict[str, Any],
    val_dataset_callable: Callable[[Any], torch.utils.data.Dataset],
    val_dataset_kwargs: Dict[str, Any],
    loader_kwargs: Dict[str, Any],
    iterations: int,
    find_unused_parameters: bool = True,
    optimizer_callable: Optional[Callable[[Any], torch.optim.Optimizer]] = None,
    optimizer_kwargs: Optional[Dict[str, Any]] = None,
    lr_scheduler_callable: Optional[Callable] = None,
    lr_scheduler_kwargs: Optional[Dict[str, Any]] = None,
    trainer_callable: Optional[Callable] = None,
    world_size: Optional[int] = None,
    backend: str = "nccl",
    **kwargs,
) -> None:
    """Run data-parallel training on multiple local GPUs; same arguments as ``torch_em.multi_gpu_training.train_multi_gpu``
    (multi_gpu_training.py:107-190) plus ``world_size`` / ``backend``.  ``kwargs`` go to the trainer."""
    if world_size is None:
        world_size = torch.cuda.device_count()
    if world_size < 1:
        raise RuntimeError("train_multi_gpu: no CUDA device visible (pass world_size and backend='gloo' for a CPU dry run)")
    train = partial(
        _train_impl, world_size=world_size, port=_free_port(), backend=backend, model_callable=model_callable,
        model_kwargs=model_kwargs, train_dataset_callable=train_dataset_callable, train_dataset_kwargs=train_dataset_kwargs,
        val_dataset_callable=val_dataset_callable, val_dataset_kwargs=val_dataset_kwargs, loader_kwargs=loader_kwargs,
        iterations=iterations, find_unused_parameters=find_unused_parameters, optimizer_callable=optimizer_callable,
        optimizer_kwargs=optimizer_kwargs, lr_scheduler_callable=lr_scheduler_callable, lr_scheduler_kwargs=lr_scheduler_kwargs,
        trainer_callable=trainer_callable, **kwargs)
    torch.multiprocessing.spawn(train, nprocs=world_size, join=True)
