"""The N > 1 path on CPU: world_size 2, gloo.  One all-reduce over the flat gradient buffer must give every rank the
mean of the per-rank gradients (what DDP gives the reference, multi_gpu_training.py:79), and replicas must stay
bit-identical after optimizer steps.  The kernels are emulated (tests/emu_backend.py); the collective plumbing in
torch-em_b200/distributed.py and the flat-buffer backward are the product code under test."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _data(rank):
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.randn((1, 1, 8, 16, 16), generator=g)
    t = (torch.rand((1, 2, 8, 16, 16), generator=g) > 0.5).float()
    return x, t


def _make_model():
    import torch_em_b200 as tb
    from tests.emu_backend import TorchEmuBackend
    torch.manual_seed(0)
    net = tb.UNet3d(1, 2, depth=2, initial_features=4, final_activation="Sigmoid")
    net._backend_override = TorchEmuBackend()
    return net


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch_em_b200 as tb
    from oracle import dice as odice
    torch.set_num_threads(1)
    net = _make_model()
    if rank == 1:                                   # perturb: broadcast must bring rank 1 back to rank 0's weights
        with torch.no_grad():
            for p in net.parameters():
                p.add_(1.0)
    tb.distributed.broadcast_parameters(net)
    tb.distributed.sync_gradients(net)
    opt = torch.optim.AdamW(net.parameters(), lr=1e-3)
    x, t = _data(rank)
    sums = []
    for step in range(2):
        opt.zero_grad()
        loss = odice.dice_loss(net(x), t)
        loss.backward()
        if step == 0:
            torch.save({k: p.grad.clone() for k, p in net.named_parameters()}, os.path.join(out_dir, f"grads{rank}.pt"))
        opt.step()
        sums.append(tb.distributed.parameter_checksum(net))
    torch.save(sums, os.path.join(out_dir, f"sums{rank}.pt"))
    dist.destroy_process_group()


def test_flat_allreduce_world2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    g0, g1 = torch.load(tmp_path / "grads0.pt"), torch.load(tmp_path / "grads1.pt")
    # single-process reference: mean of the two ranks' local gradients
    from oracle import dice as odice
    local = []
    for rank in range(2):
        net = _make_model()
        x, t = _data(rank)
        odice.dice_loss(net(x), t).backward()
        local.append({k: p.grad.clone() for k, p in net.named_parameters()})
    for k in g0:
        assert torch.equal(g0[k], g1[k]), k                       # every rank holds the same averaged gradient
        np.testing.assert_allclose(g0[k].numpy(), 0.5 * (local[0][k] + local[1][k]).numpy(), rtol=1e-5, atol=1e-8, err_msg=k)
    s0, s1 = torch.load(tmp_path / "sums0.pt"), torch.load(tmp_path / "sums1.pt")
    assert s0 == s1                                               # replicas bit-identical after each optimizer step
