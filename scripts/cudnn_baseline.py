"""Diagnostic only (not the product, not the oracle): the cfg2 train step written with stock torch.nn modules
(cuDNN Conv3d, InstanceNorm3d, MaxPool3d, F.interpolate) in the reference's block structure (unet.py:409-458), timed on
the same GPU under bf16 autocast.  This is the "reference cuDNN" denominator of BASELINE.json's >= 1.5x target; the
reference itself cannot be imported on the GPU box (/root/reference does not travel).

    python scripts/cudnn_baseline.py [--channels-last] [--steps 10]
"""
import argparse
import json

import torch
import torch.nn as nn
import torch.nn.functional as F


class Block(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.block = nn.Sequential(nn.InstanceNorm3d(cin), nn.Conv3d(cin, cout, 3, padding=1), nn.ReLU(inplace=True),
                                   nn.InstanceNorm3d(cout), nn.Conv3d(cout, cout, 3, padding=1), nn.ReLU(inplace=True))

    def forward(self, x):
        return self.block(x)


class Net(nn.Module):
    def __init__(self, cin=1, cout=2, depth=4, f0=32):
        super().__init__()
        fe = [cin] + [f0 * 2 ** i for i in range(depth)]
        self.enc = nn.ModuleList([Block(fe[i], fe[i + 1]) for i in range(depth)])
        self.base = Block(fe[-1], fe[-1] * 2)
        fd = [fe[-1] * 2] + fe[1:][::-1]
        self.samp = nn.ModuleList([nn.Conv3d(fd[i], fd[i + 1], 1) for i in range(depth)])
        self.dec = nn.ModuleList([Block(2 * fd[i + 1], fd[i + 1]) for i in range(depth)])
        self.out = nn.Conv3d(f0, cout, 1)

    def forward(self, x):
        skips = []
        for b in self.enc:
            x = b(x)
            skips.append(x)
            x = F.max_pool3d(x, 2)
        x = self.base(x)
        for s, b, sk in zip(self.samp, self.dec, skips[::-1]):
            x = s(F.interpolate(x, scale_factor=2, mode="trilinear", align_corners=False))
            x = b(torch.cat([x, sk], dim=1))
        return torch.sigmoid(self.out(x))


def dice_loss(p, t, eps=1e-7):
    C = p.shape[1]
    p = p.transpose(0, 1).reshape(C, -1)
    t = t.transpose(0, 1).reshape(C, -1)
    num = (p * t).sum(-1)
    den = (p * p).sum(-1) + (t * t).sum(-1)
    return (1 - 2 * num / den.clamp(min=eps)).sum()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--channels-last", action="store_true")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--size", type=int, default=128)
    args = ap.parse_args()
    dev = "cuda:0"
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0)
    net = Net().to(dev)
    if args.channels_last:
        net = net.to(memory_format=torch.channels_last_3d)
    opt = torch.optim.AdamW(net.parameters(), lr=1e-3)
    S = args.size
    x = torch.randn(args.batch, 1, S, S, S, device=dev)
    t = (torch.rand(args.batch, 2, S, S, S, device=dev) > 0.5).float()
    if args.channels_last:
        x = x.contiguous(memory_format=torch.channels_last_3d)

    def step():
        opt.zero_grad()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = dice_loss(net(x), t)
        loss.backward()
        opt.step()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print(json.dumps({"impl": "torch.nn cuDNN bf16 autocast", "channels_last": args.channels_last, "ms_per_step": ms,
                      "voxels_per_s": args.batch * S ** 3 / (ms * 1e-3), "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}))


if __name__ == "__main__":
    main()
