"""Timing of the HBM-bound kernels on the cfg2 shapes (CUDA events).  Diagnostic.

    python scripts/bench_elementwise.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from torch_em_b200.backend import default_backend

dev = "cuda:0"
B = default_backend()
N = 4


def timeit(name, fn, nbytes, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:44s} {ms * 1e3:9.1f} us  {nbytes / ms / 1e6:8.1f} GB/s (algorithmic)", flush=True)


for C, S in ((32, 128), (64, 64)):
    lo = torch.randn((N, S // 2, S // 2, S // 2, C), device=dev).bfloat16()
    cat = torch.empty((N, S, S, S, 2 * C), device=dev, dtype=torch.bfloat16)
    flat = torch.empty((N, S, S, S, C), device=dev, dtype=torch.bfloat16)
    sums = torch.zeros((N, C, 2), device=dev)
    out_bytes = flat.numel() * 2
    timeit(f"upsample_fwd C={C} S={S} into cat slice", lambda: B.upsample_fwd(lo, cat[..., :C], (2, 2, 2), sums), out_bytes + lo.numel() * 2)
    timeit(f"upsample_fwd C={C} S={S} contiguous", lambda: B.upsample_fwd(lo, flat, (2, 2, 2), sums), out_bytes + lo.numel() * 2)
    timeit(f"upsample_fwd C={C} S={S} contiguous, no stats", lambda: B.upsample_fwd(lo, flat, (2, 2, 2), None), out_bytes + lo.numel() * 2)
    dlo = torch.empty_like(lo)
    timeit(f"upsample_bwd C={C} S={S} from cat slice", lambda: B.upsample_bwd(cat[..., :C], dlo, (2, 2, 2)), out_bytes + lo.numel() * 2)
    timeit(f"upsample_bwd C={C} S={S} contiguous", lambda: B.upsample_bwd(flat, dlo, (2, 2, 2)), out_bytes + lo.numel() * 2)
    skip = cat[..., C:]
    pooled = torch.empty_like(lo)
    timeit(f"maxpool_fwd C={C} S={S}", lambda: B.maxpool_fwd(skip, pooled, (2, 2, 2), sums), out_bytes + lo.numel() * 2)
    dz = torch.empty_like(flat)
    timeit(f"maxpool_bwd C={C} S={S}", lambda: B.maxpool_bwd(skip, lo, cat[..., :C], dz, (2, 2, 2), 1), 3 * out_bytes + lo.numel() * 2)
    coef = torch.randn((N, C, 3), device=dev)
    timeit(f"norm_bwd_apply C={C} S={S}", lambda: B.norm_bwd_apply(flat, skip, coef, None, dz, 1), 3 * out_bytes)
x1 = torch.randn((N, 128, 128, 128, 1), device=dev).bfloat16()
ss = torch.ones((N, 1, 2), device=dev)
timeit("im2col 1->32 taps S=128", lambda: B.im2col(x1, ss, (3, 3, 3), 32), N * 128 ** 3 * (2 + 64))
w1 = torch.randn((32, 1, 3, 3, 3), device=dev) * 0.2
pk1 = B.pack(("bench-first",), w1)
y1 = torch.empty((N, 128, 128, 128, 32), device=dev, dtype=torch.bfloat16)
s1 = torch.zeros((N, 32, 2), device=dev)
b1 = torch.zeros(32, device=dev)
timeit("first conv fwd 1->32 S=128", lambda: B.conv(x1, ss, pk1, b1, y1, s1, (3, 3, 3), True, False), N * 128 ** 3 * (2 + 64))
dw1 = torch.zeros_like(w1)
db1 = torch.zeros(32, device=dev)
timeit("first conv wgrad 1->32 S=128", lambda: B.wgrad(x1, ss, y1, dw1, db1, (3, 3, 3)), N * 128 ** 3 * (2 + 64))
