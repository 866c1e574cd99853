"""Per-layer timing of the conv kernels on the cfg2 shapes (forward, data gradient, weight gradient), CUDA events,
inputs larger than L2 for the big layers.  Diagnostic; prints one line per (layer, pass).

    python scripts/bench_layers.py [--only 32x32] [--reps 5]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from torch_em_b200.backend import default_backend

LAYERS = [  # (name, Cin, Cout, spatial)
    ("L0 32->32", 32, 32, 128), ("L0 64->32", 64, 32, 128),
    ("L1 32->64", 32, 64, 64), ("L1 64->64", 64, 64, 64), ("L1 128->64", 128, 64, 64),
    ("L2 64->128", 64, 128, 32), ("L2 128->128", 128, 128, 32), ("L2 256->128", 256, 128, 32),
    ("L3 128->256", 128, 256, 16), ("L3 256->256", 256, 256, 16), ("L3 512->256", 512, 256, 16),
    ("L4 256->512", 256, 512, 8), ("L4 512->512", 512, 512, 8),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--size", type=int, default=0, help="override the spatial size of every layer")
    args = ap.parse_args()
    dev = "cuda:0"
    B = default_backend()
    N = args.batch
    tot = {"fwd": 0.0, "dgrad": 0.0, "wgrad": 0.0}
    for name, cin, cout, S in LAYERS:
        if args.only and args.only not in name.replace("->", "x"):
            continue
        S = args.size or S
        torch.manual_seed(0)
        x = torch.randn((N, S, S, S, cin), device=dev).bfloat16()
        dz = torch.randn((N, S, S, S, cout), device=dev).bfloat16()
        w = torch.randn((cout, cin, 3, 3, 3), device=dev) * 0.03
        b = torch.zeros(cout, device=dev)
        ss = torch.ones((N, cin, 2), device=dev)
        pk = B.pack(("bench", name), w)
        y = torch.empty((N, S, S, S, cout), device=dev, dtype=torch.bfloat16)
        g = torch.empty((N, S, S, S, cin), device=dev, dtype=torch.bfloat16)
        sums = torch.zeros((N, cout, 2), device=dev)
        dsums = torch.zeros((N, cin, 2), device=dev)
        dw = torch.zeros_like(w)
        db = torch.zeros(cout, device=dev)
        fl = 2.0 * N * S ** 3 * cin * cout * 27
        passes = {
            "fwd": lambda: B.conv(x, ss, pk, b, y, sums, (3, 3, 3), True, False),
            "dgrad": lambda: B.conv(dz, None, pk, None, g, dsums, (3, 3, 3), False, True, dot_x=x),
            "wgrad": lambda: B.wgrad(x, ss, dz, dw, db, (3, 3, 3)),
        }
        for pname, fn in passes.items():
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.reps
            tot[pname] += ms
            print(f"{name:12s} {pname:6s} {ms * 1e3:9.1f} us  {fl / ms / 1e9:7.1f} TFLOP/s", flush=True)
    print("totals (ms):", {k: round(v, 3) for k, v in tot.items()})


if __name__ == "__main__":
    main()
