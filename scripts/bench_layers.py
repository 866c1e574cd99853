"""Per-layer timing of the conv kernels on the cfg2 shapes (forward, data gradient, weight gradient), CUDA events,
inputs larger than L2 for the big layers.  Diagnostic; prints one line per (layer, pass).

    python scripts/bench_layers.py [--only 32x32] [--reps 5] [--config cfg4]

--config cfg4: the layers of configs[3] (depth 5, f = 64, batch 1) with fp32 activations on the h16 path (B200EM_H16=0: TF32
kernels); the conv kernel alone is timed (backend.start_timing), the fp16 operand conversion passes are listed separately.
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


LAYERS_CFG4 = [
    ("L0 64->64", 64, 64, 128), ("L0 128->64", 128, 64, 128),
    ("L1 64->128", 64, 128, 64), ("L1 128->128", 128, 128, 64), ("L1 256->128", 256, 128, 64),
    ("L2 128->256", 128, 256, 32), ("L2 256->256", 256, 256, 32), ("L2 512->256", 512, 256, 32),
    ("L3 256->512", 256, 512, 16), ("L3 512->512", 512, 512, 16), ("L3 1024->512", 1024, 512, 16),
    ("L4 512->1024", 512, 1024, 8), ("L4 1024->1024", 1024, 1024, 8), ("L4 2048->1024", 2048, 1024, 8),
    ("L5 1024->2048", 1024, 2048, 4), ("L5 2048->2048", 2048, 2048, 4),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--size", type=int, default=0, help="override the spatial size of every layer")
    args = ap.parse_args()
    dev = "cuda:0"
    B = default_backend()
    f32 = args.config == "cfg4"
    N = args.batch or (1 if f32 else 4)
    adt = torch.float32 if f32 else torch.bfloat16
    torch.backends.cudnn.allow_tf32 = True
    tot = {"fwd": 0.0, "dgrad": 0.0, "wgrad": 0.0}
    for name, cin, cout, S in (LAYERS_CFG4 if f32 else LAYERS):
        if args.only and args.only not in name.replace("->", "x"):
            continue
        S = args.size or S
        torch.manual_seed(0)
        x = torch.randn((N, S, S, S, cin), device=dev).to(adt)
        dz = torch.randn((N, S, S, S, cout), device=dev).to(adt)
        w = torch.randn((cout, cin, 3, 3, 3), device=dev) * 0.03
        b = torch.zeros(cout, device=dev)
        ss = torch.ones((N, cin, 2), device=dev)
        pk = B.pack(("bench", name), w)
        y = torch.empty((N, S, S, S, cout), device=dev, dtype=adt)
        g = torch.empty((N, S, S, S, cin), device=dev, dtype=adt)
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
            B.start_timing()
            e0.record()
            for _ in range(args.reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            fam = B.stop_timing()
            ms_all = e0.elapsed_time(e1) / args.reps
            ms = sum(v[1] for v in fam.values()) / args.reps if f32 else ms_all       # fp32: the conv kernel alone
            tot[pname] += ms
            extra = f"   (+ {1e3 * (ms_all - ms):7.1f} us operand conversion; {'/'.join(sorted(fam))})" if f32 else ""
            print(f"{name:14s} {pname:6s} {ms * 1e3:9.1f} us  {fl / ms / 1e9:7.1f} TFLOP/s{extra}", flush=True)
    print("totals (ms):", {k: round(v, 3) for k, v in tot.items()})


if __name__ == "__main__":
    main()
