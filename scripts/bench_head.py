"""Times the network-head kernels (out_conv 1x1x1 + activation, forward and backward) on the cfg3 / cfg2 head shapes.

    python scripts/bench_head.py            # (2, 64, 256, 256) x 32 -> 12 (cfg3) and (4, 128^3) x 32 -> 2 (cfg2), bf16
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from torch_em_b200.backend import default_backend

dev = "cuda:0"
B = default_backend()
torch.manual_seed(0)
shapes = [(2, 64, 256, 256, 32, 12), (4, 128, 128, 128, 32, 2)]
if os.environ.get("ODD"):      # plane strides that are not a power of two (partition-camping check)
    shapes += [(2, 63, 254, 250, 32, 12), (2, 64, 256, 256, 32, 8), (2, 64, 256, 256, 32, 4)]
for (N, D, H, W, Cin, Cout) in shapes:
    x = torch.randn((N, D, H, W, Cin), device=dev).bfloat16()
    w = torch.randn((Cout, Cin, 1, 1, 1), device=dev) * 0.2
    b = torch.randn(Cout, device=dev) * 0.1
    out = torch.empty((N, Cout, D, H, W), device=dev)
    g = torch.randn_like(out)
    dx = torch.empty_like(x)
    dw = torch.zeros((Cout, Cin), device=dev)
    db = torch.zeros(Cout, device=dev)

    def timed(fn, n=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    tf = timed(lambda: B.head_fwd(x, w, b, out, "Sigmoid"))
    tb_ = timed(lambda: B.head_bwd(g, out, x, w, dx, dw, db, "Sigmoid", True))
    vox = N * D * H * W
    bytes_f = vox * (Cin * 2 + Cout * 4)
    bytes_b = vox * (2 * Cin * 2 + 2 * Cout * 4)
    # parity of one call against plain torch (fp32 math on the bf16 input)
    dw.zero_(); db.zero_()
    B.head_fwd(x, w, b, out, "Sigmoid")
    B.head_bwd(g, out, x, w, dx, dw, db, "Sigmoid", True)
    xs = x[:, :8].float()
    ref = torch.sigmoid(torch.einsum("ndhwc,oc->nodhw", xs, w.reshape(Cout, Cin)) + b.view(1, -1, 1, 1, 1))
    dz = g[:, :, :8] * ref * (1 - ref)
    rdx = torch.einsum("nodhw,oc->ndhwc", dz, w.reshape(Cout, Cin)) * (xs > 0)
    dzf = g * out * (1 - out)
    rdw = torch.einsum("nodhw,ndhwc->oc", dzf, x.float())
    err_o = (out[:, :, :8] - ref).abs().max().item()
    err_dx = ((dx[:, :8].float() - rdx).abs().max() / rdx.abs().max()).item()
    err_dw = ((dw - rdw).abs().max() / rdw.abs().max()).item()
    err_db = ((db - dzf.sum((0, 2, 3, 4))).abs().max() / dzf.sum((0, 2, 3, 4)).abs().max()).item()
    print(f"head {Cin}->{Cout} on {(N, D, H, W)}: fwd {tf * 1e3:.0f} us ({bytes_f / tf / 1e9:.2f} TB/s)  bwd {tb_ * 1e3:.0f} us ({bytes_b / tb_ / 1e9:.2f} TB/s)"
          f"   max err out {err_o:.1e} dx {err_dx:.1e} dw {err_dw:.1e} db {err_db:.1e}")
