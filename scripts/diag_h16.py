"""Diagnostic: gradient error of the fp32 paths (h16 / tf32) against float64, next to the reference's own TF32 (cuDNN) error,
over a few seeds, on the depth-5 quarter-width topology of tests/test_gpu_tf32.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import torch_em_b200 as tb
from oracle import dice as odice
from oracle import unet as ounet
from torch_em_b200.backend import default_backend

DEV = "cuda:0"
torch.backends.cudnn.allow_tf32 = True
kw = dict(in_channels=1, out_channels=2, depth=int(os.environ.get("DEPTH", 5)), initial_features=16, final_activation="Sigmoid")
shape = (1, 1, 64, 64, 64)
B = default_backend()
for seed in range(3):
    torch.manual_seed(seed)
    net = tb.UNet3d(**kw).to(DEV)
    depth = kw["depth"]
    x = torch.randn(*shape)
    t = (torch.nn.functional.avg_pool3d(torch.randn(shape[0], 2, *shape[2:]), 5, 1, 2) > 0).float()

    def oracle(dtype, dev):
        sd = {k: v.detach().to(dev).to(dtype).clone().requires_grad_(True) for k, v in net.state_dict().items()}
        y_ = ounet.unet3d_forward(x.to(dev).to(dtype), sd, [2] * depth, final_activation="Sigmoid")
        l_ = odice.dice_loss(y_, t.to(dev).to(dtype))
        l_.backward()
        return y_.detach().cpu().double(), l_.item(), {k: v.grad.cpu().double() for k, v in sd.items()}

    y64, l64, g64 = oracle(torch.float64, "cpu")
    y_tf, l_tf, g_tf = oracle(torch.float32, DEV)
    gmax = max(float(v.norm()) for v in g64.values())
    res = {}
    for path in ("h16", "tf32"):
        B.use_h16 = path == "h16"
        net.zero_grad()
        y = net(x.to(DEV))
        loss = tb.DiceLoss()(y, t.to(DEV))
        loss.backward()
        torch.cuda.synchronize()
        ey = float((y.detach().cpu().double() - y64).norm() / y64.norm())
        eg = {k: float((p.grad.cpu().double() - g64[k]).norm()) / gmax for k, p in net.named_parameters()}
        res[path] = (ey, eg)
    ey_r = float((y_tf - y64).norm() / y64.norm())
    eg_r = {k: float((g_tf[k] - g64[k]).norm()) / gmax for k in g64}
    print(f"seed {seed}: pred rel err  h16 {res['h16'][0]:.2e}  tf32 {res['tf32'][0]:.2e}  ref {ey_r:.2e}")
    tot = {p: sum(v * v for v in res[p][1].values()) ** 0.5 for p in res}
    print(f"   total grad err / gmax: h16 {tot['h16']:.3e} tf32 {tot['tf32']:.3e} ref {sum(v * v for v in eg_r.values()) ** 0.5:.3e}")
    worst = sorted(eg_r, key=lambda k: -res['h16'][1][k])[:6]
    for k in worst:
        print(f"   {k:40s} h16 {res['h16'][1][k]:.3e} tf32 {res['tf32'][1][k]:.3e} ref {eg_r[k]:.3e}")
