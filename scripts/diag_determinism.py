import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch_em_b200 as tb
DEV = "cuda:0"
torch.manual_seed(0)
size = int(os.environ.get("SIZE", 64))
net = tb.UNet3d(1, 2, depth=4, initial_features=32, final_activation="Sigmoid").to(DEV)
x = torch.randn(2, 1, size, size, size, device=DEV)
t = (torch.rand(2, 2, size, size, size, device=DEV) > 0.5).float()
for mode in ("bf16", "fp32"):
    outs = []
    for rep in range(2):
        net.zero_grad()
        if mode == "bf16":
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y = net(x); loss = tb.DiceLoss()(y, t)
        else:
            y = net(x); loss = tb.DiceLoss()(y, t)
        loss.backward()
        outs.append((y.detach().clone(), loss.item(), {k: p.grad.clone() for k, p in net.named_parameters()}))
    (y0, l0, g0), (y1, l1, g1) = outs
    print(mode, "loss", l0, l1, "pred maxdiff", float((y0 - y1).abs().max()), "rel", float((y0 - y1).norm() / y0.norm()))
    worst = sorted(((float((g0[k] - g1[k]).norm() / (g0[k].norm() + 1e-30)), k) for k in g0), reverse=True)[:6]
    for r, k in worst:
        print(f"   {k:40s} rel diff {r:.3e}  |g| {float(g0[k].norm()):.3e}")
