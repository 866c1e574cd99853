import sys, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch_em_b200 as tb
DEV="cuda:0"
torch.manual_seed(0)
net = tb.UNet3d(1, 2, depth=4, initial_features=32, final_activation="Sigmoid").to(DEV)
x = torch.randn(2, 1, 128, 128, 128, device=DEV)
for name, t in (("noise", (torch.rand(2, 2, 128, 128, 128, device=DEV) > 0.5).float()),
                ("smooth", (torch.nn.functional.avg_pool3d(x.repeat(1, 2, 1, 1, 1), 9, 1, 4) > 0).float())):
    gs, ls = [], []
    for scale in (1.0, 1.0, 4.0):
        net.zero_grad()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = tb.DiceLoss()(net(x * scale), t)
        loss.backward()
        ls.append(loss.item())
        gs.append(torch.cat([p.grad.flatten() for p in net.parameters()]).clone())
    c = lambda a, b: float(torch.dot(a, b) / (a.norm() * b.norm()))
    print(name, ls, "cos same", c(gs[0], gs[1]), "cos x4", c(gs[0], gs[2]), "rel", float((gs[0]-gs[1]).norm()/gs[0].norm()))
