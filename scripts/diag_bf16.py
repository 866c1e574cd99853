"""Diagnostic: per-parameter gradient cosine vs the fp32 oracle for (a) our bf16 path, (b) the oracle under bf16 autocast on the GPU."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch_em_b200 as tb
from oracle import dice as odice, unet as ounet

DEV = "cuda:0"
norm = sys.argv[1] if len(sys.argv) > 1 else "InstanceNorm"
torch.manual_seed(0)
kw = dict(in_channels=1, out_channels=2, depth=3, initial_features=16, final_activation="Sigmoid", norm=norm)
net = tb.UNet3d(**kw).to(DEV)
x = torch.randn(2, 1, 32, 32, 32)
t = (torch.rand(2, 2, 32, 32, 32) > 0.5).float()
sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in net.state_dict().items()}
y_ref = ounet.unet3d_forward(x, sd, [2, 2, 2], norm=norm, final_activation="Sigmoid")
odice.dice_loss(y_ref, t).backward()
sdg = {k: v.detach().to(DEV).clone().requires_grad_(True) for k, v in net.state_dict().items()}
with torch.autocast("cuda", dtype=torch.bfloat16):
    y_ac = ounet.unet3d_forward(x.to(DEV), sdg, [2, 2, 2], norm=norm, final_activation="Sigmoid")
    l_ac = odice.dice_loss(y_ac, t.to(DEV))
l_ac.backward()
with torch.autocast("cuda", dtype=torch.bfloat16):
    y = net(x.to(DEV))
    loss = tb.DiceLoss()(y, t.to(DEV))
loss.backward()
rel = lambda a, b: float((a.float().cpu() - b).norm() / b.norm())
print("pred rel-L2: ours", rel(y.detach(), y_ref.detach()), "autocast-ref", rel(y_ac.detach(), y_ref.detach()))
for k, p in net.named_parameters():
    b = sd[k].grad.flatten().double()
    cs = []
    for g in (p.grad, sdg[k].grad):
        a = g.cpu().flatten().double()
        cs.append(float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-30)))
    print(f"{k:45s} ours {cs[0]:.4f}  autocast-ref {cs[1]:.4f}")
