"""fp32 noise diagnosis (cfg1 model): error of the parameter gradients against the float64 oracle for (a) the fp32 CPU oracle,
(b) the same functional graph in fp32 on the GPU (cuDNN, TF32 off), (c) ours in fp32.  Prints max-abs error / tensor max."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch_em_b200 as tb  # noqa: E402
from oracle import dice as odice  # noqa: E402
from oracle import unet as ounet  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
DEV = "cuda:0"
torch.manual_seed(0)
net = tb.UNet3d(1, 2, depth=3, initial_features=16, final_activation="Sigmoid").to(DEV)
x = torch.randn(1, 1, 64, 64, 64)
t = (torch.rand(1, 2, 64, 64, 64) > 0.5).float()


def oracle(dtype, dev):
    sd = {k: v.detach().to(dev).to(dtype).clone().requires_grad_(True) for k, v in net.state_dict().items()}
    y_ = ounet.unet3d_forward(x.to(dev).to(dtype), sd, [2] * 3, final_activation="Sigmoid")
    l_ = odice.dice_loss(y_, t.to(dev).to(dtype))
    l_.backward()
    return y_.detach().cpu(), {k: v.grad.cpu() for k, v in sd.items()}


y64, g64 = oracle(torch.float64, "cpu")
y32, g32 = oracle(torch.float32, "cpu")
yg, gg = oracle(torch.float32, DEV)
y = net(x.to(DEV))
tb.DiceLoss()(y, t.to(DEV)).backward()
print("pred max err: cpu32 %.2e gpu32 %.2e ours %.2e" % (float((y32.double() - y64).abs().max()), float((yg.double() - y64).abs().max()),
                                                       float((y.detach().cpu().double() - y64).abs().max())))
for k, p in net.named_parameters():
    m = float(g64[k].abs().max())
    e = [float((g.double() - g64[k]).abs().max()) / m for g in (g32[k], gg[k], p.grad.cpu())]
    print(f"{k:36s} max {m:.2e}  cpu32 {e[0]:.1e}  gpu32 {e[1]:.1e}  ours {e[2]:.1e}")
