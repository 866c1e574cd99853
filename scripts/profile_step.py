"""One profiled train step of the bench workload (use under ncu with --profile-from-start off).

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import torch_em_b200 as tb

batch = int(os.environ.get("B200EM_BATCH", bench.BATCH))
patch = tuple(int(v) for v in os.environ.get("B200EM_PATCH", "128,128,128").split(","))
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = tb.UNet3d(**bench.MODEL_KW).to(dev)
loss_fn = tb.DiceLoss()
opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
x, t = bench.synthetic_batch(batch, patch, 1)
x, t = x.to(dev), t.to(dev)


def step():
    opt.zero_grad()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = loss_fn(model(x), t)
    loss.backward()
    opt.step()
    return loss


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
l = step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("loss", l.item())
