"""One profiled train step of a bench workload (B200EM_CONFIG=cfg2|cfg3|cfg4, default cfg2) (use under ncu with --profile-from-start off).

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import torch_em_b200 as tb

cfg = bench.CONFIGS[os.environ.get("B200EM_CONFIG", "cfg2")]
batch = int(os.environ.get("B200EM_BATCH", cfg["batch"]))
patch = tuple(int(v) for v in os.environ["B200EM_PATCH"].split(",")) if "B200EM_PATCH" in os.environ else cfg["patch"]
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = getattr(tb, cfg["model"])(**cfg["model_kw"]).to(dev)
loss_fn = tb.AffinityLoss(bench.CREMI_OFFSETS, ignore_label=0) if cfg["loss"] == "affinity" else tb.DiceLoss()
boundary = tb.BoundaryTransform(add_binary_target=True) if cfg["loss"] == "boundary" else None
opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
x, t = bench.synthetic_batch(batch, patch, 1, cfg["loss"], cfg["model_kw"]["out_channels"])
x, t = x.to(dev), t.to(dev)
bf16 = cfg["dtype"] == "bf16"


def step():
    opt.zero_grad()
    tt = boundary(t) if boundary is not None else t
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=bf16):
        loss = loss_fn(model(x), tt)
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
