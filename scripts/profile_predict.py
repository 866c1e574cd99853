"""One profiled tiled prediction (cfg5 blocks) for an ncu launch list:

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_predict.csv \
        python scripts/profile_predict.py
Predicts a (256, 256, 256) volume with 128^3 blocks + halo 32 (8 haloed blocks of 192^3 = 2 forward passes of 4 blocks)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import torch_em_b200 as tb
from torch_em_b200.util import predict_with_halo

cfg = bench.CONFIGS["cfg5"]
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = tb.UNet3d(**cfg["model_kw"]).to(dev).eval()
vol = bench.cfg5_volume((256, 256, 256))


def run():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        return predict_with_halo(vol, model, [0], cfg["block_shape"], cfg["halo"])


run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
out = run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("out", out.shape, float(out.mean()))
