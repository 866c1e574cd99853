"""One launch of each hot kernel on its cfg2 L0 shape between cudaProfilerStart/Stop (for `ncu --set full`).

    ncu --set full --import-source on --clock-control none --profile-from-start off -o gpurun_out/r02_kernels python scripts/profile_kernels.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from torch_em_b200.backend import default_backend

dev = "cuda:0"
B = default_backend()
N, S = 4, 128
which = set((os.environ.get("KERNELS") or "ds_fwd,ds_dgrad,cs_wgrad,plain_fwd,plain_dgrad,plain_deep,head_mid_fwd,head_mid_bwd,h16_fwd,h16_dgrad,h16_wgrad,tf32_fwd,upsample_fwd,upsample_bwd,maxpool_bwd,"
                                          "norm_bwd_apply,cvt_f16,pack").split(","))
torch.manual_seed(0)
x = torch.randn((N, S, S, S, 32), device=dev).bfloat16()
dz = torch.randn((N, S, S, S, 32), device=dev).bfloat16()
w = torch.randn((32, 32, 3, 3, 3), device=dev) * 0.03
b = torch.zeros(32, device=dev)
ss = torch.ones((N, 32, 2), device=dev)
pk = B.pack(("prof", 32), w)
y = torch.empty_like(x)
g = torch.empty_like(x)
sums = torch.zeros((N, 32, 2), device=dev)
dw = torch.zeros_like(w)
db = torch.zeros(32, device=dev)
x2 = torch.randn((N, 32, 32, 32, 128), device=dev).bfloat16()
w2 = torch.randn((128, 128, 3, 3, 3), device=dev) * 0.03
pk2 = B.pack(("prof", 128), w2)
y2 = torch.empty_like(x2)
s2 = torch.zeros((N, 128, 2), device=dev)
ss2 = torch.ones((N, 128, 2), device=dev)
lo = torch.randn((N, S // 2, S // 2, S // 2, 32), device=dev).bfloat16()
cat = torch.empty((N, S, S, S, 64), device=dev, dtype=torch.bfloat16)
dlo = torch.empty_like(lo)
x1 = torch.randn((N, S, S, S, 1), device=dev).bfloat16()
ss1 = torch.ones((N, 1, 2), device=dev)
coef = torch.randn((N, 64, 3), device=dev)
# deep level (L4 of cfg2: 512 -> 512 on 8^3, split-K) and the fp32 paths (cfg4 L1: 128 -> 128 on 64^3, batch 1): h16 (default) / TF32
x4 = torch.randn((N, 8, 8, 8, 512), device=dev).bfloat16()
w4 = torch.randn((512, 512, 3, 3, 3), device=dev) * 0.01
pk4 = B.pack(("prof", 512), w4)
y4 = torch.empty_like(x4)
torch.backends.cudnn.allow_tf32 = True
xf = torch.randn((1, 64, 64, 64, 128), device=dev)
wf = torch.randn((128, 128, 3, 3, 3), device=dev) * 0.03
pkf = B.pack(("prof", "tf32"), wf)
from torch_em_b200.backend import CudaBackend
B32 = CudaBackend(use_h16=False)
yf = torch.empty_like(xf)
dzf = torch.randn((1, 64, 64, 64, 128), device=dev) * 1e-6
gf = torch.empty_like(xf)
dsf = torch.zeros((1, 128, 2), device=dev)
dwf = torch.zeros_like(wf)
dbf = torch.zeros(128, device=dev)
ps_big = B.pack_set({"big": torch.randn((512, 512, 3, 3, 3), device=dev) * 0.01})
ssf = torch.ones((1, 128, 2), device=dev)
sf = torch.zeros((1, 128, 2), device=dev)

dz2 = torch.randn_like(x2)
g2 = torch.empty_like(x2)
# cfg3's affinity head: 32 -> 12 channels on (2, 64, 256, 256)
xh = torch.randn((2, 64, 256, 256, 32), device=dev).bfloat16()
wh = torch.randn((12, 32, 1, 1, 1), device=dev) * 0.2
bh = torch.zeros(12, device=dev)
oh = torch.empty((2, 12, 64, 256, 256), device=dev)
gh = torch.randn_like(oh)
dxh = torch.empty_like(xh)
dwh = torch.zeros((12, 32), device=dev)
dbh = torch.zeros(12, device=dev)

runs = {
    "plain_dgrad": lambda: B.conv(dz2, None, pk2, None, g2, s2, (3, 3, 3), False, True, dot_x=x2),
    "head_mid_fwd": lambda: B.head_fwd(xh, wh, bh, oh, "Sigmoid"),
    "head_mid_bwd": lambda: B.head_bwd(gh, oh, xh, wh, dxh, dwh, dbh, "Sigmoid", True),
    "ds_fwd": lambda: B.conv(x, ss, pk, b, y, sums, (3, 3, 3), True, False),
    "ds_dgrad": lambda: B.conv(dz, None, pk, None, g, sums, (3, 3, 3), False, True, dot_x=x),
    "cs_wgrad": lambda: B.wgrad(x, ss, dz, dw, db, (3, 3, 3)),
    "plain_fwd": lambda: B.conv(x2, ss2, pk2, torch.zeros(128, device=dev), y2, s2, (3, 3, 3), True, False),
    "upsample_fwd": lambda: B.upsample_fwd(lo, cat[..., :32], (2, 2, 2), sums),
    "plain_deep": lambda: B.conv(x4, None, pk4, torch.zeros(512, device=dev), y4, None, (3, 3, 3), True, False),
    "h16_fwd": lambda: B.conv(xf, ssf, pkf, torch.zeros(128, device=dev), yf, sf, (3, 3, 3), True, False),
    "h16_dgrad": lambda: B.conv(dzf, None, pkf, None, gf, dsf, (3, 3, 3), False, True, dot_x=xf),
    "h16_wgrad": lambda: B.wgrad(xf, ssf, dzf, dwf, dbf, (3, 3, 3)),
    "tf32_fwd": lambda: B32.conv(xf, ssf, pkf, torch.zeros(128, device=dev), yf, sf, (3, 3, 3), True, False),
    "cvt_f16": lambda: B.to_h16(xf, ssf),
    "pack": lambda: ps_big.refresh_fwd(),
    "upsample_bwd": lambda: B.upsample_bwd(cat[..., :32], dlo, (2, 2, 2), zlow=lo, coef=coef[:, :32]),
    "maxpool_bwd": lambda: B.maxpool_bwd(cat[..., 32:], lo, cat[..., :32], y, (2, 2, 2), 1, coef=coef[:, 32:]),
    "norm_bwd_apply": lambda: B.norm_bwd_apply(g, x, coef[:, :32].contiguous(), None, y, 1),
}
for k, fn in runs.items():
    if k in which:
        fn(); fn()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for k, fn in runs.items():
    if k in which:
        fn()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
