"""First-conv kernels (Cin=1 -> 32 on (4,128,128,128)) for ncu: direct forward and small-Cin weight gradient."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch_em_b200.backend import default_backend
dev = "cuda:0"; B = default_backend(); torch.manual_seed(0)
N, D, cin, cout = 4, 128, 1, 32
x = torch.randn((N, D, D, D, cin), device=dev).bfloat16()
w = torch.randn((cout, cin, 3, 3, 3), device=dev) * 0.1
b = torch.zeros(cout, device=dev); ss = torch.ones((N, cin, 2), device=dev)
pk = B.pack(("p1",), w)
y = torch.empty((N, D, D, D, cout), device=dev, dtype=torch.bfloat16)
sums = torch.zeros((N, cout, 2), device=dev); dw = torch.zeros_like(w); db = torch.zeros(cout, device=dev)
for it in range(2):
    B.conv(x, ss, pk, b, y, sums, (3, 3, 3), True, False)
    B.wgrad(x, ss, y, dw, db, (3, 3, 3))
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
e[0].record(); B.conv(x, ss, pk, b, y, sums, (3, 3, 3), True, False); e[1].record(); B.wgrad(x, ss, y, dw, db, (3, 3, 3)); e[2].record()
torch.cuda.synchronize()
print(f"fwd {e[0].elapsed_time(e[1]):.3f} ms  wgrad {e[1].elapsed_time(e[2]):.3f} ms")
