"""Run single L0-shaped kernels of the bench workload (for `ncu --set full`): conv fwd 32->32 and wgrad 32x32 on (4,128,128,128)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from torch_em_b200.backend import default_backend

cin = int(os.environ.get("CIN", 32)); cout = int(os.environ.get("COUT", 32))
N, D = int(os.environ.get("NB", 4)), int(os.environ.get("DD", 128))
dev = "cuda:0"
B = default_backend()
torch.manual_seed(0)
x = torch.randn((N, D, D, D, cin), device=dev).bfloat16()
w = torch.randn((cout, cin, 3, 3, 3), device=dev) * 0.03
b = torch.zeros(cout, device=dev)
ss = torch.ones((N, cin, 2), device=dev)
pk = B.pack(("p", cin, cout), w)
y = torch.empty((N, D, D, D, cout), device=dev, dtype=torch.bfloat16)
sums = torch.zeros((N, cout, 2), device=dev)
dw = torch.zeros_like(w); db = torch.zeros(cout, device=dev)
for it in range(3):
    B.conv(x, ss, pk, b, y, sums, (3, 3, 3), True, False)
    B.wgrad(x, ss, y, dw, db, (3, 3, 3))
torch.cuda.synchronize()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record(); B.conv(x, ss, pk, b, y, sums, (3, 3, 3), True, False); e1.record(); B.wgrad(x, ss, y, dw, db, (3, 3, 3)); e2.record()
torch.cuda.synchronize()
fl = 2.0 * N * D ** 3 * cin * cout * 27
print(f"fwd {e0.elapsed_time(e1):.3f} ms {fl / e0.elapsed_time(e1) / 1e9:.1f} TFLOP/s   wgrad {e1.elapsed_time(e2):.3f} ms {fl / e1.elapsed_time(e2) / 1e9:.1f} TFLOP/s")
