"""diagnostic: absolute/relative error of the fused kernel's g and dQ against a float64 evaluation on the device"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rechorus_b200 import ops
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(3)
n, d, B, C = 1_000_000, 64, 4096, 100
U = torch.randn(n, d, device=dev, generator=g) * 0.2
I = torch.randn(n, d, device=dev, generator=g) * 0.2
uid = torch.randint(1, n, (B,), device=dev, generator=g)
iid = torch.randint(1, n, (B, C), device=dev, generator=g)
pred, gp, rl, dq = ops.bprmf_fused_fwd_bwd(U, uid, I, iid)
rows = I[iid].double()
q = U[uid].double()
x = torch.einsum("bd,bcd->bc", q, rows)
pos, neg = x[:, :1], x[:, 1:]
w = torch.softmax(neg - neg.max(), dim=1)
xg = x.detach().clone().requires_grad_(True)
S = ((xg[:, :1] - xg[:, 1:]).sigmoid() * torch.softmax(xg[:, 1:] - xg[:, 1:].max(), dim=1)).sum(1)
loss = -torch.log(S.clamp(1e-8, 1 - 1e-8)).mean()
loss.backward()
g64 = xg.grad
dq64 = torch.einsum("bc,bcd->bd", g64, rows)
for name, a, b in (("pred", pred, x), ("g", gp, g64), ("dQ", dq, dq64)):
    err = (a.double() - b).abs()
    print(f"{name}: max|ref| {float(b.abs().max()):.3e}  max err {float(err.max()):.3e}  rel-to-max {float(err.max() / b.abs().max()):.3e}  "
          f"median err {float(err.median()):.3e}")
# torch fp32 evaluation of the same thing for scale
dq32 = torch.einsum("bc,bcd->bd", g64.float(), I[iid])
print("torch fp32 einsum dQ err vs fp64:", float((dq32.double() - dq64).abs().max()))
