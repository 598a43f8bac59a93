"""one call of each dense hot kernel at config-4 shape, for ncu"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rechorus_b200 import ops
M, N, K = 204800, 64, 64
g = torch.Generator().manual_seed(1)
x = torch.randn(M, K, generator=g).cuda(); W = (torch.randn(N, K, generator=g) * 0.1).cuda(); b = torch.randn(N, generator=g).cuda()
dy = torch.randn(M, N, generator=g).cuda()
for _ in range(2):
    ops.linear_fwd_tc(x, W, b, True)
    ops.linear_fwd(x, W, b, True)
    ops.linear_bwd(dy, x, W, None, False, True, True)
B, L, d, H = 4096, 50, 64, 4
q, k, v = [torch.randn(B, L, d, generator=g).cuda().requires_grad_(True) for _ in range(3)]
go = torch.randn(B, L, d, generator=g).cuda()
for _ in range(2):
    o = ops.causal_attention(q, k, v, H); o.backward(go)
torch.cuda.synchronize()
