"""time the causal-attention kernels at config-4 shape: register-resident (attention_rt.cu) vs the first kernels
(B2R_ATTN=v1), forward and backward, with and without live lengths"""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, torch
sys.path.insert(0, %r)
from rechorus_b200 import ops
B, L, d, H = 4096, 50, 64, 4
g = torch.Generator().manual_seed(1)
q, k, v = [torch.randn(B, L, d, generator=g).cuda().requires_grad_(True) for _ in range(3)]
go = torch.randn(B, L, d, generator=g).cuda()
for name, live in (("full", None), ("live~U[1,L]", torch.randint(1, L + 1, (B,), generator=g).cuda())):
    def fwd():
        return ops.causal_attention(q, k, v, H, live=live)
    for _ in range(3):
        o = fwd(); o.backward(go)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = tb = 0.0
    n = 20
    for _ in range(n):
        e[0].record(); o = fwd(); e[1].record(); o.backward(go); e[2].record()
        torch.cuda.synchronize()
        tf += e[0].elapsed_time(e[1]); tb += e[1].elapsed_time(e[2])
    print("%%-12s fwd %%.1f us  bwd(+autograd) %%.1f us" %% (name, tf / n * 1e3, tb / n * 1e3))
''' % ROOT
for env in ({"B2R_ATTN": "rt"}, {"B2R_ATTN": "v1"}):
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True, timeout=600)
    print(env, "\n" + r.stdout.strip(), r.stderr.strip()[-600:])
