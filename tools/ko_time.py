"""diagnostic: which stage pins the pipelined tcgen05 Linear kernels?  B2R_TC_KO knocks stages out (1 MMAs, 2 epilogue
stores, 4 split, 8 global loads); results are wrong by design, only the times matter."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, torch
sys.path.insert(0, %r)
from rechorus_b200 import ops
M, N, K = 204800, 64, 64
g = torch.Generator().manual_seed(1)
x = torch.randn(M, K, generator=g).cuda(); W = (torch.randn(N, K, generator=g) * 0.1).cuda(); b = torch.randn(N, generator=g).cuda()
dy = torch.randn(M, N, generator=g).cuda()
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
print("fwd %%.1f us   dW %%.1f us" %% (t(lambda: ops.linear_fwd_tc(x, W, b, True)), t(lambda: ops.linear_bwd(dy, x, W, None, False, True, True))))
''' % ROOT
for ko in (0, 1, 2, 4, 8, 5, 12, 13, 15):
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, B2R_TC_KO=str(ko)), capture_output=True, text=True, timeout=300)
    print("KO=%2d (%s)" % (ko, ",".join(n for bit, n in ((1, "noMMA"), (2, "noStore"), (4, "noSplit"), (8, "noLoad")) if ko & bit) or "full"),
          r.stdout.strip(), r.stderr.strip()[-300:])
