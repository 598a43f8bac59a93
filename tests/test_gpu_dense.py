"""GPU: the dense building blocks (SGEMM Linear, residual LayerNorm, causal attention, small-table gradient,
strided gather) against plain PyTorch fp32 on CPU."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _close(a, b, tol=TOL):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    scale = max(1.0, float(b.abs().max()))
    assert float((a - b).abs().max()) <= tol * scale, float((a - b).abs().max())


@pytest.mark.parametrize("M,K,N,relu,bias", [(7, 64, 64, True, True), (300, 128, 64, True, True), (129, 64, 32, False, True),
                                             (65, 32, 16, True, True), (50, 80, 1, False, False), (33, 50, 10, True, True),
                                             (4100, 64, 192, False, True)])
def test_linear_forward_backward(M, K, N, relu, bias):
    from rechorus_b200 import ops
    g = torch.Generator().manual_seed(M + K + N)
    x = torch.randn(M, K, generator=g).requires_grad_(True)
    W = (torch.randn(N, K, generator=g) * 0.2).requires_grad_(True)
    b = (torch.randn(N, generator=g) * 0.2).requires_grad_(True) if bias else None
    y = F.linear(x, W, b)
    if relu:
        y = torch.relu(y)
    gy = torch.randn(M, N, generator=g)
    y.backward(gy)
    xc = x.detach().cuda().requires_grad_(True)
    Wc = W.detach().cuda().requires_grad_(True)
    bc = b.detach().cuda().requires_grad_(True) if bias else None
    yc = ops.linear(xc, Wc, bc, relu=relu)
    yc.backward(gy.cuda())
    _close(yc, y)
    _close(xc.grad, x.grad)
    _close(Wc.grad, W.grad, 2e-5)
    if bias:
        _close(bc.grad, b.grad, 2e-5)


def test_linear_on_strided_input_columns():
    from rechorus_b200 import ops
    g = torch.Generator().manual_seed(1)
    big = torch.randn(40, 128, generator=g)
    W = torch.randn(16, 64, generator=g) * 0.1
    y = ops.linear_fwd(big.cuda()[:, 64:], W.cuda(), None, False)        # row stride 128, K = 64
    _close(y, F.linear(big[:, 64:], W))


@pytest.mark.parametrize("rows,d", [(5, 64), (1000, 64), (77, 32), (300, 128), (4099, 64), (10, 48)])
def test_add_layernorm_forward_backward(rows, d):
    from rechorus_b200 import ops
    g = torch.Generator().manual_seed(rows + d)
    x = torch.randn(rows, d, generator=g).requires_grad_(True)
    r = torch.randn(rows, d, generator=g).requires_grad_(True)
    gam = (1 + 0.1 * torch.randn(d, generator=g)).requires_grad_(True)
    bet = (0.1 * torch.randn(d, generator=g)).requires_grad_(True)
    y = F.layer_norm(x + r, (d,), gam, bet, 1e-5)
    gy = torch.randn(rows, d, generator=g)
    y.backward(gy)
    cu = [t.detach().cuda().requires_grad_(True) for t in (x, r, gam, bet)]
    yc = ops.add_layernorm(cu[0].view(1, rows, d), cu[1].view(1, rows, d), cu[2], cu[3])
    yc.backward(gy.cuda().view(1, rows, d))
    _close(yc.view(rows, d), y)
    _close(cu[0].grad, x.grad)
    _close(cu[1].grad, r.grad)
    _close(cu[2].grad, gam.grad, 2e-5)
    _close(cu[3].grad, bet.grad, 2e-5)


def _ref_attention(q, k, v, H):
    B, L, d = q.shape
    dk = d // H
    def split(t):
        return t.view(B, L, H, dk).permute(0, 2, 1, 3)
    s = split(q) @ split(k).transpose(-1, -2) / dk ** 0.5
    s = s.masked_fill(~torch.ones(L, L, dtype=torch.bool).tril(), float("-inf"))
    a = torch.softmax(s - s.max(), dim=-1)
    return (a @ split(v)).permute(0, 2, 1, 3).reshape(B, L, d)


@pytest.mark.parametrize("B,L,d,H", [(3, 50, 64, 4), (2, 20, 64, 1), (5, 12, 32, 2), (1, 1, 64, 4), (2, 70, 64, 4),
                                     (2, 33, 128, 8), (3, 50, 64, 2), (2, 128, 32, 4), (2, 31, 64, 8), (2, 32, 48, 3)])
def test_causal_attention_forward_backward(B, L, d, H):
    from rechorus_b200 import ops
    g = torch.Generator().manual_seed(B * 100 + L)
    q, k, v = [(torch.randn(B, L, d, generator=g)).requires_grad_(True) for _ in range(3)]
    out = _ref_attention(q, k, v, H)
    go = torch.randn(B, L, d, generator=g)
    out.backward(go)
    qc, kc, vc = [t.detach().cuda().requires_grad_(True) for t in (q, k, v)]
    oc = ops.causal_attention(qc, kc, vc, H)
    oc.backward(go.cuda())
    _close(oc, out)
    _close(qc.grad, q.grad, 2e-5)
    _close(kc.grad, k.grad, 2e-5)
    _close(vc.grad, v.grad, 2e-5)


@pytest.mark.parametrize("B,L,d,H", [(6, 50, 64, 4), (4, 33, 128, 8), (3, 7, 32, 2), (5, 100, 64, 2), (4, 50, 64, 1)])
def test_causal_attention_with_dead_rows_skipped_equals_full_attention_on_the_live_rows(B, L, d, H):
    """live[b] positions matter, the rest is dead work (SASRec reads position len-1 only and the mask is causal): the
    live rows of the output and of dq/dk/dv equal the full computation's when the dead rows receive no upstream gradient;
    dead rows come back as zeros"""
    from rechorus_b200 import ops
    g = torch.Generator().manual_seed(B + L)
    q, k, v = [(torch.randn(B, L, d, generator=g)).requires_grad_(True) for _ in range(3)]
    live = torch.randint(1, L + 1, (B,), generator=g)
    live[0], live[-1] = L, 1
    rowmask = (torch.arange(L).view(1, L) < live.view(B, 1)).unsqueeze(-1).float()
    out = _ref_attention(q, k, v, H)
    go = torch.randn(B, L, d, generator=g) * rowmask
    out.backward(go)
    qc, kc, vc = [t.detach().cuda().requires_grad_(True) for t in (q, k, v)]
    oc = ops.causal_attention(qc, kc, vc, H, live=live.cuda())
    oc.backward(go.cuda())
    _close(oc, out * rowmask)
    assert float((oc.cpu() * (1 - rowmask)).abs().max()) == 0.0
    for a, b in ((qc.grad, q.grad), (kc.grad, k.grad), (vc.grad, v.grad)):
        _close(a, b * rowmask, 2e-5)
        assert float((a.cpu() * (1 - rowmask)).abs().max()) == 0.0


def test_embed_history_and_small_table_gradient():
    from rechorus_b200 import ops
    g = torch.Generator().manual_seed(5)
    n_items, L, d, B = 40, 9, 64, 300
    I = (torch.randn(n_items, d, generator=g)).requires_grad_(True)
    P = (torch.randn(L + 1, d, generator=g)).requires_grad_(True)
    lengths = torch.randint(1, L + 1, (B,), generator=g)
    lengths[0] = L
    hist = torch.randint(1, n_items, (B, L), generator=g) * (torch.arange(L).view(1, L) < lengths.view(B, 1))
    valid = (hist > 0).long()
    pos = (lengths.view(B, 1) - torch.arange(L).view(1, L)) * valid
    x = F.embedding(hist, I) + F.embedding(pos, P)
    gx = torch.randn(B, L, d, generator=g) * valid.unsqueeze(-1)      # padded positions carry zero gradient
    x.backward(gx)
    Ic, Pc = I.detach().cuda().requires_grad_(True), P.detach().cuda().requires_grad_(True)
    xc = ops.embed_history(Ic, Pc, hist.cuda(), lengths.cuda())
    xc.backward(gx.cuda())
    _close(xc, x)
    _close(Ic.grad, I.grad)
    _close(Pc.grad, P.grad, 3e-5)
    # select_last
    y = torch.randn(B, L, d, generator=g).requires_grad_(True)
    h = (y * valid.unsqueeze(-1).float())[torch.arange(B), lengths - 1]
    gh = torch.randn(B, d, generator=g)
    h.backward(gh)
    yc = y.detach().cuda().requires_grad_(True)
    hc = ops.select_last(yc, hist.cuda(), lengths.cuda())
    hc.backward(gh.cuda())
    assert torch.equal(hc.cpu(), h.detach()) and torch.equal(yc.grad.cpu(), y.grad)
    ops.check_ids()


def test_gather_concat_and_colscale():
    from rechorus_b200 import ops
    g = torch.Generator().manual_seed(6)
    Tu = torch.randn(30, 64, generator=g).requires_grad_(True)
    Ti = torch.randn(50, 64, generator=g).requires_grad_(True)
    uid = torch.randint(0, 30, (11,), generator=g)
    iid = torch.randint(0, 50, (11, 5), generator=g)
    x = torch.cat([F.embedding(uid.view(11, 1).expand(11, 5), Tu), F.embedding(iid, Ti)], dim=-1).view(55, 128)
    gx = torch.randn(55, 128, generator=g)
    x.backward(gx)
    Tuc, Tic = Tu.detach().cuda().requires_grad_(True), Ti.detach().cuda().requires_grad_(True)
    xc = ops.gather_concat(Tuc, Tic, uid.cuda(), iid.cuda())
    xc.backward(gx.cuda())
    assert torch.equal(xc.cpu(), x.detach())
    _close(Tuc.grad, Tu.grad)
    _close(Tic.grad, Ti.grad)
    a = torch.randn(20, 64, generator=g).requires_grad_(True)
    w = torch.randn(64, generator=g).requires_grad_(True)
    (a * w).backward(gx[:20, :64])
    ac, wc = a.detach().cuda().requires_grad_(True), w.detach().cuda().requires_grad_(True)
    ops.colscale(ac, wc).backward(gx[:20, :64].cuda())
    _close(ac.grad, a.grad)
    _close(wc.grad, w.grad)
