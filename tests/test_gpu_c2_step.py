"""Parity of the BENCHMARKED route at BASELINE config 2, full size (1 M x 1 M rows, d=64, B=4096, K=99):
``model.train_step(feed, next_feed)`` = multi-pass fused gather/score/loss/dQ kernel -> prefetched bucket plan ->
fused row-sparse Adam on both tables (what bench.py times), against the CPU oracle.

The oracle cannot hold 1 M-row autograd tables cheaply, so the touched rows are compacted: ids are remapped onto the
sorted unique rows of the batch and the oracle (the reference's arithmetic, oracle/rechorus_oracle.py) differentiates
the compact tables -- every contribution, duplicate and accumulation of the real step is in there.  Checked:
  (i)   the fused kernel at B=4096 (passes > resident CTAs: the cross-pass cp.async ring, ids prefetched two passes
        ahead) -- pred / g / dQ / loss of every sample vs fp32 torch on the device and a 64-sample slice vs the oracle;
  (ii)  every touched row's w, m, v after each of 3 prefetched steps vs the lazy-Adam formula on oracle gradients;
  (iii) untouched rows bit-unchanged;  (iv) two runs give the same bits.
Tolerances (north_star): scores / grads 1e-5 fp32 absolute; weights after Adam 1e-6 absolute (update size 1e-3).
"""
import argparse
import types

import numpy as np
import pytest
import torch

from oracle import rechorus_oracle as O

pytestmark = pytest.mark.gpu

N_USERS = N_ITEMS = 1_000_000
D, B, C = 64, 4096, 100
LR, B1, B2, EPS = 1e-3, 0.9, 0.999, 1e-8


def _dev():
    return torch.device("cuda", 0)


def _model(seed=0, scale=20.0):
    from rechorus_b200 import plugin
    from rechorus_b200.optim import RowSparseOptimizer
    from rechorus_b200.runner import BaseRunner
    p = argparse.ArgumentParser()
    p = BaseRunner.parse_runner_args(p)
    p = plugin.BPRMF.parse_model_args(p)
    a = p.parse_args(["--emb_size", str(D), "--num_neg", str(C - 1), "--table_mode", "fused"])
    a.device, a.model_path = _dev(), "/tmp/_b2r_unused.pt"
    torch.manual_seed(seed)
    m = plugin.BPRMF(a, types.SimpleNamespace(n_users=N_USERS, n_items=N_ITEMS)).to(_dev())
    with torch.no_grad():
        for q in m.parameters():
            q.mul_(scale)                       # trained-scale weights: softmax weights far from uniform
    m.optimizer = RowSparseOptimizer(m, "Adam", lr=LR, l2=0.0)
    m.train()
    return m


def _feeds(n, seed=5, zipf=False):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n):
        uid = torch.randint(1, N_USERS, (B,), generator=g)
        iid = torch.randint(1, N_ITEMS, (B, C), generator=g)
        if zipf:                                # heavy duplication in the positive column (SURVEY 8d variant)
            iid[:, 0] = torch.from_numpy(np.random.default_rng(seed).zipf(1.05, B) % (N_ITEMS - 1) + 1)
            uid[::16] = uid[0]                  # duplicated users too
        out.append({"user_id": uid.to(_dev()), "item_id": iid.to(_dev()), "batch_size": B, "phase": "train"})
    return out


def _compact_oracle_grads(U_rows, I_rows, uu, ui, uid, iid):
    """oracle loss + gradients on the compacted tables (rows = sorted unique ids of this batch)"""
    batch = {"user_id": torch.searchsorted(uu, uid), "item_id": torch.searchsorted(ui, iid)}
    w = {"u_embeddings.weight": U_rows, "i_embeddings.weight": I_rows}
    pred, loss, g = O.loss_and_grads("BPRMF", w, batch)
    return pred, float(loss), g["u_embeddings.weight"], g["i_embeddings.weight"]


def _lazy_adam(w, m, v, g, t):
    """torch.optim.Adam's formula (helpers/BaseRunner.py:110-114 builds it) for the touched rows, in fp64"""
    w, m, v, g = w.double(), m.double(), v.double(), g.double()
    m = B1 * m + (1 - B1) * g
    v = B2 * v + (1 - B2) * g * g
    w = w - (LR / (1 - B1 ** t)) * m / (v.sqrt() / (1 - B2 ** t) ** 0.5 + EPS)
    return w, m, v


def test_fused_kernel_full_size_multi_pass_matches_torch_and_oracle():
    from rechorus_b200 import ops
    m = _model()
    U, I = m.u_embeddings.weight.data, m.i_embeddings.weight.data
    for zipf in (False, True):
        f = _feeds(1, seed=11, zipf=zipf)[0]
        uid, iid = f["user_id"], f["item_id"]
        pred, gp, row_loss, dq = ops.bprmf_fused_fwd_bwd(U, uid, I, iid)
        rows = I[iid]                                                    # [B, C, d] fp32 on the device
        ref = torch.einsum("bd,bcd->bc", U[uid], rows)
        assert (pred - ref).abs().max() <= 1e-5
        l64, g64 = O.bpr_loss_and_grad_fp64(ref.cpu().numpy())
        assert np.abs(gp.cpu().numpy() - g64).max() <= 1e-5
        assert abs(float(row_loss.double().mean()) - l64) <= 1e-5
        dq_ref = torch.einsum("bc,bcd->bd", torch.from_numpy(g64).to(_dev()).float(), rows)
        assert (dq - dq_ref).abs().max() <= 1e-5
        # relative form (g is ~1/B): the gradient must be reproduced, not merely small
        assert np.abs(gp.cpu().numpy() - g64).max() <= 2e-5 * np.abs(g64).max()
        # a 64-sample slice through the oracle proper (mean over the slice -> scale by 64 / B)
        sl = torch.arange(0, B, B // 64)
        uu, ui = torch.unique(uid[sl].cpu()), torch.unique(iid[sl].cpu())
        p_o, _, gU, gI = _compact_oracle_grads(U[uu.to(_dev())].cpu(), I[ui.to(_dev())].cpu(), uu, ui, uid[sl].cpu(),
                                               iid[sl].cpu())
        assert (pred[sl].cpu() - p_o).abs().max() <= 1e-5
        # dU of the slice's users == dq rows (unique users in the slice) up to the batch-mean factor
        if not zipf:
            pos = torch.searchsorted(uu, uid[sl].cpu())
            assert (dq[sl].cpu() * (B / 64.0) - gU[pos]).abs().max() <= 1e-5
        ops.check_ids()


@pytest.mark.parametrize("zipf", [False, True])
def test_train_step_full_size_three_prefetched_steps_match_lazy_adam_on_oracle_grads(zipf):
    from rechorus_b200 import ops
    n_steps = 3
    feeds = _feeds(n_steps, seed=21, zipf=zipf)
    m = _model()
    U, I = m.u_embeddings.weight.data, m.i_embeddings.weight.data
    eu, ei = m.optimizer.entry(m.u_embeddings.weight), m.optimizer.entry(m.i_embeddings.weight)
    U0, I0 = U.clone(), I.clone()
    touched_u = torch.zeros(N_USERS, dtype=torch.bool, device=_dev())
    touched_i = torch.zeros(N_ITEMS, dtype=torch.bool, device=_dev())
    for t in range(1, n_steps + 1):
        f = feeds[t - 1]
        uid, iid = f["user_id"].cpu(), f["item_id"].cpu()
        uu, ui = torch.unique(uid), torch.unique(iid)
        uud, uid_d = uu.to(_dev()), ui.to(_dev())
        before = [x[idx].cpu() for x, idx in ((U, uud), (eu["m"], uud), (eu["v"], uud), (I, uid_d), (ei["m"], uid_d),
                                              (ei["v"], uid_d))]
        loss = m.train_step(f, feeds[t] if t < n_steps else None)
        torch.cuda.synchronize()
        _, loss_o, gU, gI = _compact_oracle_grads(before[0], before[3], uu, ui, uid, iid)
        assert abs(float(loss) - loss_o) <= 1e-5
        for name, (w0, m0, v0), g, W, e, idx in (("U", before[0:3], gU, U, eu, uud), ("I", before[3:6], gI, I, ei, uid_d)):
            w1, m1, v1 = _lazy_adam(w0, m0, v0, g, t)
            err = (W[idx].cpu().double() - w1).abs()
            well = g.abs() > 1e-7                 # Adam divides by |g| + 1e-8: entries with |g| near eps amplify rounding
            assert float(err[well].max()) <= 1e-6, (name, t, float(err[well].max()))
            assert float(err.max()) <= 5e-5, (name, t, float(err.max()))     # ill-conditioned entries: a fraction of lr
            assert (e["m"][idx].cpu().double() - m1).abs().max() <= 1e-5 * float(g.abs().max()) + 1e-12, (name, t)
            assert (e["v"][idx].cpu().double() - v1).abs().max() <= 1e-4 * float(v1.abs().max()) + 1e-16, (name, t)
            # the step must be reproduced, not lost: the update itself to 1 % of its own size (lr) on well-conditioned
            # entries (|g| >> eps)
            upd = (W[idx].cpu().double() - w0.double())
            ok = g.abs() > 1e-6
            assert ((upd - (w1 - w0.double())).abs()[ok]).max() <= 1e-2 * LR, (name, t)
        touched_u[uud] = True
        touched_i[uid_d] = True
    # (iii) untouched rows: bit-identical weights, zero moments
    assert torch.equal(U[~touched_u], U0[~touched_u]) and torch.equal(I[~touched_i], I0[~touched_i])
    assert not bool(eu["m"][~touched_u].any()) and not bool(ei["v"][~touched_i].any())
    assert bool((U[touched_u] != U0[touched_u]).any(dim=1).all())           # every touched row moved
    assert bool((I[touched_i] != I0[touched_i]).any(dim=1).all())
    ops.check_ids()
    # (iv) bit-reproducibility of the whole 3-step run (fresh model, same seed, same feeds)
    m2 = _model()
    for t in range(1, n_steps + 1):
        m2.train_step(feeds[t - 1], feeds[t] if t < n_steps else None)
    torch.cuda.synchronize()
    assert torch.equal(m2.u_embeddings.weight.data, U) and torch.equal(m2.i_embeddings.weight.data, I)
    e2 = m2.optimizer.entry(m2.i_embeddings.weight)
    assert torch.equal(e2["m"], ei["m"]) and torch.equal(e2["v"], ei["v"])


def test_contract_route_equals_train_step_at_full_size():
    """forward -> loss -> backward -> RowSparseOptimizer.step() (what the reference's unchanged BaseRunner.fit drives,
    BaseRunner.py:193-206) and the single-call train_step leave the same tables (different kernels, same sums)."""
    feeds = _feeds(2, seed=31)
    ma, mb = _model(), _model()
    for f in feeds:
        ma.optimizer.zero_grad()
        la = ma.loss(ma(f))
        la.backward()
        ma.optimizer.step()
        lb = mb.train_step(f)
        assert abs(float(la) - float(lb)) <= 2e-6
    for (k, pa), (_, pb) in zip(ma.named_parameters(), mb.named_parameters()):
        assert (pa - pb).abs().max() <= 1e-6, k


@pytest.mark.parametrize("kind", ["one_row", "two_hot_buckets"])  # "one_row": really two (see below)
def test_train_step_degenerate_id_distributions_match_oracle(kind):
    """skew the fixed-capacity bucket regions of the step's index plan cannot hold: every pair on ONE row (one bucket
    receives the whole batch through the spill list and is merge-sorted in chunks) / two hot buckets plus uniform rest"""
    from rechorus_b200 import ops
    Bs = 1024
    g = torch.Generator().manual_seed(77)
    uid = torch.randint(1, N_USERS, (Bs,), generator=g)
    iid = torch.randint(1, N_ITEMS, (Bs, C), generator=g)
    if kind == "one_row":
        iid[:, 0] = 9                      # two rows carry the whole batch: 1,024 positives on row 9, 101,376 negatives on
        iid[:, 1:] = 7                     # row 7 -- one bucket receives everything (spill list, chunked merge sort)
        uid[:] = 3
    else:
        iid[:, ::3] = torch.randint(1000, 1400, (Bs, (C + 2) // 3), generator=g)          # ~34k pairs in one 512-row bucket
        iid[:, 1::7] = torch.randint(900_000, 900_003, (Bs, len(range(1, C, 7))), generator=g)   # three very long rows
        uid[::2] = torch.randint(5000, 5040, (Bs // 2,), generator=g)                    # hot users
    f = {"user_id": uid.to(_dev()), "item_id": iid.to(_dev()), "batch_size": Bs, "phase": "train"}
    m = _model()
    U, I = m.u_embeddings.weight.data, m.i_embeddings.weight.data
    eu, ei = m.optimizer.entry(m.u_embeddings.weight), m.optimizer.entry(m.i_embeddings.weight)
    for t in (1, 2):
        uu, ui = torch.unique(uid), torch.unique(iid)
        uud, uid_d = uu.to(_dev()), ui.to(_dev())
        before = [x[idx].cpu() for x, idx in ((U, uud), (eu["m"], uud), (eu["v"], uud), (I, uid_d), (ei["m"], uid_d),
                                              (ei["v"], uid_d))]
        loss = m.train_step(f)
        torch.cuda.synchronize()
        _, loss_o, gU, gI = _compact_oracle_grads(before[0], before[3], uu, ui, uid, iid)
        assert abs(float(loss) - loss_o) <= 1e-5
        for name, (w0, m0, v0), gr, W, idx in (("U", before[0:3], gU, U, uud), ("I", before[3:6], gI, I, uid_d)):
            w1, _, _ = _lazy_adam(w0, m0, v0, gr, t)
            ok = gr.abs() > 1e-6
            assert bool(ok.any())
            assert ((W[idx].cpu().double() - w1).abs()[ok]).max() <= 2e-6, (kind, name, t)
    ops.check_ids()
    m2 = _model()                                             # and the same bits on a second run
    for t in (1, 2):
        m2.train_step(f)
    torch.cuda.synchronize()
    assert torch.equal(m2.i_embeddings.weight.data, I) and torch.equal(m2.u_embeddings.weight.data, U)
