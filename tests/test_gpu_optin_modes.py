"""GPU tests of the opt-in modes: exact dense-Adam results from the row-sparse kernels (--exact_adam 1,
csrc/adam_exact.cu), the on-device negative sampler (--device_sampler, csrc/sampler.cu), SASRec's
one-query last block (csrc/attention_last.cu, default on) and the peer-store score return of the sharded path."""
import os
import types

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu]


@pytest.mark.parametrize("wd", [0.0, 1e-2])
@pytest.mark.parametrize("d", [32, 64, 128])
def test_adam_exact_advance_equals_dense_adam_on_zero_gradients(d, wd):
    """rows updated once, then left alone for k steps: b2r_adam_exact_advance must move them like torch.optim.Adam
    over the whole table does (momentum, weight decay), listed rows only, then the flush form"""
    from rechorus_b200 import ops
    torch.manual_seed(d)
    n = 200
    W0 = torch.randn(n, d) * 0.1
    dense = W0.clone().cuda().requires_grad_(True)
    opt = torch.optim.Adam([dense], lr=1e-2, weight_decay=wd)
    g0 = torch.zeros(n, d)
    touched = torch.arange(0, n, 3)
    g0[touched] = torch.randn(len(touched), d)
    dense.grad = g0.cuda()
    opt.step()                                              # step 1: the only data gradient
    st = opt.state[dense]
    W, m, v = dense.detach().clone(), st["exp_avg"].clone(), st["exp_avg_sq"].clone()
    if wd == 0.0:                                           # a row-sparse table would not have touched the others
        mask = torch.ones(n, dtype=torch.bool)
        mask[touched] = False
        assert float(m[mask.cuda()].abs().max()) == 0.0
    last = torch.ones(n, dtype=torch.int32, device="cuda")
    K = 37
    for _ in range(K):                                      # steps 2 .. K+1 with zero data gradient everywhere
        dense.grad = torch.zeros_like(dense)
        opt.step()
    cfg = types.SimpleNamespace(lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd)
    some = touched[: len(touched) // 2].cuda()
    ops.adam_exact_advance(W, m, v, last, K + 1, cfg, rows=some)
    assert (W[some] - dense.detach()[some]).abs().max() <= 2e-6
    assert int(last[some].min()) == K + 1 and int(last.max()) == K + 1
    ops.adam_exact_advance(W, m, v, last, K + 1, cfg)       # flush: everything else
    assert (W - dense.detach()).abs().max() <= 2e-6
    assert (m - st["exp_avg"]).abs().max() <= 1e-6 and (v - st["exp_avg_sq"]).abs().max() <= 1e-6
    assert int(last.min()) == K + 1
    ops.check_ids()


@pytest.mark.parametrize("l2", [0.0, 1e-3])
def test_exact_adam_mode_reproduces_dense_adam_training(l2):
    """BPRMF trained for a number of steps through the plugin surface with RowSparseOptimizer(exact_dense=True): after
    the flush the tables equal the reference-style training with dense torch.optim.Adam on the same batches (most
    rows are skipped by most batches, so momentum / weight-decay catch-up is what is being tested)."""
    import argparse
    from oracle import rechorus_oracle as O
    from rechorus_b200 import plugin
    from rechorus_b200.optim import RowSparseOptimizer
    p = plugin.BPRMF.parse_model_args(argparse.ArgumentParser())
    a = p.parse_args(["--emb_size", "64", "--num_neg", "4", "--table_mode", "fused"])
    a.device, a.model_path = torch.device("cuda", 0), "/tmp/_b2r_exact.pt"
    corpus = types.SimpleNamespace(n_users=60, n_items=90)
    torch.manual_seed(21)
    model = plugin.BPRMF(a, corpus).to(a.device)
    with torch.no_grad():
        for prm in model.parameters():
            prm.mul_(30.0)                                  # gradients well above rounding (Adam divides by |g|)
    w0 = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    model.optimizer = RowSparseOptimizer(model, "Adam", lr=1e-2, l2=l2, exact_dense=True)
    ref = O.ReferenceStyleTrainer("BPRMF", w0, lr=1e-2, l2=l2, optimizer="Adam")
    g = torch.Generator().manual_seed(22)
    model.train()
    # Without weight decay the comparison is CHAOTIC at the 1e-4 level, for torch.optim on two machines as much as for
    # these kernels: Adam's step lr * m_hat / (sqrt(v_hat) + 1e-8) turns a 1e-10 rounding difference in a gradient entry
    # that is within a few orders of eps (a candidate with a ~0 softmax weight has a ~1e-8 gradient ROW) into ~1e-4 of
    # weight; that perturbed item row then enters the user gradients sum_c g_c * I[c], and where those cancel (|dU| ~ 1e-3
    # from terms ~1e-2) a 7e-5 perturbation is ~1 % of the entry -- which Adam's normalisation again amplifies
    # (traced entry by entry with tools/diag_exact_adam.py).  With weight decay every gradient entry carries l2 * w >> eps
    # and the run is compared at 2e-5.
    for step in range(25):
        uid = torch.randint(1, 60, (8,), generator=g)
        iid = torch.randint(1, 90, (8, 5), generator=g)
        model.optimizer.zero_grad()
        out = model({"user_id": uid.cuda(), "item_id": iid.cuda(), "batch_size": 8, "phase": "train"})
        loss = model.loss(out)
        loss.backward()
        model.optimizer.step()
        ref_loss = ref.step({"user_id": uid, "item_id": iid}, shuffle=False)
        assert abs(float(loss) - ref_loss) <= 2e-5, (step, float(loss), ref_loss)
    model.optimizer.flush()
    for k, v in model.state_dict().items():
        err = (v.cpu() - ref.p[k].detach()).abs()
        if l2 > 0:
            assert float(err.max()) <= 2e-5, (k, float(err.max()))
        else:
            assert float(err.max()) <= 2e-4 and float(err.reshape(-1).quantile(0.9)) <= 2e-6, (k, float(err.max()))


def test_device_negative_sampler_equals_its_cpu_definition_bit_for_bit():
    """integer work: the kernel's output must equal oracle.device_sampler_reference exactly; and it must have the
    reference sampler's properties (range [1, n_items), never a training click of the row's user)"""
    from oracle import rechorus_oracle as O
    from rechorus_b200 import ops
    rng = np.random.RandomState(3)
    n_users, n_items, K = 37, 211, 6
    clicked = {u: set(rng.randint(1, n_items, rng.randint(0, 60)).tolist()) for u in range(n_users)}
    clicked[5] = set(range(1, n_items - 2))                 # almost everything clicked: long rejection chains
    users = rng.randint(0, n_users, 500)
    sampler = ops.DeviceNegativeSampler(clicked, n_users, n_items, torch.device("cuda", 0), seed=0x1234567890ABCDEF)
    got = sampler.sample(torch.from_numpy(users).cuda(), K, epoch=3).cpu().numpy()
    want = O.device_sampler_reference(users.tolist(), clicked, n_items, K, seed=0x1234567890ABCDEF, epoch=3)
    assert got.dtype == np.int64 and np.array_equal(got, want)
    assert got.min() >= 1 and got.max() < n_items
    for i, u in enumerate(users):
        assert not (set(got[i].tolist()) & clicked[int(u)])
    ops.check_ids()


def test_runner_with_device_sampler_trains_on_valid_negatives():
    import argparse
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import fit_corpus
    from rechorus_b200 import plugin
    from rechorus_b200.runner import BaseRunner
    p = argparse.ArgumentParser()
    p = BaseRunner.parse_runner_args(p)
    p = plugin.BPRMF.parse_model_args(p)
    a = p.parse_args(["--emb_size", "64", "--num_neg", "5", "--batch_size", "64", "--num_workers", "0", "--lr", "0.01",
                      "--table_mode", "fused", "--fused_optimizer", "1", "--fused_step", "1", "--device_sampler", "77"])
    a.device, a.model_path, a.log_file = torch.device("cuda", 0), "/tmp/_b2r_ds.pt", ""
    corpus = fit_corpus.build()
    torch.manual_seed(1)
    model = plugin.BPRMF(a, corpus).to(a.device)
    train = plugin.BPRMF.Dataset(model, corpus, "train")
    runner = BaseRunner(a)
    l1 = runner.fit(train, epoch=1)
    neg1 = np.array(train.data["neg_items"])
    l2 = runner.fit(train, epoch=2)
    neg2 = np.array(train.data["neg_items"])
    assert np.isfinite(l1) and np.isfinite(l2)
    assert neg1.shape == (len(train), 5) and neg1.min() >= 1 and neg1.max() < corpus.n_items
    assert not np.array_equal(neg1, neg2)                   # a fresh draw every epoch
    for i, u in enumerate(train.data["user_id"]):
        assert not (set(neg1[i].tolist()) & corpus.train_clicked_set[u])


def test_collate_general_equals_host_collate_bit_for_bit():
    """integer work: b2r_collate_general vs GeneralModel.Dataset._get_feed_dict + collate_batch (BaseModel.py:192-203,135-152)
    restated with torch indexing; identity and permuted order, ragged last batch"""
    from rechorus_b200 import ops
    g = torch.Generator().manual_seed(12)
    N, K, B = 1000, 7, 96
    users = torch.randint(1, 50, (N,), generator=g).cuda()
    items = torch.randint(1, 80, (N,), generator=g).cuda()
    neg = torch.randint(1, 80, (N, K), generator=g).cuda()
    for perm in (None, torch.randperm(N, generator=g).cuda()):
        for start in (0, 96 * 5, N - 40):
            Bn = min(B, N - start)
            ou = torch.full((B,), -1, dtype=torch.int64, device="cuda")
            oi = torch.full((B, K + 1), -1, dtype=torch.int64, device="cuda")
            ops.collate_general(users, items, neg, perm, start, Bn, ou, oi)
            rows = torch.arange(start, start + Bn, device="cuda") if perm is None else perm[start:start + Bn]
            assert torch.equal(ou[:Bn], users[rows])
            assert torch.equal(oi[:Bn], torch.cat([items[rows].unsqueeze(1), neg[rows]], dim=1))
            assert bool((ou[Bn:] == -1).all()) and bool((oi[Bn:] == -1).all())        # nothing beyond the batch is written


@pytest.mark.parametrize("model_name", ["BPRMF", "NeuMF"])
def test_fit_on_device_epoch_equals_the_same_batches_fed_by_hand(model_name):
    """--device_batches: an epoch whose negatives, row order and collate come from device kernels must train exactly as the
    same batches handed to the model one by one; its negatives are never training clicks of the row's user"""
    import argparse
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import fit_corpus
    from rechorus_b200 import ops, plugin
    from rechorus_b200.optim import RowSparseOptimizer
    from rechorus_b200.runner import BaseRunner
    cls = getattr(plugin, model_name)

    def build():
        p = argparse.ArgumentParser()
        p = BaseRunner.parse_runner_args(p)
        p = cls.parse_model_args(p)
        a = p.parse_args(["--emb_size", "64", "--num_neg", "5", "--batch_size", "64", "--num_workers", "0", "--lr", "0.05",
                          "--optimizer", "SGD", "--table_mode", "fused", "--fused_optimizer", "1", "--fused_step", "1",
                          "--device_batches", "77"])
        a.device, a.model_path, a.log_file = torch.device("cuda", 0), "/tmp/_b2r_db.pt", ""
        corpus = fit_corpus.build()
        torch.manual_seed(1)
        model = cls(a, corpus).to(a.device)
        with torch.no_grad():
            for prm in model.parameters():
                prm.mul_(20.0)
        return a, corpus, model, cls.Dataset(model, corpus, "train"), BaseRunner(a)

    a, corpus, model, train, runner = build()
    l1 = runner.fit(train, epoch=1)                          # routed to fit_on_device by the flag
    st = train.__dict__["_b2r_dev"]
    N, K, B = st["users"].numel(), 5, 64
    neg = st["sampler"].sample(st["users"], K, 1)            # same (seed, epoch) -> same draw
    gen = torch.Generator(device="cuda")
    gen.manual_seed(st["seed"] * 1_000_003 + 1)
    perm = torch.randperm(N, device="cuda", generator=gen)
    users_h = st["users"].cpu().numpy()
    negs_h = neg.cpu().numpy()
    for i in range(N):
        assert not (set(negs_h[i].tolist()) & corpus.train_clicked_set[int(users_h[i])])
    assert negs_h.min() >= 1 and negs_h.max() < corpus.n_items
    # the same epoch by hand on a fresh, identically initialised model
    _, _, model2, _, _ = build()
    model2.optimizer = RowSparseOptimizer(model2, "SGD", lr=0.05, l2=0.0)
    model2.set_table_mode("fused")
    model2.train()
    losses = []
    for k in range((N + B - 1) // B):
        rows = perm[k * B:(k + 1) * B]
        feed = {"user_id": st["users"][rows].contiguous(), "batch_size": rows.numel(), "phase": "train",
                "item_id": torch.cat([st["items"][rows].unsqueeze(1), neg[rows]], dim=1).contiguous()}
        if hasattr(model2, "train_step"):
            losses.append(model2.train_step(feed))
        else:
            model2.optimizer.zero_grad()
            ls = model2.loss(model2(feed))
            ls.backward()
            model2.optimizer.step()
            losses.append(ls.detach())
    assert abs(float(torch.stack(losses).mean()) - l1) <= 1e-6
    for (k, pa), (_, pb) in zip(model.named_parameters(), model2.named_parameters()):
        assert torch.equal(pa, pb), k
    l2 = runner.fit(train, epoch=2)
    assert np.isfinite(l2) and abs(l2 - l1) > 0              # a second epoch runs (fresh negatives, fresh order)
    ops.check_ids()
