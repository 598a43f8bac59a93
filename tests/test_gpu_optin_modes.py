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
        err = (v.cpu() - ref.p[k].detach()).abs().reshape(-1)
        if l2 > 0:
            assert err.max() <= 2e-5, k                    # g includes l2 * w >> eps: every entry is well-conditioned
        else:
            # without weight decay some entries see |g| within a few orders of Adam's eps = 1e-8; there the step
            # lr * m_hat / (sqrt(v_hat) + eps) turns a 1e-10 rounding difference in g into ~1e-4 (for torch.optim on
            # two machines just as well): bound the bulk tightly and the ill-conditioned tail by a few such steps
            assert float(err.quantile(0.995)) <= 2e-5 and float(err.max()) <= 1e-3, (k, float(err.max()))


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
