"""GPU end-to-end: the stand-alone BaseRunner.fit / evaluate surface driving the kernel-backed BPRMF for one epoch.
In 'dense' table mode with the stock torch.optim.Adam the runner builds (exact reference semantics), the final weights
must equal an oracle replay of the very same batches (recorded on the way) within 1e-5; the 'fused' mode must train
(loss goes down, dev metrics computed) through the model.optimizer seam."""
import argparse
import types

import numpy as np
import pandas as pd
import pytest
import torch

from oracle import rechorus_oracle as O

pytestmark = pytest.mark.gpu


def _corpus(n_users=40, n_items=60, rows=600, seed=0):
    rng = np.random.RandomState(seed)
    df = pd.DataFrame({"user_id": rng.randint(1, n_users, rows), "item_id": rng.randint(1, n_items, rows),
                       "time": np.arange(rows)})
    ev = df.iloc[:64].copy()
    ev["neg_items"] = [list(rng.randint(1, n_items, 19)) for _ in range(len(ev))]
    clicked = {u: set(df.item_id[df.user_id == u]) for u in range(n_users)}
    return types.SimpleNamespace(n_users=n_users, n_items=n_items, data_df={"train": df, "dev": ev, "test": ev},
                                 train_clicked_set=clicked, residual_clicked_set={u: set() for u in range(n_users)})


def _setup(mode, extra=()):
    from rechorus_b200 import plugin
    from rechorus_b200.runner import BaseRunner
    p = argparse.ArgumentParser()
    p = BaseRunner.parse_runner_args(p)
    p = plugin.BPRMF.parse_model_args(p)
    a = p.parse_args(["--emb_size", "64", "--num_neg", "3", "--batch_size", "64", "--num_workers", "0", "--lr", "0.01",
                      "--l2", "1e-5", "--table_mode", mode, "--topk", "5,10", *extra])
    a.device, a.model_path, a.log_file = torch.device("cuda", 0), "/tmp/_b2r_fit.pt", ""
    corpus = _corpus()
    torch.manual_seed(0)
    model = plugin.BPRMF(a, corpus).to(a.device)
    data = {ph: plugin.BPRMF.Dataset(model, corpus, ph) for ph in ("train", "dev", "test")}
    for d in data.values():
        d.prepare()
    return a, model, data, BaseRunner(a)


def _bpr_loss_pos_col(pred, pos_col):
    """BaseModel.py:182-185 with the positive sitting in column pos_col[b] (the loss is invariant to the order of
    the negatives, so this equals the loss on the un-shuffled prediction)."""
    B, C = pred.shape
    mask = torch.ones(B, C, dtype=torch.bool)
    mask[torch.arange(B), pos_col] = False
    pos = pred[torch.arange(B), pos_col].unsqueeze(1)
    neg = pred[mask].view(B, C - 1)
    w = torch.softmax(neg - neg.max(), dim=1)
    s = (torch.sigmoid(pos - neg) * w).sum(1)
    return -torch.log(s.clamp(1e-8, 1 - 1e-8)).mean()


def test_fit_epoch_dense_mode_weights_match_oracle_training():
    a, model, data, runner = _setup("dense")
    w0 = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    seen = []
    orig_forward = model.forward

    def recording_forward(feed):
        seen.append({k: v.detach().cpu().clone() for k, v in feed.items() if isinstance(v, torch.Tensor)})
        return orig_forward(feed)

    model.forward = recording_forward
    np.random.seed(1)
    torch.manual_seed(1)
    runner.fit(data["train"], epoch=1)
    model.forward = orig_forward
    # the positive of every training row (to find its column after the runner's shuffle)
    train = data["train"]
    pos_of = {}
    for u, i in zip(train.data["user_id"], train.data["item_id"]):
        pos_of.setdefault(int(u), set()).add(int(i))
    params = {k: v.clone().requires_grad_(True) for k, v in w0.items()}
    opt = torch.optim.Adam(O.param_groups(params.items(), 1e-5), lr=0.01)
    for feed in seen:
        uid, iid = feed["user_id"], feed["item_id"]
        # negatives were rejected against the user's training clicks (BaseModel.py:206-214), so exactly the columns
        # holding one of the user's clicked items are candidates for the positive; with duplicates among them the
        # loss value is the same whichever is taken (equal ids -> equal scores)
        pos_col = torch.tensor([next(c for c in range(iid.shape[1]) if int(iid[b, c]) in pos_of[int(uid[b])])
                                for b in range(iid.shape[0])])
        opt.zero_grad()
        _bpr_loss_pos_col(O.bprmf_scores(params, uid, iid), pos_col).backward()
        opt.step()
    for k, p in model.state_dict().items():
        assert (p.cpu() - params[k].detach()).abs().max() <= 1e-5, k


def test_fit_and_evaluate_with_fused_optimizer_seam():
    from rechorus_b200.optim import RowSparseOptimizer
    a, model, data, runner = _setup("fused", ["--fused_optimizer", "1"])
    np.random.seed(2)
    torch.manual_seed(2)
    before = runner.evaluate(data["dev"], [5, 10], ["HR", "NDCG"])
    l1 = runner.fit(data["train"], epoch=1)
    assert isinstance(model.optimizer, RowSparseOptimizer)
    l2 = runner.fit(data["train"], epoch=2)
    l3 = runner.fit(data["train"], epoch=3)
    assert l3 < l1 and np.isfinite(l2)
    after = runner.evaluate(data["dev"], [5, 10], ["HR", "NDCG"])
    assert set(after) == {"HR@5", "NDCG@5", "HR@10", "NDCG@10"} and all(0 <= v <= 1 for v in after.values())
    # dev rows are training rows here: three epochs must have improved the ranking of the positives
    assert after["HR@10"] >= before["HR@10"]
    # predictions on the eval path come from the no-grad inference hook and rank exactly like the oracle on the same weights
    w = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    batch = data["dev"].collate_batch([data["dev"][i] for i in range(16)])
    pred = model.inference({k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in batch.items()})["prediction"].cpu()
    ref = O.bprmf_scores(w, batch["user_id"], batch["item_id"])
    assert (pred - ref).abs().max() <= 1e-5
    assert np.array_equal(O.gt_rank(pred.numpy()), O.gt_rank(ref.numpy()))


@pytest.mark.parametrize("test_all", [0, 1])
def test_device_metrics_equal_host_evaluate_method(test_all):
    """--device_metrics 1 (ranks + histogram on the GPU) gives the metrics of the reference's predict ->
    evaluate_method route on the same weights, for the 100-candidate protocol and for test_all with clicked-item
    masking."""
    a, model, data, runner = _setup("fused", ["--fused_optimizer", "1", "--test_all", str(test_all)])
    np.random.seed(3)
    torch.manual_seed(3)
    for ep in range(3):
        runner.fit(data["train"], epoch=ep + 1)
    with torch.no_grad():
        for prm in model.parameters():
            prm.mul_(20.0)                 # spread the scores: rank gaps far above fp32 rounding
    if test_all:
        # the reference masks train + residual clicks; give some users residual clicks too
        corpus = data["dev"].corpus
        for u in range(0, corpus.n_users, 3):
            corpus.residual_clicked_set[u] = {1 + (u * 7) % (corpus.n_items - 1), 2}
    host = runner.evaluate(data["dev"], [1, 5, 10], ["HR", "NDCG"])
    runner.device_metrics = 1
    dev = runner.evaluate(data["dev"], [1, 5, 10], ["HR", "NDCG"])
    assert set(host) == set(dev)
    for k in host:
        assert abs(host[k] - dev[k]) <= 1e-12, (k, host[k], dev[k])
