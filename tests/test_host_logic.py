"""CPU: host-side mirror of the reference interface (datasets, collate, metrics, runner bookkeeping)."""
import argparse
import types

import numpy as np
import pandas as pd
import pytest
import torch

from oracle import rechorus_oracle as O
from rechorus_b200 import plugin
from rechorus_b200.runner import BaseRunner, format_metric
from tests import golden_util as G


def _corpus(n_users=6, n_items=12, rows=40, seed=0):
    rng = np.random.RandomState(seed)
    df = pd.DataFrame({"user_id": rng.randint(1, n_users, rows), "item_id": rng.randint(1, n_items, rows),
                       "time": np.arange(rows)})
    ev = df.iloc[:8].copy()
    ev["neg_items"] = [list(rng.randint(1, n_items, 5)) for _ in range(len(ev))]
    clicked = {u: set(df.item_id[df.user_id == u]) for u in range(n_users)}
    return types.SimpleNamespace(n_users=n_users, n_items=n_items, data_df={"train": df, "dev": ev, "test": ev},
                                 train_clicked_set=clicked, residual_clicked_set={u: set() for u in range(n_users)})


def _args(model_cls, extra=()):
    p = argparse.ArgumentParser()
    p = BaseRunner.parse_runner_args(p)
    p = model_cls.parse_model_args(p)
    a = p.parse_args(list(extra))
    a.device = torch.device("cpu")
    a.model_path = "/tmp/_b2r_unused.pt"
    return a


def test_runner_flags_match_reference_defaults():
    a = _args(plugin.BPRMF)
    assert (a.epoch, a.early_stop, a.lr, a.l2, a.batch_size, a.eval_batch_size) == (200, 10, 1e-3, 0, 256, 256)
    assert (a.optimizer, a.num_workers, a.topk, a.metric) == ("Adam", 5, "5,10,20,50", "NDCG,HR")
    assert (a.emb_size, a.num_neg, a.dropout, a.test_all, a.buffer) == (64, 1, 0, 0, 1)


def test_evaluate_method_matches_reference_fixture():
    z = np.load(G.GOLDEN_DIR + "/runner_metrics.npz")
    res = BaseRunner.evaluate_method(z["pred"], [1, 5, 10, 50], ["HR", "NDCG"])
    for k, v in res.items():
        assert float(v) == pytest.approx(float(z["m:" + k]), abs=1e-12)
    assert format_metric({"HR@5": 0.5, "NDCG@5": 0.25, "HR@10": 1.0}) == "HR@5:0.5000,NDCG@5:0.2500,HR@10:1.0000"
    with pytest.raises(ValueError):
        BaseRunner.evaluate_method(z["pred"], [5], ["MAP"])


def test_state_dict_keys_and_init_match_reference():
    corpus = _corpus()
    m = plugin.BPRMF(_args(plugin.BPRMF), corpus)
    assert sorted(m.state_dict()) == ["i_embeddings.weight", "u_embeddings.weight"]
    assert m.i_embeddings.weight.shape == (12, 64) and abs(float(m.i_embeddings.weight.std()) - 0.01) < 2e-3
    groups = m.customize_parameters()
    assert len(groups[0]["params"]) == 2 and groups[1]["weight_decay"] == 0
    # golden weights (made by the reference) load into our module unchanged
    meta, w, *_ = G.load("bprmf_k1")
    m2 = plugin.BPRMF(_args(plugin.BPRMF), types.SimpleNamespace(n_users=meta["n_users"], n_items=meta["n_items"]))
    m2.load_state_dict(w)


def test_general_dataset_feed_dict_collate_and_sampling():
    corpus = _corpus()
    model = plugin.BPRMF(_args(plugin.BPRMF, ["--num_neg", "3"]), corpus)
    train = plugin.BPRMF.Dataset(model, corpus, "train")
    np.random.seed(5)
    train.actions_before_epoch()
    np.random.seed(5)
    want = O.sample_negatives(train.data["user_id"], corpus.train_clicked_set, corpus.n_items, 3, np.random)
    assert np.array_equal(train.data["neg_items"], want)          # same NumPy RNG stream as BaseModel.py:206-214
    fd = train[0]
    assert fd["item_id"].shape == (4,) and fd["item_id"][0] == train.data["item_id"][0]
    batch = train.collate_batch([train[i] for i in range(5)])
    assert batch["item_id"].shape == (5, 4) and batch["item_id"].dtype == torch.int64
    assert batch["user_id"].dtype == torch.int64 and batch["batch_size"] == 5 and batch["phase"] == "train"
    dev = plugin.BPRMF.Dataset(model, corpus, "dev")
    dev.prepare()
    assert dev[0]["item_id"].shape == (6,)
    model.test_all = 1
    assert plugin.BPRMF.Dataset(model, corpus, "test")._get_feed_dict(0)["item_id"].shape == (12,)


def test_collate_right_pads_ragged_histories():
    ds = plugin.BaseModel.Dataset.__new__(plugin.BaseModel.Dataset)
    ds.phase = "train"
    out = ds.collate_batch([{"h": np.array([3, 4, 5]), "lengths": 3}, {"h": np.array([7]), "lengths": 1}])
    assert out["h"].tolist() == [[3, 4, 5], [7, 0, 0]] and out["lengths"].tolist() == [3, 1]


def test_eval_termination_rule():
    r = BaseRunner(_args(plugin.BPRMF, ["--early_stop", "3"]))
    assert not r.eval_termination([0.1, 0.2, 0.3])
    assert r.eval_termination([0.5, 0.4, 0.3, 0.2])            # non-increasing tail
    assert r.eval_termination([0.9, 0.1, 0.2, 0.3, 0.4])       # best is more than early_stop epochs ago
    # utils.non_increasing (utils/utils.py:103-104) compares the window's FIRST value with every later one -- not
    # consecutive pairs: [0.5, 0.3, 0.4] stops although 0.3 < 0.4
    assert r.eval_termination([0.1, 0.5, 0.3, 0.4])
    assert not r.eval_termination([0.1, 0.3, 0.5, 0.4])
    r1 = BaseRunner(_args(plugin.BPRMF, ["--early_stop", "1"]))
    assert r1.eval_termination([0.1, 0.2])                     # a one-element window is trivially non-increasing


def test_metrics_from_histogram_equal_evaluate_method():
    """HR@k / NDCG@k read off a rank histogram equal helpers/BaseRunner.py:52-78 applied to the predictions."""
    import numpy as np
    from oracle import rechorus_oracle as O
    from rechorus_b200 import ops
    rng = np.random.RandomState(0)
    pred = rng.randn(999, 60).astype(np.float32)
    pred[:, 0] += 1.0
    rank = O.gt_rank(pred)
    kmax = 20
    hist = np.bincount(np.minimum(rank, kmax + 1), minlength=kmax + 2)
    ours = ops.metrics_from_histogram(hist, len(rank), [1, 5, 20], ["HR", "NDCG"])
    ref = O.rank_metrics(pred, [1, 5, 20], ["HR", "NDCG"])
    assert set(ours) == set(ref)
    for k in ref:
        assert abs(ours[k] - ref[k]) <= 1e-12
    import pytest
    with pytest.raises(ValueError):
        ops.metrics_from_histogram(hist, len(rank), [5], ["MRR"])


def test_graph_step_flag_needs_the_fused_optimizer_and_excludes_exact_adam():
    """--graph_step replays the loop body from a CUDA graph: it needs the RowSparseOptimizer's device-side clock (so
    --fused_optimizer 1) and cannot keep the exact-Adam mode's per-step host bookkeeping"""
    with pytest.raises(ValueError):
        BaseRunner(_args(plugin.BPRMF, ["--graph_step", "1"]))
    with pytest.raises(ValueError):
        BaseRunner(_args(plugin.BPRMF, ["--graph_step", "1", "--fused_optimizer", "1", "--exact_adam", "1"]))
    r = BaseRunner(_args(plugin.BPRMF, ["--graph_step", "1", "--fused_optimizer", "1"]))
    assert r.graph_step == 1 and BaseRunner(_args(plugin.BPRMF, [])).graph_step == 0
