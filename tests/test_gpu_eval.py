"""GPU parity of the evaluation kernels: integer ranks must equal the oracle's bit for bit on the
golden fixtures and on seeded cases whose score gaps exceed the kernels' rounding; the test_all path is checked
against the oracle's materialised [B, n_items] predictions including the clicked-item masking."""
import numpy as np
import pytest
import torch

from tests import golden_util
from oracle import rechorus_oracle as O

pytestmark = pytest.mark.gpu


def _ops():
    from rechorus_b200 import ops
    return ops


@pytest.mark.parametrize("N,C", [(1, 1), (7, 2), (300, 100), (64, 1024), (5, 5000), (3, 100000)])
def test_gt_rank_matches_numpy_bit_exact(N, C):
    ops = _ops()
    g = torch.Generator().manual_seed(N * 131 + C)
    pred = torch.randn(N, C, generator=g)
    pred[:, C // 2] = pred[:, 0]                        # a tie counts against the ground truth
    if C > 3:
        pred[0, 3] = float("nan")                       # NaN compares false, as in NumPy
        pred[N - 1, 1] = float("-inf")
    rank = ops.gt_rank(pred.cuda()).cpu().numpy()
    assert rank.dtype == np.int64
    assert np.array_equal(rank, O.gt_rank(pred.numpy()))


def test_gt_rank_on_golden_fixture_predictions():
    ops = _ops()
    for name in ("bprmf_k9_trained", "neumf_l64_32_16", "sasrec_l2h4"):
        pred = golden_util.load(name)[3]
        assert np.array_equal(ops.gt_rank(pred.cuda()).cpu().numpy(), O.gt_rank(pred.numpy()))


def test_histogram_metrics_equal_evaluate_method():
    ops = _ops()
    g = torch.Generator().manual_seed(5)
    pred = torch.randn(4000, 100, generator=g)
    pred[:, 0] += 1.5
    rank = ops.gt_rank(pred.cuda())
    hist = ops.rank_histogram(rank, 50)
    assert int(hist.sum()) == 4000
    ours = ops.metrics_from_histogram(hist, 4000, [5, 10, 20, 50], ["HR", "NDCG"])
    ref = O.rank_metrics(pred.numpy(), [5, 10, 20, 50], ["HR", "NDCG"])
    for k, v in ref.items():
        assert abs(ours[k] - v) <= 1e-12, k


@pytest.mark.parametrize("B,n_items,d", [(1, 2, 4), (5, 130, 20), (128, 1000, 64), (200, 4097, 64), (37, 777, 128)])
def test_rank_all_items_equals_materialised_test_all(B, n_items, d):
    ops = _ops()
    g = torch.Generator().manual_seed(B + n_items + d)
    q = torch.randn(B, d, generator=g)
    table = torch.randn(n_items, d, generator=g)
    target = torch.randint(1, n_items, (B,), generator=g)
    rng = np.random.RandomState(B)
    # clicked sets as the reference builds them: the user's other interactions, sometimes including the target
    clicked = []
    for b in range(B):
        s = set(rng.randint(1, n_items, size=rng.randint(0, min(40, n_items))).tolist())
        if b % 3 == 0:
            s.add(int(target[b]))
        clicked.append(sorted(s))
    rows = torch.tensor([b for b, s in enumerate(clicked) for _ in s], dtype=torch.int64)
    cols = torch.tensor([j for s in clicked for j in s], dtype=torch.int64)
    for use_mask in (False, True):
        ref_pred = O.test_all_predictions(q, table, target, clicked if use_mask else None)
        ref = O.gt_rank(ref_pred)
        ours = ops.rank_all_items(q.cuda(), table.cuda(), target.cuda(), rows.cuda() if use_mask else None,
                                  cols.cuda() if use_mask else None).cpu().numpy()
        # ranks are exact wherever no candidate sits within rounding distance of the target's score
        s0 = ref_pred[:, :1]
        finite = np.isfinite(ref_pred)
        near = (np.abs(np.where(finite, ref_pred, np.inf) - s0) <= 1e-4 * (1 + np.abs(s0))).sum(axis=1)
        # the target's own columns (0, and its item-id column when unmasked) are exact ties in both implementations
        self_cols = 1 + np.array([0 if (use_mask and int(target[b]) in clicked[b]) else 1 for b in range(B)])
        clean = near == self_cols
        assert clean.mean() > 0.9
        assert np.array_equal(ours[clean], ref[clean])
        assert np.abs(ours - ref).max() <= (near - self_cols).max()
    ops.check_ids()


def test_rank_all_items_flags_out_of_range_target():
    ops = _ops()
    q = torch.randn(4, 8).cuda()
    table = torch.randn(10, 8).cuda()
    ops.rank_all_items(q, table, torch.tensor([1, 2, 10, 3]).cuda())
    with pytest.raises(IndexError):
        ops.check_ids()


def test_model_eval_ranks_candidates_and_test_all_agree_with_oracle():
    """plugin surface: BPRMF.eval_ranks on a 100-candidate batch and under test_all."""
    import argparse
    import types
    from rechorus_b200 import plugin
    p = plugin.BPRMF.parse_model_args(argparse.ArgumentParser())
    a = p.parse_args(["--emb_size", "64", "--test_all", "0"])
    a.device, a.model_path = torch.device("cuda", 0), "/tmp/_b2r_eval.pt"
    corpus = types.SimpleNamespace(n_users=50, n_items=300)
    torch.manual_seed(3)
    model = plugin.BPRMF(a, corpus).to(a.device)
    with torch.no_grad():
        for prm in model.parameters():
            prm.mul_(30.0)                                # trained-scale scores: gaps far above rounding
    w = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(4)
    uid = torch.randint(1, 50, (32,), generator=g)
    iid = torch.randint(1, 300, (32, 100), generator=g)
    feed = {"user_id": uid.cuda(), "item_id": iid.cuda(), "batch_size": 32, "phase": "test"}
    ranks = model.eval_ranks(feed).cpu().numpy()
    assert np.array_equal(ranks, O.gt_rank(O.bprmf_scores(w, uid, iid).numpy()))
    # test_all: the feed carries [target] + arange(1, n_items) exactly as BaseModel.py:194-198 builds it
    model.test_all = 1
    allc = torch.cat([iid[:, :1], torch.arange(1, 300).view(1, -1).expand(32, -1)], dim=1)
    feed_all = {"user_id": uid.cuda(), "item_id": allc.cuda(), "batch_size": 32, "phase": "test"}
    clicked = [sorted(set(iid[b, 1:6].tolist())) for b in range(32)]
    rows = torch.tensor([b for b, s in enumerate(clicked) for _ in s])
    cols = torch.tensor([j for s in clicked for j in s])
    ours = model.eval_ranks(feed_all, rows.cuda(), cols.cuda()).cpu().numpy()
    ref = O.gt_rank(O.test_all_predictions(F_embed(w, uid), w["i_embeddings.weight"], iid[:, 0], clicked))
    assert np.array_equal(ours, ref)


def F_embed(w, uid):
    return torch.nn.functional.embedding(uid, w["u_embeddings.weight"])
