"""CPU: pin oracle/rechorus_oracle.py against fixtures produced by the reference's own classes."""
import numpy as np
import pytest
import torch

from oracle import rechorus_oracle as O
from tests import golden_util as G

ALL = [(m, f) for m, fs in G.MODEL_FIXTURES.items() for f in fs]


@pytest.mark.parametrize("model,fixture", ALL)
def test_oracle_matches_reference_fixture(model, fixture):
    meta, w, batch, pred_ref, loss_ref, g_ref = G.load(fixture)
    pred, loss, grads = O.loss_and_grads(model, w, batch, num_heads=meta.get("num_heads", 4))
    # same ATen ops, possibly different association order -> fp32 rounding only
    assert torch.allclose(pred, pred_ref, rtol=1e-5, atol=1e-6), (pred - pred_ref).abs().max()
    assert abs(float(loss) - loss_ref) <= 1e-6
    assert set(grads) == set(g_ref)
    for k in g_ref:
        assert torch.allclose(grads[k], g_ref[k], rtol=1e-4, atol=1e-6), (k, (grads[k] - g_ref[k]).abs().max())
    # ranks are the integer output of the path (BaseRunner.py:63): must agree exactly on the fixture
    assert np.array_equal(O.gt_rank(pred.numpy()), O.gt_rank(pred_ref.numpy()))


@pytest.mark.parametrize("model,fixture", ALL)
def test_oracle_fp64_replay_agrees(model, fixture):
    """fp64 replay of the same weights: the fp32 reference output sits within 1e-5 of it (tie-breaker
    for the north_star tolerance)."""
    meta, w, batch, pred_ref, loss_ref, _ = G.load(fixture)
    w64 = {k: v.double() for k, v in w.items()}
    pred64 = O.scores(model, w64, batch, num_heads=meta.get("num_heads", 4))
    assert (pred64 - pred_ref.double()).abs().max() < 1e-5
    assert abs(float(O.bpr_loss(pred64)) - loss_ref) < 1e-5


def test_loss_closed_form_gradient_matches_autograd():
    g = torch.Generator().manual_seed(3)
    for B, C in [(5, 2), (7, 10), (3, 100)]:
        pred = (torch.randn(B, C, generator=g, dtype=torch.float64) * 2).requires_grad_(True)
        loss = O.bpr_loss(pred)
        loss.backward()
        l2, g2 = O.bpr_loss_and_grad_fp64(pred.detach().numpy())
        assert abs(float(loss) - l2) < 1e-12
        assert np.abs(pred.grad.numpy() - g2).max() < 1e-12


def test_loss_clamp_window_has_zero_gradient():
    pred = np.zeros((2, 3))
    pred[0] = [-50.0, 10.0, 10.0]      # S ~ e^-60 < 1e-8 -> clamped, zero gradient for that row
    pred[1] = [0.3, 0.1, -0.2]
    loss, g = O.bpr_loss_and_grad_fp64(pred)
    assert np.all(g[0] == 0.0) and np.any(g[1] != 0.0)
    t = torch.tensor(pred, requires_grad=True)
    O.bpr_loss(t).backward()
    assert np.abs(t.grad.numpy() - g).max() < 1e-12


def test_runner_metrics_and_shuffle_fixture():
    z = np.load(G.GOLDEN_DIR + "/runner_metrics.npz")
    assert np.array_equal(O.gt_rank(z["pred"]), z["gt_rank"])
    res = O.rank_metrics(z["pred"], [1, 5, 10, 50], ["HR", "NDCG"])
    for k, v in res.items():
        assert v == pytest.approx(float(z["m:" + k]), abs=1e-12)
    torch.manual_seed(99)
    item_id = torch.from_numpy(z["sh:item_id"])
    shuffled, perm = O.shuffle_candidates(item_id)          # same CPU RNG stream as BaseRunner.py:189
    assert np.array_equal(perm.numpy(), z["sh:indices"])
    assert np.array_equal(shuffled.numpy(), z["sh:shuffled"])
    restored = O.unshuffle_scores(torch.from_numpy(z["sh:scores"]), perm)
    assert np.array_equal(restored.numpy(), z["sh:restored"])


def test_negative_sampler_rejects_clicked_items():
    rng = np.random.RandomState(0)
    clicked = {u: set(range(1, 40)) for u in range(5)}
    neg = O.sample_negatives([0, 1, 2, 3, 4] * 20, clicked, 50, 3, rng)
    assert neg.min() >= 40 and neg.max() < 50


@pytest.mark.parametrize("wd", [0.0, 1e-2])
def test_lazy_exact_adam_equals_dense_adam(wd):
    """the bookkeeping planned for the kernels' exact-Adam mode (advance a row through its skipped steps before it
    is read or updated) reproduces the reference's dense torch.optim.Adam, forward reads included"""
    torch.manual_seed(0)
    n, d = 12, 5
    W0 = torch.randn(n, d, dtype=torch.float64)
    dense = W0.clone().requires_grad_(True)
    opt = torch.optim.Adam([dense], lr=1e-2, weight_decay=wd)
    lazy = O.LazyExactAdam(W0.clone(), lr=1e-2, weight_decay=wd)
    for _ in range(40):
        rows = torch.unique(torch.randint(0, n, (3,)))
        seen = lazy.read(rows)
        assert torch.allclose(seen, dense.detach()[rows], rtol=0, atol=1e-12)
        g = torch.randn(len(rows), d, dtype=torch.float64) * seen          # a gradient that depends on what was read
        opt.zero_grad()
        dense.grad = torch.zeros_like(dense)
        dense.grad[rows] = g
        opt.step()
        lazy.step(rows, g)
    lazy.flush()
    assert (lazy.W - dense.detach()).abs().max() <= 1e-12


def test_philox_known_answers_and_sampler_reference_properties():
    """Random123's known-answer vectors for Philox4x32-10 pin the generator the device sampler is defined on; the
    sampler reference has the reference's distribution properties (range, clicked-set rejection, determinism)."""
    kat = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
            [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for ctr, key, want in kat:
        assert O.philox4x32_10(ctr, key) == want
    n_items = 40
    clicked = {u: set(range(1 + u, n_items, 3)) for u in range(5)}          # a third of the catalogue each
    users = [0, 1, 2, 3, 4] * 200
    neg = O.device_sampler_reference(users, clicked, n_items, 4, seed=7, epoch=1)
    assert neg.min() >= 1 and neg.max() < n_items
    for i, u in enumerate(users):
        assert not (set(neg[i].tolist()) & clicked[u])
    assert np.array_equal(neg, O.device_sampler_reference(users, clicked, n_items, 4, seed=7, epoch=1))
    assert not np.array_equal(neg, O.device_sampler_reference(users, clicked, n_items, 4, seed=7, epoch=2))
    # uniform over the allowed items of user 0: 1000 x 4 / 5 = 800 draws over 26 items
    draws = neg[0::5].reshape(-1)
    allowed = sorted(set(range(1, n_items)) - clicked[0])
    counts = np.array([(draws == a).sum() for a in allowed])
    expected = len(draws) / len(allowed)
    assert ((counts - expected) ** 2 / expected).sum() < 2.5 * len(allowed)   # chi-square, generous bound


def test_test_all_protocol_matches_reference_predict_fixture():
    """tests/golden/eval_test_all.npz = the reference's own BaseRunner.predict under --test_all 1 (candidate list,
    BPRMF scores, clicked-item masking): the oracle's restatement reproduces the masked predictions, the integer ranks
    and the metrics"""
    import os
    import sys
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, golden)
    import fit_corpus
    gold = np.load(os.path.join(golden, "eval_test_all.npz"))
    corpus = fit_corpus.build()
    U, I = torch.from_numpy(gold["w:u_embeddings.weight"]), torch.from_numpy(gold["w:i_embeddings.weight"])
    uid, target = torch.from_numpy(gold["user_id"]), torch.from_numpy(gold["item_id"])
    clicked = [corpus.train_clicked_set[int(u)] | corpus.residual_clicked_set[int(u)] for u in uid]
    pred = O.test_all_predictions(U[uid], I, target, clicked)
    ref = gold["pred"]
    assert pred.shape == ref.shape
    assert np.array_equal(np.isneginf(pred), np.isneginf(ref))
    finite = np.isfinite(ref)
    assert np.abs(pred[finite] - ref[finite]).max() <= 1e-5
    assert np.array_equal(O.gt_rank(pred), gold["gt_rank"])
    m = O.rank_metrics(pred, [5, 10, 20], ["HR", "NDCG"])
    for k in gold.files:
        if k.startswith("m:"):
            assert abs(m[k[2:]] - float(gold[k])) <= 1e-12
