"""GPU parity: NeuMF and SASRec through the plugin surface vs the fixtures the reference itself produced
(tests/golden/*.npz), in all three gradient forms, plus the fused optimizer against the oracle."""
import argparse
import types

import numpy as np
import pytest
import torch

from oracle import rechorus_oracle as O
from tests import golden_util as G

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _build(model_name, meta, weights, mode="dense"):
    from rechorus_b200 import plugin
    from rechorus_b200.runner import BaseRunner
    cls = getattr(plugin, model_name)
    p = argparse.ArgumentParser()
    p = BaseRunner.parse_runner_args(p)
    p = cls.parse_model_args(p)
    extra = ["--emb_size", str(meta["d"]), "--table_mode", mode]
    if model_name == "NeuMF":
        extra += ["--layers", str(meta["layers"])]
    if model_name == "SASRec":
        extra += ["--history_max", str(meta["history_max"]), "--num_layers", str(meta["num_layers"]),
                  "--num_heads", str(meta["num_heads"])]
    a = p.parse_args(extra)
    a.device, a.model_path = torch.device("cuda", 0), "/tmp/_b2r_unused.pt"
    m = cls(a, types.SimpleNamespace(n_users=meta.get("n_users", 10), n_items=meta["n_items"])).to(a.device)
    m.load_state_dict(weights)
    m.set_table_mode(mode)
    m.train()
    return m


def _feed(batch, B):
    out = {k: v.cuda() for k, v in batch.items()}
    out["batch_size"], out["phase"] = B, "train"
    return out


CASES = [(m, f) for m in ("NeuMF", "SASRec") for f in G.MODEL_FIXTURES[m]]


@pytest.mark.parametrize("model_name,fixture", CASES)
def test_forward_loss_backward_match_reference_fixture(model_name, fixture):
    from rechorus_b200 import ops
    meta, w, batch, pred_ref, loss_ref, g_ref = G.load(fixture)
    m = _build(model_name, meta, w)
    out = m(_feed(batch, meta["B"]))
    loss = m.loss(out)
    loss.backward()
    ops.check_ids()
    pred = out["prediction"].detach().cpu()
    assert (pred - pred_ref).abs().max() <= TOL
    assert abs(float(loss) - loss_ref) <= TOL
    assert np.array_equal(O.gt_rank(pred.numpy()), O.gt_rank(pred_ref.numpy()))
    for k, p in m.named_parameters():
        got = p.grad.cpu() if p.grad is not None else torch.zeros_like(g_ref[k])
        assert (got - g_ref[k]).abs().max() <= TOL, (k, float((got - g_ref[k]).abs().max()))


@pytest.mark.parametrize("model_name,fixture", [("NeuMF", "neumf_l64_32_16"), ("SASRec", "sasrec_l2h4")])
def test_sparse_gradient_form_equals_dense(model_name, fixture):
    meta, w, batch, *_ = G.load(fixture)
    grads = {}
    for mode in ("dense", "sparse"):
        m = _build(model_name, meta, w, mode)
        m.loss(m(_feed(batch, meta["B"]))).backward()
        grads[mode] = {k: (p.grad.to_dense() if p.grad.is_sparse else p.grad).cpu() for k, p in m.named_parameters()}
    for k in grads["dense"]:
        assert (grads["dense"][k] - grads["sparse"][k]).abs().max() <= 1e-7, k


@pytest.mark.parametrize("model_name,fixture", [("NeuMF", "neumf_l64_32_16"), ("SASRec", "sasrec_l2h4")])
def test_fused_optimizer_step_matches_oracle_gradients(model_name, fixture):
    """One fused SGD step: W_new = W - lr * grad for touched table rows and for every dense parameter."""
    from rechorus_b200.optim import RowSparseOptimizer
    meta, w, batch, _, _, g_ref = G.load(fixture)
    m = _build(model_name, meta, w, "fused")
    opt = RowSparseOptimizer(m, "SGD", lr=0.1)
    opt.zero_grad()
    m.loss(m(_feed(batch, meta["B"]))).backward()
    opt.step()
    for k, p in m.named_parameters():
        want = w[k] - 0.1 * g_ref[k]
        assert (p.detach().cpu() - want).abs().max() <= 2e-6, k


def test_sasrec_is_invariant_to_extra_right_padding():
    meta, w, batch, pred_ref, *_ = G.load("sasrec_l2h4")
    m = _build("SASRec", meta, w)
    feed = _feed(batch, meta["B"])
    # the fixture already has Lb == history_max; drop the longest rows' tail instead: shorter Lb must still agree
    short = dict(feed)
    keep = int(batch["lengths"].sort().values[-2])          # second-longest length
    mask = batch["lengths"] <= keep
    short["history_items"] = feed["history_items"][mask.cuda()][:, :keep].contiguous()
    short["lengths"] = feed["lengths"][mask.cuda()]
    short["item_id"] = feed["item_id"][mask.cuda()]
    with torch.no_grad():
        a = m(short)["prediction"].cpu()
    assert (a - pred_ref[mask]).abs().max() <= TOL
