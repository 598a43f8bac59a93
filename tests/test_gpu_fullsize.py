"""GPU: BASELINE.json configs 3 and 4 at full size (1 M-row tables, B=4096).  The oracle cannot run the whole batch
in seconds, so: (i) a 48-sample slice of the batch is replayed through the CPU oracle with the same weights,
(ii) size-independent properties (loss of the slice vs the full-batch kernel outputs, permutation equivariance,
gradient checksums) cover the rest."""
import argparse
import types

import numpy as np
import pytest
import torch

from oracle import rechorus_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _build(name, extra, n_users, n_items):
    from rechorus_b200 import plugin
    from rechorus_b200.runner import BaseRunner
    cls = getattr(plugin, name)
    p = argparse.ArgumentParser()
    p = BaseRunner.parse_runner_args(p)
    p = cls.parse_model_args(p)
    a = p.parse_args(extra)
    a.device, a.model_path = torch.device("cuda", 0), "/tmp/_b2r_unused.pt"
    torch.manual_seed(0)
    m = cls(a, types.SimpleNamespace(n_users=n_users, n_items=n_items)).to(a.device)
    with torch.no_grad():
        for n_, q in m.named_parameters():
            if "layer_norm" not in n_:
                q.mul_(10.0)                      # trained-scale weights
    m.train()
    return m


def test_config3_neumf_full_size():
    from rechorus_b200 import ops
    n_users = n_items = 1_000_000
    B, C, d = 4096, 5, 64
    m = _build("NeuMF", ["--emb_size", "64", "--layers", "[64, 32, 16]", "--num_neg", "4"], n_users, n_items)
    g = torch.Generator().manual_seed(3)
    uid = torch.randint(1, n_users, (B,), generator=g)
    iid = torch.randint(1, n_items, (B, C), generator=g)
    feed = {"user_id": uid.cuda(), "item_id": iid.cuda(), "batch_size": B, "phase": "train"}
    out = m(feed)
    loss = m.loss(out)
    loss.backward()
    ops.check_ids()
    pred = out["prediction"].detach().cpu()
    w = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    sl = slice(100, 148)
    ref = O.neumf_scores(w, uid[sl], iid[sl])
    assert (pred[sl] - ref).abs().max() <= TOL
    assert abs(float(loss) - float(O.bpr_loss(pred))) <= TOL
    # candidate-permutation equivariance (helpers/BaseRunner.py:187-191)
    perm = torch.argsort(torch.rand(B, C), dim=-1)
    with torch.no_grad():
        p2 = m({**feed, "item_id": torch.gather(iid, 1, perm).cuda()})["prediction"].cpu()
    assert (p2 - torch.gather(pred, 1, perm)).abs().max() <= 1e-6
    # gradient checksum: the dense table gradients touch exactly the batch's rows
    gi = m.mlp_i_embeddings.weight.grad
    touched = torch.zeros(n_items, dtype=torch.bool)
    touched[iid.reshape(-1)] = True
    assert bool(((gi.abs().sum(1) > 0).cpu() <= touched).all())
    # bias gradient of the first MLP layer equals the column sum an fp64 replay of the slice cannot give; check the
    # whole-batch identity sum_p dL/dpred[p] * d pred/d b instead through autograd on a 48-sample sub-batch
    m.zero_grad()
    sub = {"user_id": uid[sl].cuda(), "item_id": iid[sl].cuda(), "batch_size": 48, "phase": "train"}
    m.loss(m(sub)).backward()
    _, _, gref = O.loss_and_grads("NeuMF", w, {"user_id": uid[sl], "item_id": iid[sl]})
    for k in ("mlp.0.weight", "mlp.0.bias", "mlp.2.weight", "prediction.weight"):
        got = dict(m.named_parameters())[k].grad.cpu()
        assert (got - gref[k]).abs().max() <= TOL, k


def test_config4_sasrec_full_size():
    from rechorus_b200 import ops
    n_items = 1_000_000
    B, C, L, d = 4096, 100, 50, 64
    m = _build("SASRec", ["--emb_size", "64", "--history_max", "50", "--num_layers", "2", "--num_heads", "4",
                          "--num_neg", "99"], 10, n_items)
    g = torch.Generator().manual_seed(4)
    lengths = torch.randint(1, L + 1, (B,), generator=g)
    lengths[0] = L
    hist = torch.randint(1, n_items, (B, L), generator=g) * (torch.arange(L).view(1, L) < lengths.view(B, 1))
    iid = torch.randint(1, n_items, (B, C), generator=g)
    feed = {"user_id": torch.zeros(B, dtype=torch.int64).cuda(), "item_id": iid.cuda(), "history_items": hist.cuda(),
            "lengths": lengths.cuda(), "batch_size": B, "phase": "train"}
    out = m(feed)
    loss = m.loss(out)
    loss.backward()
    ops.check_ids()
    pred = out["prediction"].detach().cpu()
    w = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    sl = slice(0, 48)
    ref = O.sasrec_scores(w, hist[sl], lengths[sl], iid[sl], 4)
    assert (pred[sl] - ref).abs().max() <= TOL
    assert abs(float(loss) - float(O.bpr_loss(pred))) <= TOL
    assert np.array_equal(O.gt_rank(pred[sl].numpy()), O.gt_rank(ref.numpy()))
    # padding row 0 of the item table and of the position table gets exactly zero gradient
    assert float(m.i_embeddings.weight.grad[0].abs().max()) == 0.0
    # the position-table gradient is a sum over ~B*L/2 positions: compare on the 48-sample sub-batch
    m.zero_grad()
    sub = {k: (v[sl] if isinstance(v, torch.Tensor) else v) for k, v in feed.items()}
    sub["batch_size"] = 48
    m.loss(m(sub)).backward()
    _, _, gref = O.loss_and_grads("SASRec", w, {"history_items": hist[sl], "lengths": lengths[sl], "item_id": iid[sl]})
    for k in ("p_embeddings.weight", "transformer_block.0.masked_attn_head.q_linear.weight",
              "transformer_block.1.linear2.bias", "transformer_block.0.layer_norm1.weight"):
        got = dict(m.named_parameters())[k].grad.cpu()
        assert (got - gref[k]).abs().max() <= 2e-5, (k, float((got - gref[k]).abs().max()))
