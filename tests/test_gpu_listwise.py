"""GPU: the list-wise loss kernels (b2r_listwise_loss) against the reference's own ImpressionModel.loss outputs
(tests/golden/listwise.npz) and against the oracle on larger ragged batches; the Impression model variants
(BPRMF.py:65-80, SASRec.py:107-122) forward / loss / backward.  Tolerance 1e-5 (north_star)."""
import argparse
import os
import types

import numpy as np
import pytest
import torch

from oracle import rechorus_oracle as O

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "listwise.npz"))
LOSSES = ["BPR", "BPRhard", "BPRafter", "BPRhardafter", "BPRbefore", "BPRhardbefore", "listnet", "softmaxCE", "attention_rank"]


@pytest.mark.parametrize("case", ["a", "b", "c"])
@pytest.mark.parametrize("loss_n", LOSSES)
def test_listwise_kernel_equals_reference_fixture(case, loss_n):
    from rechorus_b200 import ops
    pred = torch.from_numpy(GOLD[f"{case}:pred"]).cuda().requires_grad_(True)
    target = torch.from_numpy(GOLD[f"{case}:target"]).cuda()
    loss = ops.listwise_loss(pred, target, loss_n, int(GOLD[f"{case}:max_pos"]))
    (loss * 3.0).backward()                                   # upstream factor goes through the node
    want = float(GOLD[f"{case}:{loss_n}:loss"])
    assert abs(float(loss) - want) <= 1e-5 * max(1.0, abs(want))
    assert np.abs(pred.grad.cpu().numpy() / 3.0 - GOLD[f"{case}:{loss_n}:grad"]).max() <= 1e-5


@pytest.mark.parametrize("loss_n", LOSSES)
def test_listwise_kernel_equals_oracle_on_ragged_batches(loss_n):
    from rechorus_b200 import ops
    g = torch.Generator().manual_seed(len(loss_n))
    for B, max_pos, max_neg in [(257, 20, 20), (33, 2, 70), (5, 40, 3)]:
        Cn = max_pos + max_neg
        pred = torch.randn(B, Cn, generator=g) * 2
        target = torch.full((B, Cn), -1, dtype=torch.int64)
        for b in range(B):
            target[b, :int(torch.randint(1, max_pos + 1, (1,), generator=g))] = 1
            nneg = int(torch.randint(0 if loss_n in ("listnet", "softmaxCE", "attention_rank") and b % 7 == 3 else 1,
                                     max_neg + 1, (1,), generator=g))
            target[b, max_pos:max_pos + nneg] = 0             # CE forms: some rows without negatives (have_neg = 0)
        p_ref = pred.clone().requires_grad_(True)
        l_ref = O.listwise_loss(p_ref, target, loss_n, max_pos)
        l_ref.backward()
        p = pred.cuda().requires_grad_(True)
        loss = ops.listwise_loss(p, target.cuda(), loss_n, max_pos)
        loss.backward()
        assert abs(float(loss) - float(l_ref)) <= 1e-5 * max(1.0, abs(float(l_ref))), (B, float(loss), float(l_ref))
        assert (p.grad.cpu() - p_ref.grad).abs().max() <= 1e-5, B


def test_unknown_and_vector_valued_loss_names_are_refused():
    from rechorus_b200 import ops
    with pytest.raises(ValueError):
        ops.listwise_kind("BPRsimple")
    with pytest.raises(ValueError):
        ops.listwise_kind("hinge")


def test_bprmf_impression_model_trains_and_returns_the_vectors():
    """BPRMFImpression (BPRMF.py:65-80): forward returns prediction, u_v, i_v (BPRMF.py:43-45); loss(out, target) is the
    list-wise loss; gradients reach both tables and equal the oracle's"""
    from rechorus_b200 import ops, plugin
    p = plugin.BPRMFImpression.parse_model_args(argparse.ArgumentParser())
    a = p.parse_args(["--emb_size", "64", "--loss_n", "BPRhard", "--train_max_pos_item", "3", "--train_max_neg_item", "5"])
    a.device, a.model_path = torch.device("cuda", 0), "/tmp/_b2r_imp.pt"
    torch.manual_seed(4)
    m = plugin.BPRMFImpression(a, types.SimpleNamespace(n_users=40, n_items=70)).to(a.device)
    with torch.no_grad():
        for q in m.parameters():
            q.mul_(20.0)
    g = torch.Generator().manual_seed(5)
    B, Cn = 9, 8
    uid, iid = torch.randint(1, 40, (B,), generator=g), torch.randint(1, 70, (B, Cn), generator=g)
    target = torch.full((B, Cn), -1, dtype=torch.int64)
    for b in range(B):
        target[b, :1 + b % 3] = 1
        target[b, 3:4 + b % 5] = 0
    m.train()
    out = m({"user_id": uid.cuda(), "item_id": iid.cuda(), "batch_size": B, "phase": "train"})
    assert set(out) == {"prediction", "u_v", "i_v"} and tuple(out["u_v"].shape) == (B, Cn, 64) == tuple(out["i_v"].shape)
    w = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    assert torch.equal(out["i_v"].detach().cpu(), w["i_embeddings.weight"][iid])
    assert torch.equal(out["u_v"].detach().cpu()[:, 0], w["u_embeddings.weight"][uid])
    loss = m.loss(out, target.cuda())
    loss.backward()
    U, I = w["u_embeddings.weight"].clone().requires_grad_(True), w["i_embeddings.weight"].clone().requires_grad_(True)
    pred_ref = O.bprmf_scores({"u_embeddings.weight": U, "i_embeddings.weight": I}, uid, iid)
    l_ref = O.listwise_loss(pred_ref, target, "BPRhard", 3)
    l_ref.backward()
    assert abs(float(loss) - float(l_ref)) <= 1e-5
    assert (m.u_embeddings.weight.grad.cpu() - U.grad).abs().max() <= 1e-5
    assert (m.i_embeddings.weight.grad.cpu() - I.grad).abs().max() <= 1e-5
    ops.check_ids()
