"""GPU (one device, world = 1): ShardedBPRMF on the CUDA backend must equal the oracle's single-table step; the
multi-rank wiring is covered by tests/test_shard_gloo.py on CPU and by tools/shard_bench.py under torchrun."""
import pytest
import torch

from oracle import rechorus_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("d,C", [(64, 10), (128, 33)])
def test_single_rank_sharded_step_matches_oracle(d, C):
    from rechorus_b200 import ops
    from rechorus_b200.shard import ShardedBPRMF
    dev = torch.device("cuda", 0)
    n_users, n_items, B = 300, 500, 64
    m = ShardedBPRMF(n_users, n_items, d, dev, optimizer="SGD", lr=0.5, init_std=0.3)
    g = torch.Generator().manual_seed(5)
    U, I = m.U.detach().cpu().clone().requires_grad_(True), m.I.detach().cpu().clone().requires_grad_(True)
    for _ in range(2):
        uid = torch.randint(0, n_users, (B,), generator=g)
        iid = torch.randint(0, n_items, (B, C), generator=g)
        pred, _ = m.scores(uid.to(dev), iid.to(dev))
        ref = O.bprmf_scores({"u_embeddings.weight": U, "i_embeddings.weight": I}, uid, iid)
        assert (pred.cpu() - ref.detach()).abs().max() <= 1e-5
        loss = m.train_step(uid.to(dev), iid.to(dev))
        lref = O.bpr_loss(ref)
        assert abs(float(loss) - float(lref)) <= 1e-5
        gU, gI = torch.autograd.grad(lref, [U, I])
        with torch.no_grad():
            U -= 0.5 * gU
            I -= 0.5 * gI
        assert (m.U.cpu() - U.detach()).abs().max() <= 1e-5
        assert (m.I.cpu() - I.detach()).abs().max() <= 1e-5
    ops.check_ids()
