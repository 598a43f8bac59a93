"""CPU: the oracle's list-wise losses (oracle.listwise_loss) against fixtures produced by the unmodified reference's
ImpressionModel.loss (tests/golden/make_listwise_golden.py): value and autograd gradient, every loss name."""
import os

import numpy as np
import pytest
import torch

from oracle import rechorus_oracle as O

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "listwise.npz"))
LOSSES = ["BPR", "BPRhard", "BPRafter", "BPRhardafter", "BPRbefore", "BPRhardbefore", "listnet", "softmaxCE", "attention_rank"]


@pytest.mark.parametrize("case", ["a", "b", "c"])
@pytest.mark.parametrize("loss_n", LOSSES)
def test_oracle_listwise_equals_reference_fixture(case, loss_n):
    pred = torch.from_numpy(GOLD[f"{case}:pred"]).clone().requires_grad_(True)
    target = torch.from_numpy(GOLD[f"{case}:target"])
    loss = O.listwise_loss(pred, target, loss_n, int(GOLD[f"{case}:max_pos"]))
    loss.backward()
    assert abs(float(loss) - float(GOLD[f"{case}:{loss_n}:loss"])) <= 1e-5 * max(1.0, abs(float(loss)))
    assert np.abs(pred.grad.numpy() - GOLD[f"{case}:{loss_n}:grad"]).max() <= 1e-6
