"""Generate tests/golden/listwise.npz by calling the UNMODIFIED reference's ImpressionModel.loss
(models/BaseImpressionModel.py:44-128) on seeded predictions / targets, with autograd gradients.  Build container only:

    python tests/golden/make_listwise_golden.py
"""
import os
import sys
import types

import numpy as np

REF_SRC = "/root/reference/src"
OUT = os.path.dirname(os.path.abspath(__file__))
LOSSES = ["BPR", "BPRhard", "BPRafter", "BPRhardafter", "BPRbefore", "BPRhardbefore", "listnet", "softmaxCE", "attention_rank"]


def cases():
    import torch
    g = torch.Generator().manual_seed(2024)
    out = {}
    for name, (B, max_pos, max_neg) in {"a": (6, 3, 5), "b": (17, 20, 20), "c": (4, 1, 1)}.items():
        Cn = max_pos + max_neg
        pred = torch.randn(B, Cn, generator=g) * 1.5
        target = torch.full((B, Cn), -1, dtype=torch.int64)
        for b in range(B):
            npos = int(torch.randint(1, max_pos + 1, (1,), generator=g))
            nneg = int(torch.randint(1, max_neg + 1, (1,), generator=g))
            target[b, :npos] = 1
            target[b, max_pos:max_pos + nneg] = 0
        out[name] = (pred, target, max_pos)
    return out


def main():
    np.object, np.int, np.float = object, int, float
    sys.path.insert(0, REF_SRC)
    import torch
    from models.BaseImpressionModel import ImpressionModel
    blob = {}
    for cname, (pred, target, max_pos) in cases().items():
        blob[f"{cname}:pred"], blob[f"{cname}:target"], blob[f"{cname}:max_pos"] = pred.numpy(), target.numpy(), np.int64(max_pos)
        for ln in LOSSES:
            holder = types.SimpleNamespace(loss_n=ln, train_max_pos_item=max_pos, device=torch.device("cpu"))
            p = pred.clone().requires_grad_(True)
            loss = ImpressionModel.loss(holder, {"prediction": p}, target)
            loss.backward()
            blob[f"{cname}:{ln}:loss"] = np.float64(loss.item())
            blob[f"{cname}:{ln}:grad"] = p.grad.numpy()
            print(cname, ln, float(loss))
    np.savez_compressed(os.path.join(OUT, "listwise.npz"), **blob)


if __name__ == "__main__":
    main()
