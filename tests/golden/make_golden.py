"""Generate tests/golden/*.npz by executing the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

The reference ships no golden vectors of its own, so these fixtures -- outputs of
the reference's own classes (models.general.BPRMF.BPRMF, models.general.NeuMF.NeuMF,
models.sequential.SASRec.SASRec, GeneralModel.loss, BaseRunner.evaluate_method and the BaseRunner.fit
shuffle/unshuffle) on seeded inputs -- are the pin for oracle/rechorus_oracle.py and for the CUDA path.
Each .npz holds: the state dict ("w:<key>"), the batch ("in:<key>"), the prediction, the loss and the
dense gradient of every parameter ("g:<key>"), all float32 / int64 exactly as the reference produced them.
"""
import argparse
import os
import sys
import types

import numpy as np

REF_SRC = "/root/reference/src"
OUT = os.path.dirname(os.path.abspath(__file__))


def _import_reference():
    # NumPy >= 1.24 removed the aliases the reference still uses (BaseModel.py:141, SASRec.py:69, utils.py:65)
    np.object = object
    np.int = int
    np.float = float
    sys.path.insert(0, REF_SRC)
    import torch  # noqa: F401
    from models.general.BPRMF import BPRMF
    from models.general.NeuMF import NeuMF
    from models.sequential.SASRec import SASRec
    from helpers.BaseRunner import BaseRunner
    return BPRMF, NeuMF, SASRec, BaseRunner


def _args(model_cls, runner_cls, extra):
    import torch
    parser = argparse.ArgumentParser()
    parser = runner_cls.parse_runner_args(parser)
    parser = model_cls.parse_model_args(parser)
    args = parser.parse_args(extra)
    args.device = torch.device("cpu")
    args.model_path = "/tmp/_golden_unused.pt"
    return args


def _dump(name, model, batch, scale, extra_meta):
    import torch
    torch.manual_seed(1234)
    if scale != 1.0:
        with torch.no_grad():
            for n, p in model.named_parameters():
                if "layer_norm" not in n:
                    p.mul_(scale)
    model.train()
    model.zero_grad()
    out = model(dict(batch))
    loss = model.loss(out)
    loss.backward()
    blob = {"meta": np.array(repr(extra_meta))}
    for k, v in model.state_dict().items():
        blob["w:" + k] = v.detach().numpy().copy()
    for k, v in batch.items():
        if isinstance(v, torch.Tensor):
            blob["in:" + k] = v.numpy().copy()
    blob["prediction"] = out["prediction"].detach().numpy().copy()
    blob["loss"] = loss.detach().numpy().copy()
    for k, p in model.named_parameters():
        blob["g:" + k] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy().copy()
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **blob)
    print(f"{name}: pred {blob['prediction'].shape} loss {float(blob['loss']):.6f} -> {os.path.getsize(path)} B")


def main():
    import torch
    BPRMF, NeuMF, SASRec, BaseRunner = _import_reference()
    g = torch.Generator().manual_seed(20260923)

    def ids(lo, hi, shape):
        return torch.randint(lo, hi, shape, generator=g, dtype=torch.int64)

    # ---------------- BPRMF (BPRMF.py:34-45 + BaseModel.py:175-189) ----------------
    for tag, (nu, ni, d, B, C, scale) in {
        "bprmf_k1": (37, 53, 64, 16, 2, 1.0),          # config-1 shape: one negative -> plain BPR
        "bprmf_k9_trained": (37, 53, 64, 12, 10, 20.0),  # trained-scale weights, duplicates guaranteed (53 items)
        "bprmf_d128": (19, 211, 128, 7, 33, 20.0),       # config-5 row width, ragged sizes
        "bprmf_d20": (11, 29, 20, 5, 4, 20.0),           # emb_size not a multiple of 16
    }.items():
        torch.manual_seed(7)
        corpus = types.SimpleNamespace(n_users=nu, n_items=ni)
        model = BPRMF(_args(BPRMF, BaseRunner, ["--emb_size", str(d), "--num_neg", str(C - 1)]), corpus)
        batch = {"user_id": ids(1, nu, (B,)), "item_id": ids(1, ni, (B, C)), "batch_size": B, "phase": "train"}
        _dump(tag, model, batch, scale, dict(model="BPRMF", n_users=nu, n_items=ni, d=d, B=B, C=C, scale=scale))

    # ---------------- NeuMF (NeuMF.py:56-76) ----------------
    for tag, (nu, ni, d, layers, B, C, scale) in {
        "neumf_l64": (31, 47, 64, [64], 9, 2, 20.0),                 # reference default --layers '[64]'
        "neumf_l64_32_16": (31, 47, 64, [64, 32, 16], 10, 5, 20.0),  # config 3
        "neumf_d32": (13, 17, 32, [48, 8], 6, 3, 20.0),
    }.items():
        torch.manual_seed(11)
        corpus = types.SimpleNamespace(n_users=nu, n_items=ni)
        model = NeuMF(_args(NeuMF, BaseRunner, ["--emb_size", str(d), "--layers", str(layers),
                                                 "--num_neg", str(C - 1)]), corpus)
        batch = {"user_id": ids(1, nu, (B,)), "item_id": ids(1, ni, (B, C)), "batch_size": B, "phase": "train"}
        _dump(tag, model, batch, scale, dict(model="NeuMF", n_users=nu, n_items=ni, d=d, layers=layers,
                                             B=B, C=C, scale=scale))

    # ---------------- SASRec (SASRec.py:51-86, layers.py:34-63,112-118) ----------------
    for tag, (ni, d, L, nl, nh, B, C, scale) in {
        "sasrec_l1h1": (61, 64, 20, 1, 1, 6, 4, 10.0),     # demo script setting (Topk_Amazon.sh:26)
        "sasrec_l2h4": (61, 64, 50, 2, 4, 8, 10, 10.0),    # config 4
        "sasrec_d32": (43, 32, 12, 2, 2, 5, 3, 10.0),
    }.items():
        torch.manual_seed(13)
        corpus = types.SimpleNamespace(n_users=10, n_items=ni)
        model = SASRec(_args(SASRec, BaseRunner, ["--emb_size", str(d), "--history_max", str(L),
                                                   "--num_layers", str(nl), "--num_heads", str(nh),
                                                   "--num_neg", str(C - 1)]), corpus)
        lengths = ids(1, L + 1, (B,))
        lengths[0] = L                       # batch max length = history_max
        lengths[1] = 1                       # shortest possible history
        Lb = int(lengths.max())
        hist = ids(1, ni, (B, Lb))
        hist = hist * (torch.arange(Lb).view(1, Lb) < lengths.view(B, 1))   # right-padded with 0 (collate)
        batch = {"user_id": ids(1, 10, (B,)), "item_id": ids(1, ni, (B, C)), "history_items": hist,
                 "history_times": torch.zeros_like(hist), "lengths": lengths, "batch_size": B, "phase": "train"}
        _dump(tag, model, batch, scale, dict(model="SASRec", n_items=ni, d=d, history_max=L, num_layers=nl,
                                             num_heads=nh, B=B, C=C, scale=scale))

    # ---------------- runner arithmetic (BaseRunner.py:52-78 and :187-202) ----------------
    torch.manual_seed(17)
    pred = torch.randn(64, 100, generator=g).numpy().astype(np.float32)
    pred[3, 5] = pred[3, 0]                  # a tie counts against the positive
    pred[4, :] = 0.25                        # all-equal row -> rank 100
    res = BaseRunner.evaluate_method(pred, [1, 5, 10, 50], ["HR", "NDCG"])
    gt_rank = (pred >= pred[:, 0].reshape(-1, 1)).sum(axis=-1)
    blob = {"pred": pred, "gt_rank": gt_rank.astype(np.int64)}
    for k, v in res.items():
        blob["m:" + k] = np.float64(v)
    # the fit() shuffle / unshuffle round trip with a fixed CPU RNG state
    torch.manual_seed(99)
    item_ids = ids(1, 1000, (8, 6))
    indices = torch.argsort(torch.rand(*item_ids.shape), dim=-1)
    shuffled = item_ids[torch.arange(8).unsqueeze(-1), indices]
    scores = torch.randn(8, 6, generator=g)
    restored = torch.zeros(8, 6)
    restored[torch.arange(8).unsqueeze(-1), indices] = scores
    blob.update({"sh:item_id": item_ids.numpy(), "sh:indices": indices.numpy(), "sh:shuffled": shuffled.numpy(),
                 "sh:scores": scores.numpy(), "sh:restored": restored.numpy()})
    np.savez_compressed(os.path.join(OUT, "runner_metrics.npz"), **blob)
    print("runner_metrics:", {k: round(v, 4) for k, v in res.items()})


if __name__ == "__main__":
    main()
