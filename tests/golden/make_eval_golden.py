"""Generate tests/golden/eval_test_all.npz by running the UNMODIFIED reference's evaluation under the test_all protocol
on CPU: models.general.BPRMF with --test_all 1, its Dataset (candidates = [target] + arange(1, n_items),
models/BaseModel.py:194-198), BaseRunner.predict (clicked-item masking, helpers/BaseRunner.py:244-251) and
evaluate_method, on the tiny corpus of tests/golden/fit_corpus.py.  Run in the build container only:

    python tests/golden/make_eval_golden.py

The fixture holds the weights, the masked prediction matrix, the integer ranks and the metrics: the pin for
oracle.test_all_predictions and, through it, for b2r_rank_all_items."""
import argparse
import os
import sys

import numpy as np

REF_SRC = "/root/reference/src"
OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, OUT)
import fit_corpus  # noqa: E402


def main():
    np.object = object
    np.int = int
    np.float = float
    sys.path.insert(0, REF_SRC)
    import torch
    from helpers.BaseRunner import BaseRunner
    from models.general.BPRMF import BPRMF
    p = argparse.ArgumentParser()
    p = BaseRunner.parse_runner_args(p)
    p = BPRMF.parse_model_args(p)
    a = p.parse_args(["--emb_size", "64", "--test_all", "1"] + fit_corpus.COMMON)
    a.device, a.model_path, a.log_file, a.train = torch.device("cpu"), "/tmp/_eval_golden.pt", "/tmp/_eval_golden.log", 1
    corpus = fit_corpus.build()
    torch.manual_seed(31)
    model = BPRMF(a, corpus)
    model.apply(model.init_weights)
    with torch.no_grad():
        for prm in model.parameters():
            prm.mul_(30.0)                       # trained-scale scores: rank gaps far above fp32 rounding
    dev = BPRMF.Dataset(model, corpus, "dev")
    dev.prepare()
    runner = BaseRunner(a)
    pred = runner.predict(dev)
    metrics = runner.evaluate_method(pred, [5, 10, 20], ["HR", "NDCG"])
    blob = {"w:" + k: v.detach().numpy().copy() for k, v in model.state_dict().items()}
    blob["pred"] = pred.astype(np.float32)
    blob["gt_rank"] = (pred >= pred[:, 0].reshape(-1, 1)).sum(axis=-1).astype(np.int64)
    blob["user_id"] = np.asarray(dev.data["user_id"], dtype=np.int64)
    blob["item_id"] = np.asarray(dev.data["item_id"], dtype=np.int64)
    for k, v in metrics.items():
        blob["m:" + k] = np.float64(v)
    path = os.path.join(OUT, "eval_test_all.npz")
    np.savez_compressed(path, **blob)
    print("pred", pred.shape, "ranks", blob["gt_rank"][:8], {k: round(float(v), 4) for k, v in metrics.items()},
          os.path.getsize(path), "B")


if __name__ == "__main__":
    main()
