"""Generate tests/golden/fit_*.npz by running the UNMODIFIED reference end to end on CPU: its own Dataset classes
(negative sampling, collate), its own BaseRunner.fit (DataLoader shuffle, candidate shuffle / un-shuffle, torch.optim
built by name over customize_parameters()) and BaseRunner.evaluate, for two epochs on the tiny corpus of
tests/golden/fit_corpus.py.  Run in the build container only:

    python tests/golden/make_fit_golden.py

Each .npz holds the initial and the final state dict ("w0:<key>", "w1:<key>"), the per-epoch mean loss and the dev
metrics after training -- what this repository's runner + kernels must reproduce under the same RNG seeds."""
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
    from models.general.NeuMF import NeuMF
    from models.sequential.SASRec import SASRec
    classes = {"BPRMF": BPRMF, "NeuMF": NeuMF, "SASRec": SASRec}
    for name, (cls_name, flags) in fit_corpus.CASES.items():
        cls = classes[cls_name]
        p = argparse.ArgumentParser()
        p = BaseRunner.parse_runner_args(p)
        p = cls.parse_model_args(p)
        a = p.parse_args(flags + fit_corpus.COMMON)
        a.device, a.model_path, a.log_file, a.train = torch.device("cpu"), "/tmp/_fit_golden.pt", "/tmp/_fit_golden.log", 1
        corpus = fit_corpus.build()
        torch.manual_seed(5)
        model = cls(a, corpus)
        model.apply(model.init_weights)
        data = {ph: cls.Dataset(model, corpus, ph) for ph in ("train", "dev")}
        for d in data.values():
            d.prepare()
        blob = {"w0:" + k: v.detach().numpy().copy() for k, v in model.state_dict().items()}
        runner = BaseRunner(a)
        np.random.seed(42)
        torch.manual_seed(42)
        losses = [runner.fit(data["train"], epoch=e + 1) for e in range(fit_corpus.EPOCHS)]
        metrics = runner.evaluate(data["dev"], [5, 10], ["NDCG", "HR"])
        for k, v in model.state_dict().items():
            blob["w1:" + k] = v.detach().numpy().copy()
        blob["losses"] = np.array(losses, dtype=np.float64)
        # The same two epochs in float64 (same initial weights, batches and permutations): how far the reference's OWN
        # fp32 result sits from the exact one.  Adam divides every gradient entry by its running magnitude, so entries
        # whose gradient is rounding-level noise take lr-sized steps in any fp32 implementation; "w1_64:" lets a test
        # bound another implementation's deviation by the reference's own rounding sensitivity instead of a guess.
        corpus64 = fit_corpus.build()
        torch.manual_seed(5)
        model64 = cls(a, corpus64)
        model64.apply(model64.init_weights)
        model64.double()
        data64 = cls.Dataset(model64, corpus64, "train")
        rand32 = torch.rand
        torch.rand = lambda *sz, **kw: rand32(*sz, dtype=torch.float32, **kw).double()   # same CPU RNG consumption
        torch.set_default_dtype(torch.float64)
        try:
            runner64 = BaseRunner(a)
            np.random.seed(42)
            torch.manual_seed(42)
            losses64 = [runner64.fit(data64, epoch=e + 1) for e in range(fit_corpus.EPOCHS)]
        finally:
            torch.set_default_dtype(torch.float32)
            torch.rand = rand32
        for k, v in model64.state_dict().items():
            blob["w1_64:" + k] = v.detach().numpy().astype(np.float64)
        blob["losses64"] = np.array(losses64, dtype=np.float64)
        for k, v in metrics.items():
            blob["m:" + k] = np.float64(v)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **blob)
        dd = np.concatenate([np.abs(blob["w1:" + k].astype(np.float64) - blob["w1_64:" + k]).ravel()
                             for k in model.state_dict()])
        print(name, "fp32-vs-fp64 reference: median %.3g q90 %.3g max %.3g" % (np.median(dd), np.quantile(dd, 0.9), dd.max()),
              "losses64", [round(x, 6) for x in losses64])
        print(name, "losses", [round(x, 6) for x in losses], {k: round(float(v), 4) for k, v in metrics.items()},
              os.path.getsize(path), "B")


if __name__ == "__main__":
    main()
