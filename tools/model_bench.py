#!/usr/bin/env python
"""Informational timings of the parity-test configurations 3 (NeuMF) and 4 (SASRec) of BASELINE.json on one GPU:
training step = forward + BPR loss + backward (autograd nodes over the library kernels) + RowSparseOptimizer.step().
Prints one JSON line per model.  Not the bench.py contract (configs 3/4 are parity cases, not bench lines)."""
import argparse
import json
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def build(name, extra, n_users, n_items, dev):
    from rechorus_b200 import plugin
    from rechorus_b200.optim import RowSparseOptimizer
    from rechorus_b200.runner import BaseRunner
    cls = getattr(plugin, name)
    p = argparse.ArgumentParser()
    p = BaseRunner.parse_runner_args(p)
    p = cls.parse_model_args(p)
    a = p.parse_args(extra + ["--table_mode", "fused"])
    a.device, a.model_path = dev, "/tmp/_b2r_unused.pt"
    torch.manual_seed(0)
    m = cls(a, types.SimpleNamespace(n_users=n_users, n_items=n_items)).to(dev)
    m.optimizer = RowSparseOptimizer(m, "Adam", lr=1e-3)
    m.train()
    return m


def timeit(m, feeds, steps, warmup):
    def step(k):
        m.optimizer.zero_grad()
        loss = m.loss(m(feeds[k % len(feeds)]))
        loss.backward()
        m.optimizer.step()
        return loss
    for k in range(warmup):
        step(k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(steps):
        loss = step(k)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, float(loss)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(0)
    n = 1_000_000
    B = 4096
    # config 3: NeuMF d=64, layers [64,32,16], K=4
    m = build("NeuMF", ["--emb_size", "64", "--layers", "[64, 32, 16]", "--num_neg", "4"], n, n, dev)
    feeds = [{"user_id": torch.randint(1, n, (B,), generator=g).to(dev), "item_id": torch.randint(1, n, (B, 5), generator=g).to(dev),
              "batch_size": B, "phase": "train"} for _ in range(4)]
    ms, loss = timeit(m, feeds, a.steps, a.warmup)
    print(json.dumps({"config": "c3 NeuMF d=64 layers[64,32,16] K=4 B=4096 1M tables", "ms_per_step": ms,
                      "user_item_per_s": B * 5 / ms * 1e3, "loss": loss}), flush=True)
    del m
    torch.cuda.empty_cache()
    # config 4: SASRec L=50, 2 blocks, 4 heads, K=99
    m = build("SASRec", ["--emb_size", "64", "--history_max", "50", "--num_layers", "2", "--num_heads", "4", "--num_neg", "99"],
              10, n, dev)
    L = 50
    feeds = []
    for _ in range(4):
        lengths = torch.randint(1, L + 1, (B,), generator=g)
        lengths[0] = L
        hist = torch.randint(1, n, (B, L), generator=g) * (torch.arange(L).view(1, L) < lengths.view(B, 1))
        feeds.append({"user_id": torch.zeros(B, dtype=torch.int64).to(dev), "item_id": torch.randint(1, n, (B, 100), generator=g).to(dev),
                      "history_items": hist.to(dev), "lengths": lengths.to(dev), "batch_size": B, "phase": "train"})
    ms, loss = timeit(m, feeds, a.steps, a.warmup)
    print(json.dumps({"config": "c4 SASRec L=50 d=64 2 blocks 4 heads K=99 B=4096 1M items", "ms_per_step": ms,
                      "user_item_per_s": B * 100 / ms * 1e3, "loss": loss}), flush=True)


if __name__ == "__main__":
    main()
