#!/usr/bin/env python
"""Times the evaluation kernels on one GPU (not part of bench.py's contract):
  * test_all ranking of B users against an n_items x d table: ops.rank_all_items (scores never materialised) vs the
    reference-shaped route (torch matmul -> [B, n_items] scores -> mask -> compare -> sum), both on the device;
  * ranks of a [N, 100] prediction matrix: ops.gt_rank vs the torch expression of BaseRunner.py:63.
Prints one JSON line per case."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=256)
    ap.add_argument("--n_items", type=int, default=1_000_000)
    ap.add_argument("--emb", type=int, default=64)
    a = ap.parse_args()
    from rechorus_b200 import ops
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(0)
    q = torch.randn(a.B, a.emb, device=dev, generator=g)
    table = torch.randn(a.n_items, a.emb, device=dev, generator=g)
    target = torch.randint(1, a.n_items, (a.B,), device=dev, generator=g)
    n_mask = 200
    mrow = torch.arange(a.B, device=dev).repeat_interleave(n_mask)
    mitem = torch.randint(1, a.n_items, (a.B * n_mask,), device=dev, generator=g)

    def ours():
        return ops.rank_all_items(q, table, target, mrow, mitem)

    def torch_route():
        s = q @ table.t()                                  # [B, n_items]; column j is item j
        s0 = (q * table[target]).sum(-1, keepdim=True)
        s[mrow, mitem] = float("-inf")
        return 1 + (s[:, 1:] >= s0).sum(-1)

    ms_o, r_o = timed(ours)
    ms_t, r_t = timed(torch_route)
    flops = 2.0 * a.B * a.n_items * a.emb
    print(json.dumps({"case": "test_all rank", "B": a.B, "n_items": a.n_items, "d": a.emb, "ms_ours": round(ms_o, 4),
                      "TFLOPs_ours": round(flops / ms_o / 1e9, 2), "table_GBps_ours": round(a.n_items * a.emb * 4 / ms_o / 1e6, 1),
                      "ms_torch_materialised": round(ms_t, 4), "rank_mismatches_vs_torch(tf32/rounding)": int((r_o != r_t).sum())}))
    pred = torch.randn(14681, 100, device=dev, generator=g)
    ms_o, r_o = timed(lambda: ops.gt_rank(pred), 50)
    ms_t, r_t = timed(lambda: (pred >= pred[:, :1]).sum(-1), 50)
    print(json.dumps({"case": "gt_rank [14681,100]", "ms_ours": round(ms_o, 4), "ms_torch": round(ms_t, 4),
                      "equal": bool(torch.equal(r_o, r_t))}))


if __name__ == "__main__":
    main()
