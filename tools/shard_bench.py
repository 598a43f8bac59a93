#!/usr/bin/env python
"""Config-5 style run of the row-sharded BPRMF path under torchrun (one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/shard_bench.py --n_items 100000000 --n_users 1000000 --d 128 --B 4096 --K 255 --steps 20 --warmup 5 [--check]

--check: small sizes; every rank verifies scores against a replicated full-table computation (all-gathered shards).
Prints one JSON line on rank 0: user x item / s over all ranks (max-over-ranks CUDA-event time).
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n_items", type=int, default=100_000_000)
    ap.add_argument("--n_users", type=int, default=1_000_000)
    ap.add_argument("--emb", type=int, default=128)
    ap.add_argument("--B", type=int, default=4096)
    ap.add_argument("--K", type=int, default=255)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--optimizer", type=str, default="Adam")
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--torch_profile", action="store_true", help="print a torch.profiler kernel table for 3 steps (rank 0)")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from rechorus_b200 import ops
    from rechorus_b200.shard import ShardedBPRMF
    m = ShardedBPRMF(a.n_users, a.n_items, a.emb, dev, optimizer=a.optimizer, lr=1e-3, init_std=0.1 if a.check else 0.01)
    C = a.K + 1
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    pool = [(torch.randint(1, a.n_users, (a.B,), device=dev, generator=g),
             torch.randint(1, a.n_items, (a.B, C), device=dev, generator=g)) for _ in range(4)]
    if a.check:
        uid, iid = pool[0]
        pred, _ = m.scores(uid, iid)
        if world > 1:
            U = torch.empty(world * m.rows_u, a.emb, device=dev); dist.all_gather_into_tensor(U, m.U)
            I = torch.empty(world * m.rows_i, a.emb, device=dev); dist.all_gather_into_tensor(I, m.I)
        else:
            U, I = m.U, m.I
        ref = torch.einsum("bd,bcd->bc", U[uid], I[iid])
        err = float((pred - ref).abs().max())
        Ib = I.clone()
        loss = m.train_step(uid, iid)
        # item-shard update check (SGD-free: just that touched rows moved and untouched did not, on this rank's shard)
        lo = rank * m.rows_i
        moved = (m.I != Ib[lo:lo + m.rows_i]).any(1)
        all_ids = iid.reshape(-1)
        if world > 1:
            allg = torch.empty(world * all_ids.numel(), dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(allg, all_ids)
        else:
            allg = all_ids
        touched = torch.zeros(m.rows_i, dtype=torch.bool, device=dev)
        mine = allg[(allg >= lo) & (allg < lo + m.rows_i)] - lo
        touched[mine] = True
        ok = bool(torch.equal(moved, touched))
        ops.check_ids(dev)
        print(json.dumps({"rank": rank, "check_max_abs_err": err, "touched_rows_match": ok, "loss": float(loss)}), flush=True)
        assert err <= 1e-5 and ok
    for k in range(a.warmup):
        m.train_step(*pool[k % 4])
    torch.cuda.synchronize()
    if a.torch_profile:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            for k in range(3):
                m.train_step(*pool[k % 4])
            torch.cuda.synchronize()
        if rank == 0:
            print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70), flush=True)
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(a.steps):
        loss = m.train_step(*pool[k % 4])
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item()) / a.steps
    ops.check_ids(dev)
    if rank == 0:
        print(json.dumps({"metric": "training samples/sec (user x (1+neg))", "value": world * a.B * C / (ms * 1e-3),
                          "unit": "user*item/s", "n_gpus": world, "ms_per_step": ms, "scaling": "weak",
                          "config": {"workload": f"sharded BPRMF d={a.emb} n_items={a.n_items} n_users={a.n_users} "
                                                 f"B={a.B}/GPU K={a.K}", "parallelism": f"row-range shards x{world}, score routing",
                                     "optimizer": a.optimizer + " (row-sparse)"},
                          "loss": float(loss), "mem_GB": torch.cuda.max_memory_allocated() / 1e9}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
