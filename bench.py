#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the ReChorus training hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2]

A "step" is one pass of the hot path over one batch of synthetic input: forward (gather + dot), BPR loss,
backward, and the optimizer update -- the body of helpers/BaseRunner.py:193-206.  Default workload
(BASELINE.json configs[1], the one `metric` is quoted on): BPRMF emb_dim=64, 1 M synthetic users and items,
num_neg=99, batch=4096 on one B200.  metric = training samples/s counted as user x (1+neg) = B*C per step.

Prints ONE JSON line (rank 0).  `value`: ids already resident in HBM.  `e2e`: the same step through the
public plugin call (model.train_step(feed_dict)) with the batch coming from pinned host memory every step
and the loss read back every step.  `roofline`: dominant kernel, algorithmic bytes / CUDA-event time /
measured peak.  `cpu_baseline`: the oracle's reference-style CPU step on this box's host cores.
N > 1: configs 1-4 do not shard ("replicas only"): every rank trains an independent replica, no collective
on the data path, scaling = weak.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

WORKLOADS = {
    # name: (model, n_users, n_items, d, B, K)
    "c2": dict(model="BPRMF", n_users=1_000_000, n_items=1_000_000, d=64, B=4096, K=99,
               desc="BPRMF emb_dim=64, 1M synthetic items (1M users), num_neg=99, batch=4096"),
}
METRIC = "training samples/sec (user x (1+neg))"
UNIT = "user*item/s"
POOL = 8          # distinct pre-generated batches cycled through (fresh ids every step)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines, self.windows = index, None, [], []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def window(self, t0, t1):
        """mark [t0, t1] (time.time()) as 'GPU under the benchmark load'"""
        self.windows.append((t0, t1))

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()          # exact PID we started
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, ln in self.lines:
            if self.windows and not any(a <= ts <= b + 0.02 for a, b in self.windows):
                continue
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------
# synthetic data (seeded; ids uniform in [1, n) -- id 0 is the reference's padding id, data/README.md:13)
# ------------------------------------------------------------------------------------------------------

def make_batches(w, seed, pool=POOL):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(pool):
        uid = torch.randint(1, w["n_users"], (w["B"],), generator=g, dtype=torch.int64)
        iid = torch.randint(1, w["n_items"], (w["B"], w["K"] + 1), generator=g, dtype=torch.int64)
        out.append((uid, iid))
    return out


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------

def build_model(w, device):
    import types
    from rechorus_b200 import plugin
    from rechorus_b200.optim import RowSparseOptimizer
    from rechorus_b200.runner import BaseRunner
    p = argparse.ArgumentParser()
    p = BaseRunner.parse_runner_args(p)
    p = plugin.BPRMF.parse_model_args(p)
    a = p.parse_args(["--emb_size", str(w["d"]), "--num_neg", str(w["K"]), "--table_mode", "fused"])
    a.device, a.model_path = device, "/tmp/_b2r_bench.pt"
    torch.manual_seed(0)
    model = plugin.BPRMF(a, types.SimpleNamespace(n_users=w["n_users"], n_items=w["n_items"])).to(device)
    model.optimizer = RowSparseOptimizer(model, "Adam", lr=1e-3, l2=0.0)     # reference defaults (BaseRunner.py:28-38)
    model.train()
    return model


def run_ours(args, rank, world, local_rank):
    from rechorus_b200 import lib as L, ops
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    w = WORKLOADS[args.workload]
    B, C, d = w["B"], w["K"] + 1, w["d"]
    model = build_model(w, device)
    lib = L.load()
    host = make_batches(w, seed=1000 + rank)
    pinned = [(u.pin_memory(), i.pin_memory()) for u, i in host]
    dev_batches = [(u.to(device), i.to(device)) for u, i in host]
    n_uniq = [int(torch.unique(i).numel()) for _, i in host]
    n_uniq_u = [int(torch.unique(u).numel()) for u, _ in host]

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    feeds = [{"user_id": u, "item_id": i, "batch_size": B, "phase": "train"} for u, i in dev_batches]

    def step_resident(k):
        # the next batch is handed over too (as a prefetching data loader would): its index plan is built on
        # the library's side stream while this step's kernels run
        return model.train_step(feeds[k % POOL], feeds[(k + 1) % POOL])

    # ---- value: ids resident in HBM ------------------------------------------------------------------
    for k in range(args.warmup):
        step_resident(k)
    barrier()
    ops.check_ids(device)
    tags = {"score_fwd": L.PROF_SCORE_FWD, "score_bwd_query": L.PROF_SCORE_BWDQ, "segment_adam_items": L.PROF_SEGMENT_I,
            "segment_adam_users": L.PROF_SEGMENT_U, "plan_items": L.PROF_PLAN_I, "loss": L.PROF_LOSS}
    prof_every = max(16, args.steps // 64)           # bracket the tagged kernels on every prof_every-th step (>= 1 sample;
                                                      # sparse, so the brackets' events do not weigh on a short timed run)
    prof_steps = list(range(0, args.steps, prof_every))
    evs = {name: {k: (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                  for k in prof_steps} for name in tags}
    for name in tags:                      # events must exist (be created) before cudaEventRecord from C
        for a, b in evs[name].values():
            a.record(); b.record()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    # host cost of enqueuing one step with an empty launch queue (no back-pressure): 40 steps right after a sync
    torch.cuda.synchronize()
    h0 = time.perf_counter()
    for k in range(40):
        step_resident(k)
    host_cost_ms = (time.perf_counter() - h0) * 1e3 / 40
    torch.cuda.synchronize()
    sampler.start()
    time.sleep(0.3)                      # let nvidia-smi come up before the timed region
    launches0 = lib.b2r_launch_count()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    w0 = time.time()
    t0.record()
    for k in range(args.steps):
        if k % prof_every == 0:
            for name, tag in tags.items():
                a, b = evs[name][k]
                lib.b2r_profile_arm(tag, a.cuda_event, b.cuda_event)
        loss = step_resident(args.warmup + k)
    t1.record()
    host_ms_step = (time.time() - w0) * 1e3 / args.steps       # host time to enqueue a step (no sync inside the loop)
    barrier()
    sampler.window(w0, time.time())
    launches = lib.b2r_launch_count() - launches0
    ms_total = t0.elapsed_time(t1)
    if world > 1:
        t = torch.tensor([ms_total], device=device)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * B * C / (ms_step * 1e-3)
    for tag in tags.values():
        lib.b2r_profile_arm(tag, None, None)
    kern_ms = {name: statistics.mean(a.elapsed_time(b) for a, b in evs[name].values()) for name in tags}
    fused = kern_ms["score_bwd_query"] < 0.012       # the fused kernel replaces fwd + loss + bwd_query
    if fused:
        kern_ms["fused_score_loss_bwd"] = kern_ms.pop("score_fwd")
        kern_ms.pop("score_bwd_query")
        kern_ms.pop("loss")
    merged_apply = kern_ms.get("segment_adam_users", 1.0) < 0.009   # both tables updated by ONE launch (b2r_bucket_apply_pair)
    if merged_apply:
        kern_ms.pop("segment_adam_users")
    final_loss = float(loss.item())

    # ---- e2e: pinned host batch -> H2D -> step -> loss D2H, every step, through model.train_step -----
    # Triple-buffered device id buffers filled by a copy stream (the DataLoader's pin_memory/prefetch role);
    # the host reads every step's loss from pinned memory, a few steps behind the enqueue front (helpers/BaseRunner.py:207
    # reads it every step too).  All copies are enqueued inside the timed region.
    NB = 3
    copy_s = torch.cuda.Stream(device=device)
    main_s = torch.cuda.current_stream(device)
    bufs = [{"user_id": torch.empty(B, dtype=torch.int64, device=device),
             "item_id": torch.empty((B, C), dtype=torch.int64, device=device), "batch_size": B, "phase": "train"}
            for _ in range(NB)]
    ready = [torch.cuda.Event() for _ in range(NB)]
    done = [torch.cuda.Event() for _ in range(NB)]
    LAG, RING = 3, 4                      # the host reads every step's loss, LAG steps behind the enqueue front
    loss_h = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(RING)]
    loss_ev = [torch.cuda.Event() for _ in range(RING)]
    keep = [None] * RING                  # keeps each step's loss tensor alive until the host has read it
    d2h_s = torch.cuda.Stream(device=device)

    def stage(k):
        j = k % NB
        u, i = pinned[k % POOL]
        with torch.cuda.stream(copy_s):
            copy_s.wait_event(done[j])
            bufs[j]["user_id"].copy_(u, non_blocking=True)
            bufs[j]["item_id"].copy_(i, non_blocking=True)
            ready[j].record(copy_s)

    def run_e2e(n_steps, k0):
        for j in range(NB):
            done[j].record(main_s)
        stage(k0)
        stage(k0 + 1)
        seen = []
        for k in range(k0, k0 + n_steps):
            stage(k + 2)
            main_s.wait_event(ready[k % NB])
            main_s.wait_event(ready[(k + 1) % NB])
            ls = model.train_step(bufs[k % NB], bufs[(k + 1) % NB])
            done[k % NB].record(main_s)
            # the 4-byte loss read-back rides on its own stream: a D2H copy enqueued on the main stream would sit
            # between this step's last kernel and the next step's first one
            keep[k % RING] = ls
            with torch.cuda.stream(d2h_s):
                d2h_s.wait_event(done[k % NB])
                loss_h[k % RING].copy_(ls, non_blocking=True)
                loss_ev[k % RING].record(d2h_s)
            if k - LAG >= k0:
                loss_ev[(k - LAG) % RING].synchronize()
                seen.append(float(loss_h[(k - LAG) % RING]))
        for k in range(max(k0, k0 + n_steps - LAG), k0 + n_steps):
            loss_ev[k % RING].synchronize()
            seen.append(float(loss_h[k % RING]))
        copy_s.synchronize()
        return seen

    run_e2e(args.warmup, 0)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.time()
    e0.record()
    e2e_losses = run_e2e(args.steps, args.warmup)
    e1.record()
    barrier()
    sampler.window(w0, time.time())
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_e2e], device=device)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms_e2e = float(t.item())
    e2e_value = world * B * C / (ms_e2e / args.steps * 1e-3)
    ops.check_ids(device)
    # if the timed regions were too short for nvidia-smi to sample, keep the same loop running ~0.5 s more
    # (untimed) so the clock record reflects this workload; flagged in the output
    probe = False
    if len([1 for ts, _ in sampler.lines if any(a <= ts <= b for a, b in sampler.windows)]) < 3:
        probe = True
        w0 = time.time()
        k = 0
        while time.time() - w0 < 0.5:
            step_resident(k)
            k += 1
            if k % 50 == 0:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
        sampler.window(w0, time.time())
    clocks = sampler.stop()
    clocks["includes_untimed_continuation"] = probe

    # ---- roofline of the dominant kernel -------------------------------------------------------------
    peak, peak_src = measured_peaks()
    nu = statistics.mean(n_uniq[(args.warmup + k) % POOL] for k in range(args.steps))
    nuu = statistics.mean(n_uniq_u[(args.warmup + k) % POOL] for k in range(args.steps))
    n = B * C
    alg = {   # algorithmic bytes per launch (DESIGN.md section "kernels")
        "fused_score_loss_bwd": n * (4 * d + 8 + 4) + 2 * B * 4 * d + 12 * B,
        "score_fwd": n * (4 * d + 8 + 4) + B * 4 * d,
        "score_bwd_query": n * (4 * d + 8 + 4) + B * 4 * d,
        "segment_adam_items": nu * 6 * 4 * d + n * 16,
        "segment_adam_users": nuu * 6 * 4 * d + B * 4 * d + B * 12,
    }
    if merged_apply:
        alg["segment_adam_items"] += alg["segment_adam_users"]
    alg = {k_: v for k_, v in alg.items() if k_ in kern_ms}
    dom = max(alg, key=lambda k_: kern_ms[k_])
    achieved = alg[dom] / (kern_ms[dom] * 1e-3) / 1e9
    traffic = None                              # DRAM bytes per launch of that kernel from the committed ncu capture
    tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if os.path.exists(tpath) and args.workload == "c2":
        with open(tpath) as f:
            traffic = json.load(f).get(dom)
    cuda_names = {"segment_adam_items": "k_apply_sorted<16,2> (item + user table in one launch: segment reduce + Adam)"
                  if merged_apply else "k_apply_sorted<16,2> (item table: segment reduce + Adam)",
                  "fused_score_loss_bwd": "k_bprmf_fused<16,8,3>", "segment_adam_users": "k_apply_sorted<16,2> (user table)",
                  "score_fwd": "k_rowdot_fwd", "score_bwd_query": "k_rowdot_bwd_query"}
    roofline = {"bound": "hbm", "kernel": dom, "cuda_kernel": cuda_names.get(dom, dom), "achieved": round(achieved, 1),
                "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic,
                "peak_source": peak_src,
                "alg_bytes_per_launch": int(alg[dom]), "kernel_ms": round(kern_ms[dom], 5)}
    kernels = {k_: {"ms": round(v, 5), "GBps": round(alg[k_] / (v * 1e-3) / 1e9, 1) if k_ in alg else None}
               for k_, v in kern_ms.items()}
    # whole step against SURVEY 8(d)'s per-sample figure (fwd+bwd rows once, no optimizer) + the Adam rows
    survey_bytes = B * (2 * (C + 1) * 4 * d + 8 * (C + 1) + 4 * C)
    step_alg = survey_bytes - (nu + nuu) * 4 * d + (nu + nuu) * 6 * 4 * d   # grad-row write replaced by w,m,v rmw
    out = {
        "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_step, 5), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic (seeded uniform ids, N(0,0.01) tables)",
        "config": {"workload": args.workload + ": " + w["desc"], "optimizer": "Adam lr=1e-3 (row-sparse/lazy, fused)",
                   "parallelism": "replicas only" if world > 1 else "single GPU",
                   "l2_policy": "inputs larger than L2: 512 MB tables + 1 GB Adam state, fresh ids every step"},
        "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": 8 * B + 8 * B * C,
                "d2h_bytes_per_step": 4, "ms_per_step": round(ms_e2e / args.steps, 5)},
        "gpu_launches": int(launches), "host_enqueue_ms_per_step": round(host_ms_step, 5), "host_cost_ms_per_step_empty_queue": round(host_cost_ms, 5), "clocks": clocks, "roofline": roofline, "kernels": kernels,
        "step_roofline": {"alg_bytes_per_step": int(step_alg), "achieved": round(step_alg / (ms_step * 1e-3) / 1e9, 1),
                          "frac": round(step_alg / (ms_step * 1e-3) / 1e9 / peak, 4),
                          "survey_8d_bytes_no_optimizer": int(survey_bytes)},
        "final_loss": round(final_loss, 6),
    }
    return out


# ------------------------------------------------------------------------------------------------------
# CPU legs (the only place bench.py touches oracle/)
# ------------------------------------------------------------------------------------------------------

def _use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU legs are meant to use the box's cores (the physical
    ones: half of os.cpu_count() on an SMT host, torch's own default), and only rank 0 runs them."""
    want = max(1, (os.cpu_count() or 2) // 2)
    if torch.get_num_threads() < want:
        torch.set_num_threads(want)


def run_cpu_reference(w, steps, warmup, batch_B=None, seed=1000):
    """The reference's own CPU step (oracle port of helpers/BaseRunner.py:184-207 incl. dense Adam)."""
    from oracle import rechorus_oracle as O
    _use_all_host_threads()
    B = batch_B or w["B"]
    g = torch.Generator().manual_seed(0)
    params = O.bprmf_init(w["n_users"], w["n_items"], w["d"], g)
    trainer = O.ReferenceStyleTrainer(w["model"], params, lr=1e-3, l2=0.0, optimizer="Adam")
    batches = make_batches(dict(w, B=B), seed=seed, pool=max(2, min(POOL, steps + warmup)))
    for k in range(warmup):
        u, i = batches[k % len(batches)]
        trainer.step({"user_id": u, "item_id": i})
    t0 = time.perf_counter()
    for k in range(steps):
        u, i = batches[(warmup + k) % len(batches)]
        loss = trainer.step({"user_id": u, "item_id": i})
    dt = time.perf_counter() - t0
    C = w["K"] + 1
    return {"value": B * C * steps / dt, "ms_per_step": dt / steps * 1e3, "B": B, "loss": loss}


def cpu_baseline_block(w):
    r = run_cpu_reference(w, steps=10, warmup=1)
    return {"value": round(r["value"], 1), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "os_cpu_count": os.cpu_count(), "ms_per_step": round(r["ms_per_step"], 2),
            "sample": "10 timed + 1 warm-up full-size steps (B=4096, C=100) of the oracle's reference-style "
                      "CPU step: shuffle, forward, BPR loss, dense backward, dense torch.optim.Adam"}


def run_reference_arm(args, rank):
    if rank != 0:
        return None
    w = WORKLOADS[args.workload]
    # each reference step is one full batch (~0.7 s on this box's cores): cap the count so the arm ends in minutes
    steps, warmup = min(args.steps, 40), min(args.warmup, 2)
    r = run_cpu_reference(w, steps=steps, warmup=warmup)
    C = w["K"] + 1
    return {
        "impl": "reference", "metric": METRIC, "value": round(r["value"], 1), "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": round(r["ms_per_step"], 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic (seeded uniform ids, N(0,0.01) tables)",
        "config": {"workload": args.workload + ": " + w["desc"], "optimizer": "torch.optim.Adam dense (reference)",
                   "steps_requested": args.steps,
                   "parallelism": "CPU, rank 0 only"},
        "cpu_baseline": {"value": round(r["value"], 1), "unit": UNIT, "cores": torch.get_num_threads(),
                         "os_cpu_count": os.cpu_count(), "kind": "port",
                         "sample": f"each step = one full batch (B={r['B']}, C={C}) of the oracle's "
                                   "reference-style CPU step incl. dense Adam"},
        "e2e": {"value": round(r["value"], 1), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", type=str, default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", type=str, default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no_cpu_baseline", action="store_true", help="skip the ~15 s CPU leg (dev runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        out = run_reference_arm(args, rank)
        if out is not None:
            print(json.dumps(out), flush=True)
        return

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    out = run_ours(args, rank, world, local_rank)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline_block(WORKLOADS[args.workload])
        print(json.dumps(out), flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
