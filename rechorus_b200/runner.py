"""Stand-alone mirror of the reference's runner surface (helpers/BaseRunner.py) for use without the
reference checkout: same flags, same ``train / fit / evaluate / predict / evaluate_method / print_res``
semantics, same CPU RNG stream for the per-batch candidate shuffle.  Additions (all opt-in flags):
``--fused_optimizer 1`` installs ``RowSparseOptimizer`` through the ``model.optimizer`` seam
(BaseRunner.py:176-177) and switches the embedding tables to 'fused' gradient mode.
"""
from __future__ import annotations

import gc
import logging
import os
from time import time
from typing import Dict, List

import numpy as np
import torch
from torch.utils.data import DataLoader

from . import ops
from .optim import RowSparseOptimizer


def batch_to_device(batch: dict, device) -> dict:
    # utils/utils.py:30-34
    for k, v in batch.items():
        if isinstance(v, torch.Tensor):
            batch[k] = v.to(device)
    return batch


def format_metric(result: Dict[str, float]) -> str:
    # utils/utils.py:54-69: "HR@5:0.1234,NDCG@5:0.0567" sorted by k then metric name
    keys = sorted(result, key=lambda s: (int(s.split("@")[1]) if "@" in s else 0, s.split("@")[0]))
    parts = []
    for k in keys:
        v = result[k]
        parts.append(f"{k}:{v:<.4f}" if isinstance(v, (float, np.floating)) else f"{k}:{v}")
    return ",".join(parts)


def _poll_ids(model) -> None:
    """One synchronising poll of the kernels' out-of-range-id counter per epoch / evaluation: the reference's
    nn.Embedding raises IndexError on a bad id (BPRMF.py:39-40); the kernels clamp and count, so the runner raises."""
    dev = getattr(model, "device", None)
    if isinstance(dev, torch.device) and dev.type == "cuda":
        ops.check_ids(dev if dev.index is not None else None)


class BaseRunner:
    @staticmethod
    def parse_runner_args(parser):
        # helpers/BaseRunner.py:20-49 (same names, defaults and help strings' meaning)
        parser.add_argument("--epoch", type=int, default=200, help="Number of epochs.")
        parser.add_argument("--check_epoch", type=int, default=1, help="Check some tensors every check_epoch.")
        parser.add_argument("--test_epoch", type=int, default=-1, help="Print test results every test_epoch (-1: never).")
        parser.add_argument("--early_stop", type=int, default=10, help="Epochs of continuous dev drop before stopping.")
        parser.add_argument("--lr", type=float, default=1e-3, help="Learning rate.")
        parser.add_argument("--l2", type=float, default=0, help="Weight decay in optimizer.")
        parser.add_argument("--batch_size", type=int, default=256, help="Batch size during training.")
        parser.add_argument("--eval_batch_size", type=int, default=256, help="Batch size during testing.")
        parser.add_argument("--optimizer", type=str, default="Adam", help="optimizer: SGD, Adam, Adagrad, Adadelta")
        parser.add_argument("--num_workers", type=int, default=5, help="DataLoader worker processes.")
        parser.add_argument("--pin_memory", type=int, default=0, help="pin_memory in DataLoader")
        parser.add_argument("--topk", type=str, default="5,10,20,50", help="Cut-offs of the ranking metrics.")
        parser.add_argument("--metric", type=str, default="NDCG,HR", help="metrics: NDCG, HR")
        parser.add_argument("--main_metric", type=str, default="", help="Metric that selects the best model.")
        parser.add_argument("--fused_optimizer", type=int, default=0,
                            help="1: row-sparse fused SGD/Adam/Adagrad (rechorus_b200.optim) instead of torch.optim")
        parser.add_argument("--fused_step", type=int, default=0,
                            help="1 (with --fused_optimizer 1): models that offer train_step run every training step "
                                 "as one C call (forward, loss, backward, optimizer; next batch's plan prefetched)")
        parser.add_argument("--graph_step", type=int, default=0,
                            help="1 (with --fused_optimizer 1): capture the loop body of fit (zero_grad, forward, loss, "
                                 "backward, optimizer.step) once in a CUDA graph and replay it per full-size batch "
                                 "(rechorus_b200.graph.GraphedStep); the ragged last batch runs eagerly")
        parser.add_argument("--exact_adam", type=int, default=0,
                            help="1 (with --fused_optimizer 1, Adam): row-sparse cost, dense torch.optim.Adam results "
                                 "(rows are advanced through their skipped steps)")
        parser.add_argument("--device_sampler", type=int, default=0,
                            help="non-zero seed: draw the training negatives on the GPU every epoch (same distribution "
                                 "as BaseModel.py:206-214, counter-based stream)")
        parser.add_argument("--device_batches", type=int, default=0,
                            help="non-zero (a value > 1 is the seed): the whole epoch's batch production on the GPU -- "
                                 "negatives, row shuffle and collate as device kernels (runner.fit_on_device; GeneralModel)")
        parser.add_argument("--device_metrics", type=int, default=0,
                            help="1: rank the ground truth on the GPU (model.eval_ranks) instead of copying predictions "
                                 "to the host for evaluate_method")
        return parser

    @staticmethod
    def evaluate_method(predictions: np.ndarray, topk: list, metrics: list) -> Dict[str, float]:
        """helpers/BaseRunner.py:52-78.  Column 0 holds the ground-truth item's score; its rank is the number
        of candidates scoring >= it (ties count against it)."""
        gt_rank = (predictions >= predictions[:, 0].reshape(-1, 1)).sum(axis=-1)
        out = {}
        for k in topk:
            hit = gt_rank <= k
            for metric in metrics:
                key = f"{metric}@{k}"
                if metric == "HR":
                    out[key] = hit.mean()
                elif metric == "NDCG":
                    out[key] = (hit / np.log2(gt_rank + 1)).mean()
                else:
                    raise ValueError(f"Undefined evaluation metric: {metric}.")
        return out

    def __init__(self, args):
        self.train_models = getattr(args, "train", 1)
        self.epoch = args.epoch
        self.check_epoch = args.check_epoch
        self.test_epoch = args.test_epoch
        self.early_stop = args.early_stop
        self.learning_rate = args.lr
        self.batch_size = args.batch_size
        self.eval_batch_size = args.eval_batch_size
        self.l2 = args.l2
        self.optimizer_name = args.optimizer
        self.num_workers = args.num_workers
        self.pin_memory = args.pin_memory
        self.fused_optimizer = getattr(args, "fused_optimizer", 0)
        self.device_metrics = getattr(args, "device_metrics", 0)
        self.fused_step = getattr(args, "fused_step", 0)
        self.exact_adam = getattr(args, "exact_adam", 0)
        self.graph_step = getattr(args, "graph_step", 0)
        if self.graph_step and (self.exact_adam or not self.fused_optimizer):
            raise ValueError("--graph_step 1 needs --fused_optimizer 1 and is not available with --exact_adam 1 "
                             "(the exact mode keeps per-step host bookkeeping)")
        self.device_sampler = getattr(args, "device_sampler", 0)
        self.device_batches = getattr(args, "device_batches", 0)
        self.topk = [int(x) for x in args.topk.split(",")]
        self.metrics = [m.strip().upper() for m in args.metric.split(",")]
        self.main_metric = args.main_metric or f"{self.metrics[0]}@{self.topk[0]}"
        self.main_topk = int(self.main_metric.split("@")[1]) if "@" in self.main_metric else 0
        self.time = None
        log_file = getattr(args, "log_file", "") or ""
        self.log_path = os.path.dirname(log_file)
        self.save_appendix = log_file.split("/")[-1].split(".")[0]

    def _check_time(self, start=False):
        if self.time is None or start:
            self.time = [time()] * 2
            return self.time[0]
        prev = self.time[1]
        self.time[1] = time()
        return self.time[1] - prev

    def _build_optimizer(self, model):
        logging.info("Optimizer: " + self.optimizer_name)
        if self.fused_optimizer:
            model.set_table_mode("fused")
            return RowSparseOptimizer(model, self.optimizer_name, lr=self.learning_rate, l2=self.l2,
                                      exact_dense=bool(self.exact_adam), device_clock=bool(self.graph_step))
        cls = getattr(torch.optim, self.optimizer_name)       # helpers/BaseRunner.py:112
        return cls(model.customize_parameters(), lr=self.learning_rate, weight_decay=self.l2)

    # ---------------------------------------------------------------------------------------------
    def train(self, data_dict):
        model = data_dict["train"].model
        main_results, dev_results = [], []
        self._check_time(start=True)
        try:
            for epoch in range(self.epoch):
                self._check_time()
                gc.collect()
                loss = self.fit(data_dict["train"], epoch=epoch + 1)
                if np.isnan(loss):
                    logging.info("Loss is Nan. Stop training at %d." % (epoch + 1))
                    break
                train_t = self._check_time()
                dev = self.evaluate(data_dict["dev"], [self.main_topk], self.metrics)
                dev_results.append(dev)
                main_results.append(dev[self.main_metric])
                line = "Epoch {:<5} loss={:<.4f} [{:<3.1f} s]	dev=({})".format(epoch + 1, loss, train_t,
                                                                               format_metric(dev))
                if self.test_epoch > 0 and epoch % self.test_epoch == 0:
                    test = self.evaluate(data_dict["test"], self.topk[:1], self.metrics)
                    line += " test=({})".format(format_metric(test))
                line += " [{:<.1f} s]".format(self._check_time())
                if max(main_results) == main_results[-1] or (hasattr(model, "stage") and model.stage == 1):
                    model.save_model()
                    line += " *"
                logging.info(line)
                if self.early_stop > 0 and self.eval_termination(main_results):
                    logging.info("Early stop at %d based on dev result." % (epoch + 1))
                    break
        except KeyboardInterrupt:
            logging.info("Early stop manually")
        best = main_results.index(max(main_results))
        logging.info(os.linesep + "Best Iter(dev)={:>5}\t dev=({}) [{:<.1f} s] ".format(
            best + 1, format_metric(dev_results[best]), self.time[1] - self.time[0]))
        model.load_model()

    def _fit_whole_steps(self, dataset, epoch=-1) -> float:
        """The same epoch with the loop body of helpers/BaseRunner.py:185-207 replaced by the model's single-call
        ``train_step`` (forward + loss + backward + row-sparse optimizer in one C call, the next batch's index plan
        prefetched).  Batches and RNG consumption are those of ``fit``: the per-row candidate permutation is still
        drawn (so the torch RNG stream stays in step with the reference loop) but not applied -- it is a no-op for a
        model whose candidates are scored independently, and the whole-step kernel wants the positive in column 0.
        Losses stay on the device until the epoch ends (the reference syncs on every step)."""
        model = dataset.model
        dl = DataLoader(dataset, batch_size=self.batch_size, shuffle=True, num_workers=self.num_workers,
                        collate_fn=dataset.collate_batch, pin_memory=bool(self.pin_memory))
        it = iter(dl)
        cur = next(it, None)
        cur = batch_to_device(cur, model.device) if cur is not None else None
        losses = []
        ops.bprmf_step_reset()                 # no plan prefetched by an earlier (possibly aborted) epoch survives
        try:
            while cur is not None:
                nxt = next(it, None)
                if nxt is not None:
                    nxt = batch_to_device(nxt, model.device)
                torch.rand(*cur["item_id"].shape)                     # BaseRunner.py:189's draw, see above
                same_shape = nxt is not None and nxt["item_id"].shape == cur["item_id"].shape
                losses.append(model.train_step(cur, nxt if same_shape else None))
                cur = nxt
        finally:
            ops.bprmf_step_reset()
        out = float(np.mean(torch.stack(losses).cpu().numpy())) if losses else float("nan")
        _poll_ids(model)
        return out

    def _fit_graphed(self, dataset, epoch=-1) -> float:
        """The same epoch with the loop body of helpers/BaseRunner.py:193-206 captured once in a CUDA graph
        (rechorus_b200.graph.GraphedStep) and replayed for every batch of the captured shape: one launch per step instead of
        ~70 (NeuMF) to ~150 (SASRec).  The first full-size batch trains through the capture's eager warm-up step; the ragged
        last batch runs eagerly.  As in ``_fit_whole_steps`` the per-row candidate permutation of :189 is drawn (RNG stream in
        step with the reference) but not applied: it is a no-op for candidates scored independently.  Models with dropout keep
        the eager loop (``fit`` does not route them here)."""
        from .graph import GraphedStep
        model = dataset.model
        dl = DataLoader(dataset, batch_size=self.batch_size, shuffle=True, num_workers=self.num_workers,
                        collate_fn=dataset.collate_batch, pin_memory=bool(self.pin_memory))
        losses = []
        for batch in dl:
            batch = batch_to_device(batch, model.device)
            torch.rand(*batch["item_id"].shape)                       # BaseRunner.py:189's draw
            gs = model.__dict__.get("_b2r_graphed_step")
            if gs is None and batch["item_id"].shape[0] == self.batch_size:
                gs = GraphedStep(model, batch, warmup=1)              # its warm-up step IS this batch's training step
                model.__dict__["_b2r_graphed_step"] = gs
                losses.append(gs.warmup_losses[0])
            elif gs is not None and gs.matches(batch):
                losses.append(gs(batch).clone())
            else:
                model.optimizer.zero_grad()
                loss = model.loss(model(batch))
                loss.backward()
                model.optimizer.step()
                losses.append(loss.detach())
        model.optimizer.sync_clock()                                  # the host step counter follows the device clock
        out = float(np.mean(torch.stack([l.reshape(()) for l in losses]).cpu().numpy())) if losses else float("nan")
        _poll_ids(model)
        return out

    def fit_on_device(self, dataset, epoch=-1) -> float:
        """the same epoch with BATCH PRODUCTION on the GPU (``--device_batches <seed>``).  What the
        reference does on the host -- draw the negatives (models/BaseModel.py:206-214), shuffle the rows (the
        DataLoader's RandomSampler), build one feed dict per sample and collate them (BaseModel.py:192-203,135-152) --
        becomes three device operations: ``b2r_sample_negatives`` (one Philox draw + one binary search per negative, the
        reference's distribution from a counter-based stream), a device permutation, and ``b2r_collate_general`` per
        batch.  The training step is the model's ``train_step`` (one C call) or, for models without one, the plugin
        contract's forward / loss / backward / optimizer.step on the device batch.  GeneralModel datasets only (a
        SequentialModel needs the history CSR on the device too).  RNG: the host streams of the reference are NOT
        reproduced (a sequential NumPy / torch CPU stream cannot be drawn in parallel); results are reproducible from
        (seed, epoch)."""
        model = dataset.model
        if hasattr(model, "history_max"):
            raise ValueError("fit_on_device: sequential datasets are not supported (history collate is host-side)")
        if model.optimizer is None:
            model.optimizer = self._build_optimizer(model)
        dev = model.device
        st = dataset.__dict__.get("_b2r_dev")
        if st is None:
            corpus = dataset.corpus
            seed = int(self.device_batches) if int(self.device_batches) > 1 else 2023
            st = {"users": torch.as_tensor(np.asarray(dataset.data["user_id"], dtype=np.int64)).to(dev),
                  "items": torch.as_tensor(np.asarray(dataset.data["item_id"], dtype=np.int64)).to(dev),
                  "sampler": ops.DeviceNegativeSampler(corpus.train_clicked_set, corpus.n_users, corpus.n_items, dev, seed=seed),
                  "gen": torch.Generator(device=dev), "seed": seed, "epoch": 0}
            dataset.__dict__["_b2r_dev"] = st
        st["epoch"] += 1
        N, K, B = st["users"].numel(), int(model.num_neg), int(self.batch_size)
        neg = st["sampler"].sample(st["users"], K, st["epoch"])                     # [N, K] int64 on the device
        st["gen"].manual_seed(st["seed"] * 1_000_003 + st["epoch"])
        perm = torch.randperm(N, device=dev, generator=st["gen"])
        model.train()
        use_step = hasattr(model, "train_step") and isinstance(model.optimizer, RowSparseOptimizer)
        n_steps = (N + B - 1) // B
        NB = 3                                                                       # batch buffers: current, next, one being refilled
        bufs = st.get("bufs")
        if bufs is None or bufs[0]["item_id"].shape != (B, K + 1):
            bufs = st["bufs"] = [{"user_id": torch.empty(B, dtype=torch.int64, device=dev),
                                  "item_id": torch.empty((B, K + 1), dtype=torch.int64, device=dev),
                                  "batch_size": B, "phase": "train"} for _ in range(NB)]

        def batch(k):
            Bn = min(B, N - k * B)
            buf = bufs[k % NB]
            ops.collate_general(st["users"], st["items"], neg, perm, k * B, Bn, buf["user_id"], buf["item_id"])
            if Bn == B:
                return buf
            return {"user_id": buf["user_id"][:Bn], "item_id": buf["item_id"][:Bn], "batch_size": Bn, "phase": "train"}

        losses = []
        ops.bprmf_step_reset()
        try:
            cur = batch(0)
            for k in range(n_steps):
                nxt = batch(k + 1) if k + 1 < n_steps else None
                if use_step:
                    same = nxt is not None and nxt["item_id"].shape == cur["item_id"].shape
                    losses.append(model.train_step(cur, nxt if same else None))
                else:
                    model.optimizer.zero_grad()
                    loss = model.loss(model(cur))
                    loss.backward()
                    model.optimizer.step()
                    losses.append(loss.detach())
                cur = nxt
        finally:
            ops.bprmf_step_reset()
        out = float(torch.stack(losses).mean().cpu()) if losses else float("nan")
        _poll_ids(model)
        return out

    def fit(self, dataset, epoch=-1) -> float:
        """helpers/BaseRunner.py:174-208, step for step."""
        model = dataset.model
        opt = model.optimizer
        if opt is None or (self.graph_step and isinstance(opt, RowSparseOptimizer) and not opt._want_clock and opt.t == 0):
            model.optimizer = self._build_optimizer(model)     # (a lazily built one without the device clock is replaced)
        if self.device_sampler and model.__dict__.get("_b2r_device_sampler") is None:
            corpus = dataset.corpus
            model.__dict__["_b2r_device_sampler"] = ops.DeviceNegativeSampler(
                corpus.train_clicked_set, corpus.n_users, corpus.n_items, model.device, seed=self.device_sampler)
        if self.device_batches:
            return self.fit_on_device(dataset, epoch)
        dataset.actions_before_epoch()
        model.train()
        if self.fused_step and hasattr(model, "train_step") and isinstance(model.optimizer, RowSparseOptimizer):
            return self._fit_whole_steps(dataset, epoch)
        if self.graph_step and isinstance(model.optimizer, RowSparseOptimizer) and not getattr(model, "dropout", 0):
            return self._fit_graphed(dataset, epoch)
        losses = []
        dl = DataLoader(dataset, batch_size=self.batch_size, shuffle=True, num_workers=self.num_workers,
                        collate_fn=dataset.collate_batch, pin_memory=bool(self.pin_memory))
        for batch in dl:
            batch = batch_to_device(batch, model.device)
            item_ids = batch["item_id"]
            # per-row candidate permutation drawn from the CPU generator, as the reference does (:189)
            indices = torch.argsort(torch.rand(*item_ids.shape), dim=-1)
            rows = torch.arange(item_ids.shape[0]).unsqueeze(-1)
            batch["item_id"] = item_ids[rows, indices]
            model.optimizer.zero_grad()
            out = model(batch)
            pred = out["prediction"]
            if pred.dim() == 2:
                restored = torch.zeros(*pred.shape).to(pred.device)
                restored[rows, indices] = pred                           # autograd-tracked index_put (:201)
                out["prediction"] = restored
            loss = model.loss(out)
            loss.backward()
            model.optimizer.step()
            losses.append(loss.detach().cpu().data.numpy())
        _poll_ids(model)
        return float(np.mean(losses))

    def eval_termination(self, criterion: List[float]) -> bool:
        """helpers/BaseRunner.py:132-137 with utils.non_increasing (utils/utils.py:103-104): stop when the FIRST value
        of the last ``early_stop`` results is >= every later one of that window (not a pairwise-monotone test), or when
        the best result is more than ``early_stop`` epochs old."""
        if len(criterion) > self.early_stop:
            w = criterion[-self.early_stop:]
            if all(w[0] >= y for y in w[1:]):
                return True
        return len(criterion) - criterion.index(max(criterion)) > self.early_stop

    def evaluate(self, dataset, topks: list, metrics: list) -> Dict[str, float]:
        if self.device_metrics and hasattr(dataset.model, "eval_ranks"):
            return self.evaluate_on_device(dataset, topks, metrics)
        return self.evaluate_method(self.predict(dataset), topks, metrics)

    def evaluate_on_device(self, dataset, topks: list, metrics: list) -> Dict[str, float]:
        """helpers/BaseRunner.py:216-252 with the predictions never leaving the GPU: per batch the model returns
        the integer ranks of the ground truth (``eval_ranks``), a rank histogram is accumulated on the device and
        HR@k / NDCG@k are read off it -- (max k + 2) integers cross PCIe per evaluation instead of N x C floats."""
        model = dataset.model
        model.eval()
        kmax = max(int(k) for k in topks)
        hist, n_rows = None, 0
        dl = DataLoader(dataset, batch_size=self.eval_batch_size, shuffle=False, num_workers=self.num_workers,
                        collate_fn=dataset.collate_batch, pin_memory=bool(self.pin_memory))
        start = 0
        for batch in dl:
            bs = batch["batch_size"]
            mask_row = mask_item = None
            if model.test_all:
                rows, cols = [], []
                for i, u in enumerate(dataset.data["user_id"][start:start + bs]):
                    clicked = list(dataset.corpus.train_clicked_set[u] | dataset.corpus.residual_clicked_set[u])
                    rows.extend([i] * len(clicked))
                    cols.extend(clicked)
                mask_row = torch.tensor(rows, dtype=torch.int64, device=model.device)
                mask_item = torch.tensor(cols, dtype=torch.int64, device=model.device)
            batch = batch_to_device(batch, model.device)
            ranks = model.eval_ranks(batch, mask_row, mask_item)
            h = ops.rank_histogram(ranks, kmax)
            hist = h if hist is None else hist + h
            n_rows += bs
            start += bs
        _poll_ids(model)
        return ops.metrics_from_histogram(hist, n_rows, topks, metrics)

    def predict(self, dataset, save_prediction: bool = False) -> np.ndarray:
        """helpers/BaseRunner.py:225-252 (uses the model's no-grad ``inference`` hook when present)."""
        model = dataset.model
        model.eval()
        preds = []
        dl = DataLoader(dataset, batch_size=self.eval_batch_size, shuffle=False, num_workers=self.num_workers,
                        collate_fn=dataset.collate_batch, pin_memory=bool(self.pin_memory))
        for batch in dl:
            batch = batch_to_device(batch, model.device)
            fn = model.inference if hasattr(model, "inference") else model
            preds.extend(fn(batch)["prediction"].cpu().data.numpy())
        _poll_ids(model)
        preds = np.array(preds)
        if model.test_all:
            rows, cols = [], []
            for i, u in enumerate(dataset.data["user_id"]):
                clicked = list(dataset.corpus.train_clicked_set[u] | dataset.corpus.residual_clicked_set[u])
                rows.extend([i] * len(clicked))
                cols.extend(clicked)
            preds[rows, cols] = -np.inf
        return preds

    def print_res(self, dataset) -> str:
        return "(" + format_metric(self.evaluate(dataset, self.topk, self.metrics)) + ")"
