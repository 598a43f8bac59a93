"""Tensor-level wrappers over the C ABI (include/b200rec.h) and the autograd nodes built on them.

PyTorch is plumbing here: it owns device memory, streams and the autograd graph the reference's unchanged
runner expects (helpers/BaseRunner.py:193-206).  Every arithmetic step is a kernel in libb200rec.so.

Embedding-table gradients can leave the autograd graph in three forms, selected per table with
``set_table_mode(param, mode)``:
  'dense'   a dense [n_rows, d] tensor, bit-reproducible (one writer per row) -- what the reference's
            nn.Embedding(sparse=False) hands to stock torch.optim (exact reference semantics);
  'sparse'  a coalesced torch.sparse_coo tensor (unique rows only) for torch.optim.SparseAdam/SGD/Adagrad;
  'fused'   nothing is materialised: the contribution streams are parked on the parameter and
            ``rechorus_b200.optim.RowSparseOptimizer.step()`` reduces and applies them in one kernel.
"""
from __future__ import annotations

import os

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch

from . import lib as _lib

# --------------------------------------------------------------------------------------------------
# plumbing
# --------------------------------------------------------------------------------------------------


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _need_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _lib.B200RecError(
                "rechorus_b200 runs on a CUDA device only (got a %s tensor); there is no CPU fallback" % t.device)


def _f32c(t: torch.Tensor, what: str) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise TypeError(f"{what} must be float32, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def _i64c(t: torch.Tensor, what: str) -> torch.Tensor:
    if t.dtype != torch.int64:
        raise TypeError(f"{what} must be int64 (the reference's collate dtype), got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


_err_flags = {}


def err_flag(device: torch.device) -> torch.Tensor:
    """Per-device int32 counter of out-of-range ids seen by the kernels (they clamp, never read OOB)."""
    key = device.index if device.index is not None else torch.cuda.current_device()
    if key not in _err_flags:
        _err_flags[key] = torch.zeros(1, dtype=torch.int32, device=device)
    return _err_flags[key]


def check_ids(device: Optional[torch.device] = None) -> None:
    """Synchronising poll of the out-of-range counter; raises IndexError like ATen's embedding would."""
    device = device or torch.device("cuda", torch.cuda.current_device())
    flag = err_flag(device)
    bad = int(flag.item())
    if bad:
        flag.zero_()
        raise IndexError(f"{bad} embedding id(s) out of range (kernels clamped them to row 0)")


def poll_ids(device: torch.device) -> None:
    """check_ids, but only if a kernel has used this device's counter yet (cheap no-op before the first forward)"""
    key = device.index if device.index is not None else torch.cuda.current_device()
    if key in _err_flags:
        check_ids(device)


# --------------------------------------------------------------------------------------------------
# kernels, no autograd
# --------------------------------------------------------------------------------------------------

def rowdot(Q: torch.Tensor, qid: Optional[torch.Tensor], T: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
    """pred[b,c] = <Q[qid[b]], T[ids[b,c]]> (qid None: row b).  BPRMF.py:39-42 / SASRec.py:80-81."""
    _need_cuda(Q, T, ids, qid)
    Q, T, ids = _f32c(Q, "Q"), _f32c(T, "T"), _i64c(ids, "ids")
    if ids.dim() != 2:
        raise ValueError("ids must be [B, C]")
    B, Cn = ids.shape
    d = T.shape[1]
    if Q.shape[1] != d:
        raise ValueError(f"emb width mismatch: Q {Q.shape[1]} vs T {d}")
    if qid is not None:
        qid = _i64c(qid, "qid")
        if qid.numel() != B:
            raise ValueError("qid must have B entries")
    elif Q.shape[0] != B:
        raise ValueError("dense Q must have B rows")
    pred = torch.empty((B, Cn), dtype=torch.float32, device=T.device)
    L = _lib.load()
    _lib.check(L.b2r_rowdot_fwd(_p(Q), _p(qid), Q.shape[0], _p(T), _p(ids), T.shape[0], _p(pred), B, Cn, d,
                                _p(err_flag(T.device)), _stream()), "b2r_rowdot_fwd")
    return pred


def rowdot_bwd_query(g: torch.Tensor, T: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
    """dQ[b,:] = sum_c g[b,c] * T[ids[b,c],:]  -> dense [B, d], fixed summation order."""
    _need_cuda(g, T, ids)
    g, T, ids = _f32c(g, "g"), _f32c(T, "T"), _i64c(ids, "ids")
    B, Cn = ids.shape
    d = T.shape[1]
    dQ = torch.empty((B, d), dtype=torch.float32, device=T.device)
    L = _lib.load()
    _lib.check(L.b2r_rowdot_bwd_query(_p(g), _p(T), _p(ids), T.shape[0], _p(dQ), B, Cn, d, _stream()),
               "b2r_rowdot_bwd_query")
    return dQ


def gather_rows(T: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
    """out[..., :] = T[ids[...], :]   (nn.Embedding forward)."""
    _need_cuda(T, ids)
    T, ids = _f32c(T, "T"), _i64c(ids, "ids")
    d = T.shape[1]
    out = torch.empty(tuple(ids.shape) + (d,), dtype=torch.float32, device=T.device)
    L = _lib.load()
    _lib.check(L.b2r_gather_rows(_p(T), _p(ids), T.shape[0], _p(out), ids.numel(), d, _p(err_flag(T.device)),
                                 _stream()), "b2r_gather_rows")
    return out


def pairdot(Q: torch.Tensor, qidx: torch.Tensor, T: torch.Tensor, rows: torch.Tensor) -> torch.Tensor:
    """out[e] = <Q[qidx[e]], T[rows[e]]>, rows[e] < 0 -> 0   (b2r_pairdot_fwd; sharded-table owner side)"""
    _need_cuda(Q, T, qidx, rows)
    Q, T = _f32c(Q, "Q"), _f32c(T, "T")
    qidx, rows = _i64c(qidx.reshape(-1), "qidx"), _i64c(rows.reshape(-1), "rows")
    n = rows.numel()
    out = torch.empty(n, dtype=torch.float32, device=T.device)
    L = _lib.load()
    _lib.check(L.b2r_pairdot_fwd(_p(Q), _p(qidx), Q.shape[0], _p(T), _p(rows), T.shape[0], _p(out), n, T.shape[1],
                                 _p(err_flag(T.device)), _stream()), "b2r_pairdot_fwd")
    return out


def pairdot_p2p(Q: torch.Tensor, qidx: torch.Tensor, T: torch.Tensor, rows: torch.Tensor, out_tab: torch.Tensor,
                seg: int) -> None:
    """pairdot whose score for pair e goes to out_tab[e // seg][e % seg]; out_tab: int64 device tensor of float*
    addresses (peer-mapped receive buffers).  Opt-in groundwork for the fused exchange of config 5 (b2r_pairdot_fwd_p2p)."""
    _need_cuda(Q, qidx, T, rows, out_tab)
    Q, T = _f32c(Q, "Q"), _f32c(T, "T")
    qidx, rows = _i64c(qidx.reshape(-1), "qidx"), _i64c(rows.reshape(-1), "rows")
    out_tab = _i64c(out_tab, "out_tab")
    n = rows.numel()
    if n != out_tab.numel() * int(seg):
        raise ValueError("pairdot_p2p: n must equal len(out_tab) * seg")
    _lib.check(_lib.load().b2r_pairdot_fwd_p2p(_p(Q), _p(qidx), Q.shape[0], _p(T), _p(rows), T.shape[0], _p(out_tab),
                                               int(seg), n, T.shape[1], _p(err_flag(Q.device)), _stream()),
               "b2r_pairdot_fwd_p2p")


def pair_runs_sum(out: torch.Tensor, key: torch.Tensor, rows: torch.Tensor, coef: torch.Tensor, T: torch.Tensor) -> None:
    """out[key[e]] = sum over each contiguous run of valid pairs with that key of coef[e] * T[rows[e]]
    (b2r_pair_runs_sum; rows < 0 = unused slot)"""
    _need_cuda(out, key, rows, coef, T)
    key, rows = _i64c(key.reshape(-1), "key"), _i64c(rows.reshape(-1), "rows")
    coef, T = _f32c(coef.reshape(-1), "coef"), _f32c(T, "T")
    L = _lib.load()
    _lib.check(L.b2r_pair_runs_sum(_p(key), _p(rows), _p(coef), _p(T), T.shape[0], _p(out), out.shape[0], rows.numel(),
                                   T.shape[1], _stream()), "b2r_pair_runs_sum")


def gt_rank(pred: torch.Tensor) -> torch.Tensor:
    """rank[r] = #{c : pred[r, c] >= pred[r, 0]} as int64 on the device (helpers/BaseRunner.py:63)."""
    _need_cuda(pred)
    pred = _f32c(pred, "pred")
    if pred.dim() != 2 or pred.shape[1] < 1:
        raise ValueError("pred must be [N, C] with C >= 1")
    N, Cn = pred.shape
    rank = torch.empty(N, dtype=torch.int64, device=pred.device)
    _lib.check(_lib.load().b2r_gt_rank(_p(pred), N, Cn, Cn, _p(rank), _stream()), "b2r_gt_rank")
    return rank


def rank_histogram(rank: torch.Tensor, kmax: int) -> torch.Tensor:
    """hist[k] = #{rank == k} for k <= kmax, hist[kmax + 1] = #{rank > kmax}; int64 [kmax + 2] on the device."""
    _need_cuda(rank)
    rank = _i64c(rank.reshape(-1), "rank")
    hist = torch.empty(kmax + 2, dtype=torch.int64, device=rank.device)
    _lib.check(_lib.load().b2r_rank_histogram(_p(rank), rank.numel(), int(kmax), _p(hist), _stream()),
               "b2r_rank_histogram")
    return hist


def metrics_from_histogram(hist, n_rows: int, topk, metrics) -> dict:
    """HR@k / NDCG@k of helpers/BaseRunner.py:64-76 from the rank histogram: every row of rank r <= k contributes
    1 (HR) or 1/log2(r+1) (NDCG); the mean is over all n_rows rows."""
    import numpy as np
    h = np.asarray(hist.cpu() if isinstance(hist, torch.Tensor) else hist, dtype=np.int64)
    out = {}
    for k in topk:
        r = np.arange(1, int(k) + 1)
        cnt = h[1:int(k) + 1].astype(np.float64)
        for metric in metrics:
            key = f"{metric}@{k}"
            if metric == "HR":
                out[key] = cnt.sum() / n_rows
            elif metric == "NDCG":
                out[key] = (cnt / np.log2(r + 1)).sum() / n_rows
            else:
                raise ValueError(f"Undefined evaluation metric: {metric}.")
    return out


def rank_all_items(q: torch.Tensor, table: torch.Tensor, target: torch.Tensor,
                   mask_row: Optional[torch.Tensor] = None, mask_item: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Ranks under the test_all protocol (BaseModel.py:194-198 + BaseRunner.py:244-251) without the [B, n_items]
    score matrix: q [B, d] query rows, table [n_items, d], target [B]; (mask_row, mask_item) unique (batch row,
    item id) pairs whose scores the reference overwrites with -inf."""
    _need_cuda(q, table, target, mask_row, mask_item)
    q, table, target = _f32c(q, "q"), _f32c(table, "table"), _i64c(target.reshape(-1), "target")
    if q.dim() != 2 or table.dim() != 2 or q.shape[1] != table.shape[1] or target.numel() != q.shape[0]:
        raise ValueError("rank_all_items: q [B, d], table [n_items, d], target [B]")
    n_mask = 0
    if mask_row is not None:
        mask_row, mask_item = _i64c(mask_row.reshape(-1), "mask_row"), _i64c(mask_item.reshape(-1), "mask_item")
        if mask_row.numel() != mask_item.numel():
            raise ValueError("mask_row and mask_item must have the same length")
        n_mask = mask_row.numel()
    B, d = q.shape
    rank = torch.empty(B, dtype=torch.int64, device=q.device)
    s0 = torch.empty(B, dtype=torch.float32, device=q.device)
    _lib.check(_lib.load().b2r_rank_all_items(_p(q), d, _p(table), _p(target), B, table.shape[0], d,
                                              _p(mask_row) if n_mask else None, _p(mask_item) if n_mask else None,
                                              n_mask, _p(s0), _p(rank), _p(err_flag(q.device)), _stream()),
               "b2r_rank_all_items")
    return rank


class DeviceNegativeSampler:
    """Groundwork for SURVEY.md §8 f3: models/BaseModel.py:206-214 on the device.  Holds the training clicks as a CSR
    on the GPU and draws ``num_neg`` non-clicked items per training row with b2r_sample_negatives.  Same distribution
    as the reference's NumPy loop, a different (counter-based, reproducible from (seed, epoch)) random stream."""

    def __init__(self, train_clicked_set: dict, n_users: int, n_items: int, device, seed: int = 0):
        import numpy as np
        ptr = np.zeros(n_users + 1, dtype=np.int64)
        for u, items in train_clicked_set.items():
            if 0 <= int(u) < n_users:
                ptr[int(u) + 1] = len(items)
        ptr = np.cumsum(ptr)
        flat = np.zeros(max(int(ptr[-1]), 1), dtype=np.int64)
        for u, items in train_clicked_set.items():
            if 0 <= int(u) < n_users and len(items):
                flat[ptr[int(u)]:ptr[int(u) + 1]] = np.sort(np.fromiter(items, dtype=np.int64, count=len(items)))
        self.ptr = torch.from_numpy(ptr).to(device)
        self.items = torch.from_numpy(flat).to(device)
        self.n_users, self.n_items, self.seed = int(n_users), int(n_items), int(seed)

    def sample(self, user_ids: torch.Tensor, num_neg: int, epoch: int) -> torch.Tensor:
        _need_cuda(user_ids, self.ptr)
        user_ids = _i64c(user_ids.reshape(-1), "user_ids")
        out = torch.empty((user_ids.numel(), int(num_neg)), dtype=torch.int64, device=user_ids.device)
        _lib.check(_lib.load().b2r_sample_negatives(_p(user_ids), user_ids.numel(), int(num_neg), _p(self.ptr),
                                                    _p(self.items), self.n_users, self.n_items,
                                                    self.seed & 0xFFFFFFFFFFFFFFFF, int(epoch) & 0xFFFFFFFF, _p(out),
                                                    _p(err_flag(user_ids.device)), _stream()), "b2r_sample_negatives")
        return out


def collate_general(users: torch.Tensor, items: torch.Tensor, neg: torch.Tensor, perm: Optional[torch.Tensor], start: int,
                    Bn: int, out_uid: torch.Tensor, out_iid: torch.Tensor) -> None:
    """b2r_collate_general: one GeneralModel training batch assembled on the device (BaseModel.py:192-203,135-152)"""
    _need_cuda(users, items, neg, perm, out_uid, out_iid)
    K = neg.shape[1]
    if out_iid.shape[1] != K + 1 or out_iid.shape[0] < Bn or out_uid.numel() < Bn:
        raise ValueError("collate_general: output buffers do not fit the batch")
    for t in (users, items, neg, out_uid, out_iid):
        if t.dtype != torch.int64 or not t.is_contiguous():
            raise TypeError("collate_general: contiguous int64 tensors expected")
    if perm is not None and (perm.dtype != torch.int64 or not perm.is_contiguous()):
        raise TypeError("collate_general: perm must be contiguous int64")
    _lib.check(_lib.load().b2r_collate_general(_p(users), _p(items), _p(neg), _p(perm), int(start), int(Bn), int(K),
                                               _p(out_uid), _p(out_iid), _stream()), "b2r_collate_general")


def adam_exact_advance(W: torch.Tensor, m: torch.Tensor, v: torch.Tensor, last: torch.Tensor, upto: int, opt,
                       rows: Optional[torch.Tensor] = None, stamp: int = 0, state_ld: int = 0) -> None:
    """Groundwork for the exact dense-Adam mode (b2r_adam_exact_advance): bring the listed unique rows (all rows when
    ``rows`` is None) of a row-sparsely updated table to optimizer step ``upto``, as torch.optim.Adam over the whole
    table would have moved them on zero data gradients.  ``last`` int32 [n_rows] is read and re-stamped."""
    _need_cuda(W, m, v, last, rows)
    if last.dtype != torch.int32 or not last.is_contiguous():
        raise TypeError("last must be a contiguous int32 tensor")
    n = 0
    if rows is not None:
        rows = _i64c(rows.reshape(-1), "rows")
        n = rows.numel()
    o = _lib.Optim(1, opt.lr, opt.betas[0], opt.betas[1], opt.eps, opt.weight_decay, 1.0, 1.0, state_ld)
    _lib.check(_lib.load().b2r_adam_exact_advance(_p(rows), n, W.shape[0], W.shape[1], _p(W), _p(m), _p(v), _p(last),
                                                  int(upto), int(stamp), C.byref(o), _p(err_flag(W.device)), _stream()),
               "b2r_adam_exact_advance")


def bpr_loss_and_grad(pred: torch.Tensor, want_grad: bool = True) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """BaseModel.py:175-189 value and closed-form d loss / d pred in one pass."""
    _need_cuda(pred)
    pred = _f32c(pred, "pred")
    if pred.dim() != 2:
        raise ValueError("pred must be [B, C]")
    B, Cn = pred.shape
    loss = torch.empty((), dtype=torch.float32, device=pred.device)
    grad = torch.empty_like(pred) if want_grad else None
    rows = torch.empty(B, dtype=torch.float32, device=pred.device)
    L = _lib.load()
    _lib.check(L.b2r_bpr_loss(_p(pred), _p(loss), _p(grad), _p(rows), B, Cn, _stream()), "b2r_bpr_loss")
    return loss, grad


# --------------------------------------------------------------------------------------------------
# index plan + contribution sources
# --------------------------------------------------------------------------------------------------

@dataclass
class Source:
    """One contribution stream into a table gradient (b2r_grad_source): position p adds
    coef[p] * src[row(p)], row(p) = src_id[p // div] if src_id is given else p // div."""
    src: torch.Tensor                      # [*, d] float32
    n: int                                 # number of positions
    coef: Optional[torch.Tensor] = None    # [n] float32
    src_id: Optional[torch.Tensor] = None  # int64
    div: int = 1
    ld: int = 0                            # leading dimension of src in floats (0: d)

    def c_struct(self) -> _lib.GradSource:
        return _lib.GradSource(_p(self.src), _p(self.coef), _p(self.src_id), self.n, self.div, self.ld)


class IndexPlan:
    """Sorted, de-duplicated view of the ids a batch touches in one table (b2r_plan_build)."""

    def __init__(self, ids: torch.Tensor, n_rows: int, ignore_id: int = -1, ignore_n: int = 0):
        _need_cuda(ids)
        ids = _i64c(ids.reshape(-1), "ids")
        n = ids.numel()
        if n == 0:
            raise ValueError("empty id list")
        dev = ids.device
        self.n, self.n_rows, self.device = n, int(n_rows), dev
        self.sorted_key = torch.empty(n, dtype=torch.int32, device=dev)   # uint32 bits
        self.sorted_pos = torch.empty(n, dtype=torch.int32, device=dev)   # uint32 bits
        self.seg_start = torch.empty(n, dtype=torch.int32, device=dev)
        self.n_uniq = torch.empty(1, dtype=torch.int32, device=dev)
        L = _lib.load()
        nbytes = L.b2r_plan_workspace_bytes(n, self.n_rows)
        if nbytes == 0:
            raise _lib.B200RecError(f"b2r_plan_workspace_bytes({n}, {n_rows}) = 0: " + L.b2r_last_error().decode())
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _lib.check(L.b2r_plan_build_ex(_p(ids), n, self.n_rows, int(ignore_id), int(ignore_n), _p(self.sorted_key),
                                       _p(self.sorted_pos), _p(self.seg_start), _p(self.n_uniq), _p(ws), nbytes,
                                       _p(err_flag(dev)), _stream()), "b2r_plan_build_ex")
        self._ws = ws   # keep alive until the stream has consumed it

    def count(self) -> int:
        """number of unique rows (synchronises)"""
        return int(self.n_uniq.item())

    def _apply(self, d: int, sources: Sequence[Source], mode: int, uniq=None, rows=None, dense=None, W=None,
               m=None, v=None, opt: Optional[_lib.Optim] = None) -> None:
        if not 1 <= len(sources) <= 2:
            raise ValueError("one or two sources")
        if sum(s.n for s in sources) != self.n:
            raise ValueError("sources do not cover the plan")
        s0 = sources[0].c_struct()
        s1 = sources[1].c_struct() if len(sources) == 2 else None
        L = _lib.load()
        _lib.check(L.b2r_segment_apply(_p(self.sorted_key), _p(self.sorted_pos), _p(self.seg_start),
                                       _p(self.n_uniq), self.n, self.n_rows, d, C.byref(s0),
                                       C.byref(s1) if s1 is not None else None, mode, _p(uniq), _p(rows),
                                       _p(dense), _p(W), _p(m), _p(v),
                                       C.byref(opt) if opt is not None else None, _stream()),
                   "b2r_segment_apply")

    def reduce_rows(self, d: int, sources: Sequence[Source]) -> Tuple[torch.Tensor, torch.Tensor]:
        """row-sparse gradient: (unique row ids [nu] int64 ascending, grad rows [nu, d]); synchronises once."""
        uniq = torch.empty(self.n, dtype=torch.int64, device=self.device)
        rows = torch.empty((self.n, d), dtype=torch.float32, device=self.device)
        self._apply(d, sources, 0, uniq=uniq, rows=rows)
        nu = self.count()
        if nu > 0 and int(uniq[nu - 1].item()) >= self.n_rows:      # trailing sentinel run of ignored positions
            nu -= 1
        return uniq[:nu], rows[:nu]

    def add_to_dense(self, dense: torch.Tensor, sources: Sequence[Source]) -> None:
        self._apply(dense.shape[1], sources, 1, dense=dense)

    def apply_optimizer(self, W: torch.Tensor, m, v, opt: _lib.Optim, sources: Sequence[Source]) -> None:
        self._apply(W.shape[1], sources, 2, W=W, m=m, v=v, opt=opt)


_bucket_ws = {}


class BucketPlan:
    """Bucketed partition of a batch's (row id, position) pairs (b2r_bucket_partition) -- the hot-path sibling of
    IndexPlan for the dense-add and fused-optimizer outputs (d in {32, 64, 128}).  The workspace is cached per
    (device, stream, n, n_rows) and reused in stream order."""

    SUPPORTED_D = (32, 64, 128)

    def __init__(self, ids: torch.Tensor, n_rows: int, ignore_id: int = -1, ignore_n: int = 0):
        _need_cuda(ids)
        ids = _i64c(ids.reshape(-1), "ids")
        n = ids.numel()
        if n == 0:
            raise ValueError("empty id list")
        self.n, self.n_rows, self.device = n, int(n_rows), ids.device
        L = _lib.load()
        key = (ids.device.index, _stream(), n, self.n_rows)
        ws = _bucket_ws.get(key)
        if ws is None:
            nbytes = L.b2r_bucket_workspace_bytes(n, self.n_rows)
            if nbytes == 0:
                raise _lib.B200RecError(f"b2r_bucket_workspace_bytes({n}, {n_rows}) = 0")
            ws = torch.empty(nbytes, dtype=torch.uint8, device=ids.device)
            _lib.check(L.b2r_bucket_workspace_init(_p(ws), nbytes, n, self.n_rows, _stream()),
                       "b2r_bucket_workspace_init")
            _bucket_ws[key] = ws
        self.ws = ws
        self._ids = ids
        _lib.check(L.b2r_bucket_partition(_p(ids), n, self.n_rows, int(ignore_id), int(ignore_n), _p(ws), ws.numel(),
                                          _p(err_flag(ids.device)), _stream()), "b2r_bucket_partition")

    def _apply(self, d, sources, mode, dense=None, W=None, m=None, v=None, opt=None):
        if not 1 <= len(sources) <= 2:
            raise ValueError("one or two sources")
        s0 = sources[0].c_struct()
        s1 = sources[1].c_struct() if len(sources) == 2 else None
        L = _lib.load()
        _lib.check(L.b2r_bucket_apply(_p(self.ws), self.n, self.n_rows, d, C.byref(s0),
                                      C.byref(s1) if s1 is not None else None, mode, _p(dense), _p(W), _p(m), _p(v),
                                      C.byref(opt) if opt is not None else None, _stream()), "b2r_bucket_apply")

    def add_to_dense(self, dense: torch.Tensor, sources: Sequence[Source]) -> None:
        self._apply(dense.shape[1], sources, 1, dense=dense)

    def apply_optimizer(self, W: torch.Tensor, m, v, opt: _lib.Optim, sources: Sequence[Source]) -> None:
        self._apply(W.shape[1], sources, 2, W=W, m=m, v=v, opt=opt)


_direct_ws = {}


class DirectPlan:
    """b2r_direct_plan_build / _apply: BucketPlan's job in two launches (scatter into fixed-capacity bucket regions + one
    sort), same apply kernel.  Workspace cached per (device, stream, n, n_rows) and reused in stream order."""

    def __init__(self, ids: torch.Tensor, n_rows: int, ignore_id: int = -1, ignore_n: int = 0):
        _need_cuda(ids)
        ids = _i64c(ids.reshape(-1), "ids")
        n = ids.numel()
        if n == 0:
            raise ValueError("empty id list")
        self.n, self.n_rows, self.device = n, int(n_rows), ids.device
        L = _lib.load()
        key = (ids.device.index, _stream(), n, self.n_rows)
        ws = _direct_ws.get(key)
        if ws is None:
            nbytes = L.b2r_direct_plan_workspace_bytes(n, self.n_rows)
            if nbytes == 0:
                raise _lib.B200RecError(f"b2r_direct_plan_workspace_bytes({n}, {n_rows}) = 0")
            ws = torch.empty(nbytes, dtype=torch.uint8, device=ids.device)
            _lib.check(L.b2r_direct_plan_init(_p(ws), nbytes, n, self.n_rows, _stream()), "b2r_direct_plan_init")
            _direct_ws[key] = ws
        self.ws, self._ids = ws, ids
        _lib.check(L.b2r_direct_plan_build(_p(ids), n, self.n_rows, int(ignore_id), int(ignore_n), _p(ws), ws.numel(),
                                           _p(err_flag(ids.device)), _stream()), "b2r_direct_plan_build")

    @staticmethod
    def available(n: int, n_rows: int) -> bool:
        return _lib.load().b2r_direct_plan_workspace_bytes(int(n), int(n_rows)) != 0

    def _apply(self, d, sources, mode, dense=None, W=None, m=None, v=None, opt=None):
        if not 1 <= len(sources) <= 2:
            raise ValueError("one or two sources")
        s0 = sources[0].c_struct()
        s1 = sources[1].c_struct() if len(sources) == 2 else None
        _lib.check(_lib.load().b2r_direct_plan_apply(_p(self.ws), self.n, self.n_rows, d, C.byref(s0),
                                                     C.byref(s1) if s1 is not None else None, mode, _p(dense), _p(W), _p(m),
                                                     _p(v), C.byref(opt) if opt is not None else None, _stream()),
                   "b2r_direct_plan_apply")

    def add_to_dense(self, dense: torch.Tensor, sources: Sequence[Source]) -> None:
        self._apply(dense.shape[1], sources, 1, dense=dense)

    def apply_optimizer(self, W: torch.Tensor, m, v, opt: _lib.Optim, sources: Sequence[Source]) -> None:
        self._apply(W.shape[1], sources, 2, W=W, m=m, v=v, opt=opt)


_PLAN_KIND = os.environ.get("B2R_PLAN", "direct")      # direct | bucket (A/B)


def make_plan(ids: torch.Tensor, n_rows: int, d: int, ignore_id: int = -1, ignore_n: int = 0):
    """DirectPlan (two launches) where the bucket kernels exist (d = 32/64/128), IndexPlan (device radix sort) otherwise;
    B2R_PLAN=bucket selects the four-launch BucketPlan for A/B runs"""
    if d in BucketPlan.SUPPORTED_D:
        if _PLAN_KIND == "direct" and DirectPlan.available(ids.numel(), n_rows):
            return DirectPlan(ids, n_rows, ignore_id, ignore_n)
        return BucketPlan(ids, n_rows, ignore_id, ignore_n)
    return IndexPlan(ids, n_rows, ignore_id, ignore_n)


def scatter_add_atomic(dense: torch.Tensor, ids: torch.Tensor, source: Source) -> None:
    """order-nondeterministic dense += via red.global.add.v4.f32 (b2r_scatter_add_atomic)"""
    ids = _i64c(ids.reshape(-1), "ids")
    s = source.c_struct()
    L = _lib.load()
    _lib.check(L.b2r_scatter_add_atomic(_p(ids), dense.shape[0], C.byref(s), dense.shape[1], _p(dense),
                                        _p(err_flag(dense.device)), _stream()), "b2r_scatter_add_atomic")


def dense_optim(W: torch.Tensor, grad: torch.Tensor, m, v, opt: _lib.Optim) -> None:
    L = _lib.load()
    _lib.check(L.b2r_dense_optim(_p(W), _p(grad), _p(m), _p(v), W.numel(), C.byref(opt), _stream()),
               "b2r_dense_optim")


# --------------------------------------------------------------------------------------------------
# how a table's gradient leaves autograd
# --------------------------------------------------------------------------------------------------

TABLE_MODES = ("dense", "sparse", "fused")


def set_table_mode(param: torch.nn.Parameter, mode: str) -> None:
    if mode not in TABLE_MODES:
        raise ValueError(f"table mode must be one of {TABLE_MODES}")
    param._b2r_mode = mode
    if not hasattr(param, "_b2r_pending"):
        param._b2r_pending = []


def table_mode(param: torch.Tensor) -> str:
    return getattr(param, "_b2r_mode", "dense")


def _table_grad(table: torch.Tensor, ids: torch.Tensor, source: Source, ignore_id: int = -1
                ) -> Optional[torch.Tensor]:
    """ignore_id: positions holding this id contribute exactly zero (padding) and are dropped from the plan"""
    mode = table_mode(table)
    if mode == "fused":
        table._b2r_pending.append((ids.reshape(-1), source, ignore_id))
        return None
    ign_n = ids.numel() if ignore_id >= 0 else 0
    if mode == "sparse":
        plan = IndexPlan(ids, table.shape[0], ignore_id, ign_n)
        uniq, rows = plan.reduce_rows(table.shape[1], [source])
        return torch.sparse_coo_tensor(uniq.unsqueeze(0), rows, size=tuple(table.shape), is_coalesced=True)
    dense = torch.zeros_like(table)
    make_plan(ids, table.shape[0], table.shape[1], ignore_id, ign_n).add_to_dense(dense, [source])
    return dense


# --------------------------------------------------------------------------------------------------
# autograd nodes
# --------------------------------------------------------------------------------------------------

class _Gather(torch.autograd.Function):
    """out = T[ids]; backward: row-sparse scatter of grad_out (nn.Embedding fwd/bwd)."""

    @staticmethod
    def forward(ctx, table, ids):
        ctx.save_for_backward(ids)
        ctx.table = table
        return gather_rows(table, ids)

    @staticmethod
    def backward(ctx, gout):
        (ids,) = ctx.saved_tensors
        table = ctx.table
        gout = _f32c(gout, "grad").reshape(-1, table.shape[1])
        g = _table_grad(table, ids, Source(src=gout, n=ids.numel()))
        return g, None


class _Score(torch.autograd.Function):
    """pred[b,c] = <q[b], T[ids[b,c]]> with q a dense [B,d] activation; backward gives dq and the
    row-sparse table gradient g[b,c] * q[b] (BPRMF.py:42 / SASRec.py:81 and their autograd)."""

    @staticmethod
    def forward(ctx, q, table, ids):
        q = _f32c(q, "q")
        ctx.save_for_backward(q, ids)
        ctx.table = table
        return rowdot(q, None, table, ids)

    @staticmethod
    def backward(ctx, g):
        q, ids = ctx.saved_tensors
        table = ctx.table
        g = _f32c(g, "grad_pred")
        dq = rowdot_bwd_query(g, table, ids) if ctx.needs_input_grad[0] else None
        gt = None
        if ctx.needs_input_grad[1]:
            gt = _table_grad(table, ids, Source(src=q, n=ids.numel(), coef=g.reshape(-1), div=ids.shape[1]))
        return dq, gt, None


class _BprLoss(torch.autograd.Function):
    """GeneralModel.loss (BaseModel.py:175-189): value and closed-form gradient from one kernel."""

    @staticmethod
    def forward(ctx, pred):
        loss, grad = bpr_loss_and_grad(pred, want_grad=True)
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, gl):
        (grad,) = ctx.saved_tensors
        return grad * gl


def listwise_kind(loss_n: str) -> Tuple[int, int]:
    """(kind, hard) of b2r_listwise_loss for a reference --loss_n string (BaseImpressionModel.py:26-27,58-128)"""
    if "BPR" in loss_n:
        if "simple" in loss_n:
            raise ValueError("loss 'BPR...simple' returns a vector in the reference (BaseImpressionModel.py:79-81) and cannot "
                             "be back-propagated there; not provided")
        kind = 1 if "after" in loss_n else (2 if "before" in loss_n else 0)
        return kind, 1 if "hard" in loss_n else 0
    table = {"listnet": 3, "softmaxCE": 4, "attention_rank": 5}
    if loss_n not in table:
        raise ValueError("Undefined loss function: {}".format(loss_n))
    return table[loss_n], 0


class _ListwiseLoss(torch.autograd.Function):
    """ImpressionModel.loss (BaseImpressionModel.py:44-128): value and closed-form gradient from b2r_listwise_loss."""

    @staticmethod
    def forward(ctx, pred, target, kind, hard, max_pos):
        _need_cuda(pred, target)
        pred, target = _f32c(pred, "prediction"), _i64c(target, "target")
        if pred.dim() != 2 or tuple(target.shape) != tuple(pred.shape):
            raise ValueError("prediction and target must both be [B, Cn]")
        B, Cn = pred.shape
        loss = torch.empty((), dtype=torch.float32, device=pred.device)
        grad = torch.empty_like(pred)
        L = _lib.load()
        ws = _ws(L.b2r_listwise_workspace_bytes(B), pred.device)
        _lib.check(L.b2r_listwise_loss(_p(pred), _p(target), B, Cn, int(max_pos), int(kind), int(hard), _p(loss), _p(grad),
                                       _p(ws), ws.numel(), _stream()), "b2r_listwise_loss")
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, gl):
        (grad,) = ctx.saved_tensors
        return grad * gl, None, None, None, None


def listwise_loss(pred: torch.Tensor, target: torch.Tensor, loss_n: str, max_pos: int) -> torch.Tensor:
    kind, hard = listwise_kind(loss_n)
    return _ListwiseLoss.apply(pred, target, kind, hard, max_pos)


def embedding(table: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
    return _Gather.apply(table, ids)


def score(q: torch.Tensor, table: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
    return _Score.apply(q, table, ids)


def bpr_loss(pred: torch.Tensor) -> torch.Tensor:
    return _BprLoss.apply(pred)


# --------------------------------------------------------------------------------------------------
# whole-step entry points (one C call per training step)
# --------------------------------------------------------------------------------------------------

def bprmf_fused_fwd_bwd(U, uid, I, iid):
    """b2r_bprmf_fused_fwd_bwd -> (pred [B,C], grad_pred [B,C], row_loss [B], dQ [B,d]); raises when the shape has
    no fused variant."""
    _need_cuda(U, I, uid, iid)
    U, I, uid, iid = _f32c(U, "U"), _f32c(I, "I"), _i64c(uid, "uid"), _i64c(iid, "iid")
    B, Cn = iid.shape
    d = U.shape[1]
    pred = torch.empty((B, Cn), dtype=torch.float32, device=U.device)
    gp = torch.empty((B, Cn), dtype=torch.float32, device=U.device)
    rl = torch.empty(B, dtype=torch.float32, device=U.device)
    dq = torch.empty((B, d), dtype=torch.float32, device=U.device)
    L = _lib.load()
    _lib.check(L.b2r_bprmf_fused_fwd_bwd(_p(U), _p(uid), U.shape[0], _p(I), _p(iid), I.shape[0], _p(pred), _p(gp),
                                         _p(rl), _p(dq), B, Cn, d, _p(err_flag(U.device)), _stream()),
               "b2r_bprmf_fused_fwd_bwd")
    return pred, gp, rl, dq


_step_ctx = {}


class _StepContext:
    """b2r_bprmf_ctx: workspace + side stream + double-buffered index plans for one (B, C, d, table sizes)."""

    def __init__(self, device, B, Cn, d, n_users, n_items):
        L = _lib.load()
        nbytes = L.b2r_bprmf_step_workspace_bytes(B, Cn, d, n_users, n_items)
        if nbytes == 0:
            raise _lib.B200RecError("b2r_bprmf_step_workspace_bytes returned 0")
        self.ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        self.handle = C.c_void_p()
        _lib.check(L.b2r_bprmf_ctx_create(C.byref(self.handle), B, Cn, d, n_users, n_items, _p(self.ws), nbytes),
                   "b2r_bprmf_ctx_create")

    def __del__(self):
        try:
            if self.handle:
                _lib.load().b2r_bprmf_ctx_destroy(self.handle)
        except Exception:
            pass


def bprmf_step_reset() -> None:
    """Drop every step context's prefetched plan (b2r_bprmf_ctx_reset).  Runners call this at the start of an epoch
    and when one is aborted: a prefetched plan is recognised by the id tensors' addresses, which the caching
    allocator may reuse for another batch."""
    L = _lib.load()
    for ctx in _step_ctx.values():
        L.b2r_bprmf_ctx_reset(ctx.handle)


def bprmf_train_step(U: torch.nn.Parameter, I: torch.nn.Parameter, optimizer, uid: torch.Tensor,
                     iid: torch.Tensor, next_uid: Optional[torch.Tensor] = None,
                     next_iid: Optional[torch.Tensor] = None) -> torch.Tensor:
    """b2r_bprmf_train_step: forward + BPR loss + backward + fused row-sparse optimizer, in place.
    next_uid/next_iid (optional): the NEXT batch's id tensors, already on the device and left untouched until
    the next call -- their index plan is built on the side stream while this step's kernels run."""
    from .optim import RowSparseOptimizer
    if not isinstance(optimizer, RowSparseOptimizer):
        raise _lib.B200RecError("train_step needs model.optimizer to be a rechorus_b200.optim.RowSparseOptimizer")
    _need_cuda(U, I, uid, iid, next_uid, next_iid)
    uid, iid = _i64c(uid, "user_id"), _i64c(iid, "item_id")
    B, Cn = iid.shape
    d = U.shape[1]
    eu, ei = optimizer.entry(U), optimizer.entry(I)
    if eu["wd"] != ei["wd"] or eu["state_ld"] != ei["state_ld"]:
        raise _lib.B200RecError("fused step expects one weight decay / state layout for both tables")
    L = _lib.load()
    key = (U.device.index, B, Cn, d, U.shape[0], I.shape[0], _stream())
    ctx = _step_ctx.get(key)
    if ctx is None:
        ctx = _step_ctx[key] = _StepContext(U.device, B, Cn, d, U.shape[0], I.shape[0])
    if next_uid is not None and next_iid is not None:
        if tuple(next_iid.shape) != (B, Cn) or next_uid.numel() != B:
            raise ValueError("prefetched batch must have the same shape as the current one")
        if not (next_uid.is_contiguous() and next_iid.is_contiguous()) or next_uid.dtype != torch.int64 \
                or next_iid.dtype != torch.int64:
            # the side stream reads these after this call returns: a temporary contiguous copy would be freed under it
            raise ValueError("prefetched id tensors must be contiguous int64 tensors that stay alive until the next step")
    else:
        next_uid = next_iid = None
    optimizer.advance()
    opt = optimizer._opt(ei["wd"], ei["state_ld"])
    tables = _lib.BprmfTables(_p(U.data), _p(I.data), _p(eu["m"]), _p(eu["v"]), _p(ei["m"]), _p(ei["v"]),
                              U.shape[0], I.shape[0], d, 0)
    loss = torch.empty((), dtype=torch.float32, device=U.device)
    _lib.check(L.b2r_bprmf_train_step(ctx.handle, C.byref(tables), _p(uid), _p(iid), _p(next_uid), _p(next_iid),
                                      C.byref(opt), _p(loss), _p(err_flag(U.device)), _stream()),
               "b2r_bprmf_train_step")
    return loss


# --------------------------------------------------------------------------------------------------
# dense layers (NeuMF MLP tower, SASRec q/k/v + FFN, LayerNorm, attention) -- kernels + autograd nodes
# --------------------------------------------------------------------------------------------------

def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


def _rows_ld(t: torch.Tensor) -> Tuple[int, int]:
    """(rows, leading dimension) of a 2-D float32 tensor whose last dim is contiguous"""
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError("expected a 2-D tensor with a contiguous last dimension")
    return t.shape[0], t.stride(0)


def linear_fwd(x: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor], relu: bool) -> torch.Tensor:
    """y = act(x W^T + b); x [M, K] (row stride may exceed K), W [N, K] contiguous (nn.Linear layout)."""
    _need_cuda(x, W)
    M, ldx = _rows_ld(x)
    N, K = W.shape
    W = _f32c(W, "W")
    y = torch.empty((M, N), dtype=torch.float32, device=x.device)
    L = _lib.load()
    _lib.check(L.b2r_linear_fwd(_p(x), ldx, _p(W), _p(bias), _p(y), N, M, N, K, 1 if relu else 0, _stream()),
               "b2r_linear_fwd")
    return y


def linear_fwd_tc(x: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor], relu: bool, strict: bool = True):
    """same as linear_fwd on the tcgen05 tensor cores (TF32 hi/lo split, fp32-class accuracy); for shapes outside the
    kernel's class it raises (strict) or returns None"""
    _need_cuda(x, W)
    M, ldx = _rows_ld(x)
    N, K = W.shape
    W = _f32c(W, "W")
    y = torch.empty((M, N), dtype=torch.float32, device=x.device)
    L = _lib.load()
    rc = L.b2r_linear_fwd_tc(_p(x), ldx, _p(W), _p(bias), _p(y), N, M, N, K, 1 if relu else 0, _stream())
    if rc == -3 and not strict:
        return None
    _lib.check(rc, "b2r_linear_fwd_tc")
    return y


def linear_bwd(dy: torch.Tensor, x: torch.Tensor, W: torch.Tensor, y_relu: Optional[torch.Tensor],
               need_dx: bool, need_dw: bool, need_db: bool):
    """gradients of y = act(x W^T + b) given dy; y_relu = saved post-ReLU output (None when no ReLU)."""
    M, lddy = _rows_ld(dy)
    _, ldx = _rows_ld(x)
    N, K = W.shape
    L = _lib.load()
    dx = dW = db = None
    if y_relu is not None and _rows_ld(y_relu)[1] != lddy:
        y_relu = y_relu.contiguous()
        dy = dy.contiguous()
        lddy = N
    if need_dx:
        dx = torch.empty((M, K), dtype=torch.float32, device=dy.device)
        done = False
        if _TC_LINEAR and M >= 4096 and N % 32 == 0 and N <= 128 and K % 16 == 0 and K <= 256 and lddy % 4 == 0:
            # dX = (dY o mask) W = (dY o mask) (W^T)^T on the tcgen05 kernel: the transposed weight (N*K floats) is a copy
            Wt = W.t().contiguous()
            rc = L.b2r_linear_tc(_p(dy), lddy, _p(y_relu), _p(Wt), None, _p(dx), K, M, K, N, 0, _stream())
            done = rc == 0
            if rc not in (0, -3):
                _lib.check(rc, "b2r_linear_tc")
        if not done:
            _lib.check(L.b2r_linear_bwd_input(_p(dy), lddy, _p(y_relu), _p(W), _p(dx), K, M, N, K, _stream()),
                       "b2r_linear_bwd_input")
    if need_dw or need_db:
        dW = torch.empty((N, K), dtype=torch.float32, device=dy.device)
        db = torch.empty((N,), dtype=torch.float32, device=dy.device) if need_db else None
        if _TC_DW and M >= 4096 and N % 4 == 0 and N <= 128 and K % 16 == 0 and K <= 240 and lddy % 4 == 0 and ldx % 4 == 0:
            # long reductions (M = B*L rows) on the tcgen05 tensor cores: k_linear_dw_tc
            ws = _ws(L.b2r_linear_bwd_weight_tc_workspace_bytes(M, N, K), dy.device)
            rc = L.b2r_linear_bwd_weight_tc(_p(dy), lddy, _p(y_relu), _p(x), ldx, _p(dW), _p(db), M, N, K, _p(ws), ws.numel(),
                                            _stream())
            if rc == 0:
                return dx, dW, db
            if rc != -3:                                   # B2R_E_UNSUPPORTED falls through to the CUDA-core kernel
                _lib.check(rc, "b2r_linear_bwd_weight_tc")
        nbytes = L.b2r_linear_bwd_weight_workspace_bytes(M, N, K)
        ws = _ws(nbytes, dy.device)
        _lib.check(L.b2r_linear_bwd_weight(_p(dy), lddy, _p(y_relu), _p(x), ldx, _p(dW), _p(db), M, N, K, _p(ws),
                                           ws.numel(), _stream()), "b2r_linear_bwd_weight")
    return dx, dW, db


# Linear forward and dX with a long batch dimension (M >= 4096) run on the tcgen05 kernels (csrc/linear_tc.cu: TF32 hi/lo
# split, 3 products, fp32-class accuracy checked in tests/test_gpu_tc.py; the pipelined K <= 64 kernel is ~2x the SGEMM at
# M = 204800, N = K = 64); B2R_TC_LINEAR=0: CUDA-core SGEMM everywhere
_TC_LINEAR = os.environ.get("B2R_TC_LINEAR", "1") != "0"
# weight gradients with a long batch reduction run on the tensor cores (csrc/linear_dw_tc.cu); B2R_TC_DW=0: CUDA cores
_TC_DW = os.environ.get("B2R_TC_DW", "1") != "0"


class _Linear(torch.autograd.Function):
    """nn.Linear (+ optional ReLU) on the library's SGEMM: NeuMF.py:70 / layers.py:26-28,106-107."""

    @staticmethod
    def forward(ctx, x, W, bias, relu):
        shape = x.shape
        x2 = x.reshape(-1, shape[-1])
        if x2.stride(1) != 1:
            x2 = x2.contiguous()
        N, K = W.shape
        if _TC_LINEAR and x2.shape[0] >= 4096 and K % 32 == 0 and K <= 128 and N % 16 == 0 and N <= 256:
            y = linear_fwd_tc(x2, W, bias, relu, strict=False)          # tcgen05 TF32-split forward
        else:
            y = None
        if y is None:
            y = linear_fwd(x2, W, bias, relu)
        ctx.save_for_backward(x2, W, y if relu else None)
        ctx.has_bias, ctx.shape = bias is not None, shape
        return y.view(*shape[:-1], W.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, W, y_relu = ctx.saved_tensors
        dy2 = _f32c(dy.reshape(-1, W.shape[0]), "dy")
        dx, dW, db = linear_bwd(dy2, x2, W, y_relu, ctx.needs_input_grad[0], ctx.needs_input_grad[1],
                                ctx.has_bias and ctx.needs_input_grad[2])
        if dx is not None:
            dx = dx.view(*ctx.shape)
        return dx, dW, db, None


def linear(x, W, bias=None, relu=False):
    return _Linear.apply(x, W, bias, relu)


class _AddLayerNorm(torch.autograd.Function):
    """LayerNorm(x + res) with affine (layers.py:113,117)."""

    @staticmethod
    def forward(ctx, x, res, gamma, beta, eps):
        _need_cuda(x, res, gamma, beta)
        d = x.shape[-1]
        x2, r2 = _f32c(x.reshape(-1, d), "x"), _f32c(res.reshape(-1, d), "res")
        rows = x2.shape[0]
        y = torch.empty_like(x2)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        L = _lib.load()
        _lib.check(L.b2r_add_layernorm_fwd(_p(x2), _p(r2), _p(gamma), _p(beta), _p(y), _p(mean), _p(rstd), rows, d,
                                           float(eps), _stream()), "b2r_add_layernorm_fwd")
        ctx.save_for_backward(x2, r2, gamma, mean, rstd)
        ctx.shape = x.shape
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, r2, gamma, mean, rstd = ctx.saved_tensors
        d = x2.shape[1]
        rows = x2.shape[0]
        dy2 = _f32c(dy.reshape(-1, d), "dy")
        dz = torch.empty_like(x2)
        dg = torch.empty(d, dtype=torch.float32, device=dy.device)
        db = torch.empty(d, dtype=torch.float32, device=dy.device)
        L = _lib.load()
        ws = _ws(L.b2r_add_layernorm_bwd_workspace_bytes(rows, d), dy.device)
        _lib.check(L.b2r_add_layernorm_bwd(_p(dy2), _p(x2), _p(r2), _p(gamma), _p(mean), _p(rstd), _p(dz), _p(dg),
                                           _p(db), rows, d, _p(ws), ws.numel(), _stream()), "b2r_add_layernorm_bwd")
        dz = dz.view(ctx.shape)
        return dz, dz, dg, db, None


def add_layernorm(x, res, gamma, beta, eps=1e-5):
    return _AddLayerNorm.apply(x, res, gamma, beta, eps)


# B2R_ATTN=v1: the first (shared-memory operand) attention kernels for every shape, for A/B
_ATTN_RT = os.environ.get("B2R_ATTN", "rt") != "v1"


class _CausalAttention(torch.autograd.Function):
    """layers.py:52-63 for q,k,v [B, L, d] with H heads (contiguous d/H chunks), causal mask, no W_o.  ``live`` (optional
    int64 [B]): positions >= live[b] are dead rows (nothing downstream reads them): skipped, written as zeros."""

    @staticmethod
    def forward(ctx, q, k, v, H, live):
        _need_cuda(q, k, v, live)
        B, Ln, d = q.shape
        q, k, v = _f32c(q, "q"), _f32c(k, "k"), _f32c(v, "v")
        live = None if live is None else _i64c(live, "live lengths")
        out = torch.empty((B, Ln, d), dtype=torch.float32, device=q.device)
        L = _lib.load()
        lse = None
        if _ATTN_RT and d % H == 0 and d // H in (8, 16, 32) and Ln <= 128:
            # register-resident kernels (csrc/attention_rt.cu); the row log-sum-exp is kept for the backward
            lse = torch.empty((B, Ln, H), dtype=torch.float32, device=q.device)
            rc = L.b2r_attention_fwd_rt(_p(q), _p(k), _p(v), d, _p(live), _p(out), _p(lse), B, Ln, d, H, _stream())
            if rc == -3:                                   # B2R_E_UNSUPPORTED: shape / alignment outside the fast path
                lse = None
            else:
                _lib.check(rc, "b2r_attention_fwd_rt")
        if lse is None:
            _lib.check(L.b2r_attention_fwd_live(_p(q), _p(k), _p(v), d, _p(live), _p(out), B, Ln, d, H, _stream()),
                       "b2r_attention_fwd")
        ctx.save_for_backward(q, k, v, live, out if lse is not None else None, lse)
        ctx.H = H
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, live, out, lse = ctx.saved_tensors
        B, Ln, d = q.shape
        dout = _f32c(dout, "dctx")
        dq, dk, dv = torch.empty_like(q), torch.empty_like(q), torch.empty_like(q)
        L = _lib.load()
        if lse is not None:
            _lib.check(L.b2r_attention_bwd_rt(_p(q), _p(k), _p(v), d, _p(live), _p(out), _p(lse), _p(dout), _p(dq), _p(dk),
                                              _p(dv), d, B, Ln, d, ctx.H, _stream()), "b2r_attention_bwd_rt")
        else:
            _lib.check(L.b2r_attention_bwd_live(_p(q), _p(k), _p(v), d, _p(live), _p(dout), _p(dq), _p(dk), _p(dv), d, B, Ln,
                                                d, ctx.H, _stream()), "b2r_attention_bwd")
        return dq, dk, dv, None, None


def causal_attention(q, k, v, num_heads, live=None):
    return _CausalAttention.apply(q, k, v, num_heads, live)


class _AttentionLast(torch.autograd.Function):
    """Attention for the one query per sequence that SASRec's last block needs (position len-1): q_last [B, d] against
    k, v [B, L, d]; identical to row len-1 of _CausalAttention.  Opt-in groundwork (B2R_SASREC_LASTQ=1)."""

    @staticmethod
    def forward(ctx, q_last, k, v, lengths, H):
        _need_cuda(q_last, k, v, lengths)
        B, Ln, d = k.shape
        q_last, k, v = _f32c(q_last, "q_last"), _f32c(k, "k"), _f32c(v, "v")
        lengths = _i64c(lengths, "lengths")
        out = torch.empty((B, d), dtype=torch.float32, device=k.device)
        prob = torch.empty((B, H, Ln), dtype=torch.float32, device=k.device)
        _lib.check(_lib.load().b2r_attention_last_fwd(_p(q_last), _p(k), _p(v), d, _p(lengths), _p(out), _p(prob), B, Ln,
                                                      d, H, _stream()), "b2r_attention_last_fwd")
        ctx.save_for_backward(q_last, k, v, lengths, prob)
        ctx.H = H
        return out

    @staticmethod
    def backward(ctx, dout):
        q_last, k, v, lengths, prob = ctx.saved_tensors
        B, Ln, d = k.shape
        dout = _f32c(dout, "dctx_last")
        dq, dk, dv = torch.empty_like(q_last), torch.empty_like(k), torch.empty_like(k)
        _lib.check(_lib.load().b2r_attention_last_bwd(_p(q_last), _p(k), _p(v), d, _p(lengths), _p(prob), _p(dout), _p(dq),
                                                      _p(dk), _p(dv), d, B, Ln, d, ctx.H, _stream()),
                   "b2r_attention_last_bwd")
        return dq, dk, dv, None, None


def attention_last(q_last, k, v, lengths, num_heads):
    return _AttentionLast.apply(q_last, k, v, lengths, num_heads)


class _EmbedHistory(torch.autograd.Function):
    """x = I[hist] + P[(len - t) * valid]  (SASRec.py:58-66); backward: row-sparse scatter into I (padding
    positions dropped: their gradient is exactly zero) and the dense small-table gradient of P."""

    @staticmethod
    def forward(ctx, I, P, hist, lengths):
        _need_cuda(I, P, hist, lengths)
        hist, lengths = _i64c(hist, "history_items"), _i64c(lengths, "lengths")
        B, Ln = hist.shape
        d = I.shape[1]
        x = torch.empty((B, Ln, d), dtype=torch.float32, device=I.device)
        pos = torch.empty((B, Ln), dtype=torch.int64, device=I.device)
        L = _lib.load()
        _lib.check(L.b2r_embed_history(_p(I), I.shape[0], _p(P), P.shape[0], _p(hist), _p(lengths), _p(x), _p(pos),
                                       B, Ln, d, _p(err_flag(I.device)), _stream()), "b2r_embed_history")
        ctx.save_for_backward(hist, pos)
        ctx.I, ctx.P = I, P
        return x

    @staticmethod
    def backward(ctx, dx):
        hist, pos = ctx.saved_tensors
        I, P = ctx.I, ctx.P
        d = I.shape[1]
        dx2 = _f32c(dx.reshape(-1, d), "dx")
        gI = gP = None
        if ctx.needs_input_grad[0]:
            gI = _table_grad(I, hist, Source(src=dx2, n=hist.numel()), ignore_id=0)
        if ctx.needs_input_grad[1]:
            n = pos.numel()
            gP = torch.empty_like(P)
            L = _lib.load()
            ws = _ws(L.b2r_small_table_grad_workspace_bytes(n, P.shape[0], d), dx.device)
            _lib.check(L.b2r_small_table_grad(_p(dx2), d, _p(pos), n, P.shape[0], d, _p(gP), _p(ws), ws.numel(),
                                              _stream()), "b2r_small_table_grad")
        return gI, gP, None, None


def embed_history(I, P, hist, lengths):
    return _EmbedHistory.apply(I, P, hist, lengths)


class _SelectLast(torch.autograd.Function):
    """h[b] = y[b, len[b]-1] * valid  (SASRec.py:74-76)."""

    @staticmethod
    def forward(ctx, y, hist, lengths):
        _need_cuda(y, hist, lengths)
        y = _f32c(y, "y")
        B, Ln, d = y.shape
        h = torch.empty((B, d), dtype=torch.float32, device=y.device)
        L = _lib.load()
        _lib.check(L.b2r_select_last(_p(y), _p(hist), _p(lengths), _p(h), B, Ln, d, _stream()), "b2r_select_last")
        ctx.save_for_backward(hist, lengths)
        ctx.shape = (B, Ln, d)
        return h

    @staticmethod
    def backward(ctx, dh):
        hist, lengths = ctx.saved_tensors
        B, Ln, d = ctx.shape
        dy = torch.empty((B, Ln, d), dtype=torch.float32, device=dh.device)
        L = _lib.load()
        _lib.check(L.b2r_select_last_bwd(_p(_f32c(dh, "dh")), _p(hist), _p(lengths), _p(dy), B, Ln, d, _stream()),
                   "b2r_select_last_bwd")
        return dy, None, None


def select_last(y, hist, lengths):
    return _SelectLast.apply(y, _i64c(hist, "history_items"), _i64c(lengths, "lengths"))


class _ColScale(torch.autograd.Function):
    """out[r,k] = a[r,k] * w[k]  (the GMF half of NeuMF's bias-free output layer, NeuMF.py:68,74-75)."""

    @staticmethod
    def forward(ctx, a, w):
        _need_cuda(a, w)
        a, w = _f32c(a, "a"), _f32c(w, "w")
        out = torch.empty_like(a)
        L = _lib.load()
        _lib.check(L.b2r_colscale(_p(a), _p(w), _p(out), a.shape[0], a.shape[1], _stream()), "b2r_colscale")
        ctx.save_for_backward(a, w)
        return out

    @staticmethod
    def backward(ctx, g):
        a, w = ctx.saved_tensors
        g = _f32c(g, "g")
        L = _lib.load()
        da = dw = None
        if ctx.needs_input_grad[0]:
            da = torch.empty_like(a)
            _lib.check(L.b2r_colscale(_p(g), _p(w), _p(da), a.shape[0], a.shape[1], _stream()), "b2r_colscale")
        if ctx.needs_input_grad[1]:
            dw = torch.empty_like(w)
            _lib.check(L.b2r_colsum_prod(_p(g), _p(a), _p(dw), a.shape[0], a.shape[1], _stream()), "b2r_colsum_prod")
        return da, dw


def colscale(a, w):
    return _ColScale.apply(a, w)


class _GatherConcat(torch.autograd.Function):
    """x[p, :] = [ Tu[uid[p // C]] ; Ti[iid[p]] ]  (NeuMF.py:61-69: repeated user ids + concat), written by two
    strided gathers into one [B*C, 2d] buffer; backward: one row-sparse scatter per table reading its column
    block of the dense gradient in place."""

    @staticmethod
    def forward(ctx, Tu, Ti, uid, iid):
        _need_cuda(Tu, Ti, uid, iid)
        uid, iid = _i64c(uid, "user_id"), _i64c(iid, "item_id")
        B, Cn = iid.shape
        d = Tu.shape[1]
        n = B * Cn
        x = torch.empty((n, 2 * d), dtype=torch.float32, device=Tu.device)
        L = _lib.load()
        ef = _p(err_flag(Tu.device))
        _lib.check(L.b2r_gather_rows_strided(_p(Tu), _p(uid), Tu.shape[0], _p(x), 2 * d, n, d, Cn, ef, _stream()),
                   "b2r_gather_rows_strided")
        _lib.check(L.b2r_gather_rows_strided(_p(Ti), _p(iid), Ti.shape[0], x.data_ptr() + 4 * d, 2 * d, n, d, 1, ef,
                                             _stream()), "b2r_gather_rows_strided")
        ctx.save_for_backward(uid, iid)
        ctx.Tu, ctx.Ti = Tu, Ti
        return x

    @staticmethod
    def backward(ctx, dx):
        uid, iid = ctx.saved_tensors
        Tu, Ti = ctx.Tu, ctx.Ti
        B, Cn = iid.shape
        d = Tu.shape[1]
        n = B * Cn
        dx = _f32c(dx, "dx")
        gu = gi = None
        if ctx.needs_input_grad[0]:
            # one position per sample-candidate pair, id = uid[p // C]: materialise the repeated ids for the plan
            rep = uid.repeat_interleave(Cn)
            gu = _table_grad(Tu, rep, Source(src=dx, n=n, ld=2 * d))
        if ctx.needs_input_grad[1]:
            gi = _table_grad(Ti, iid, Source(src=dx[:, d:], n=n, ld=2 * d))
        return gu, gi, None, None


def gather_concat(Tu, Ti, uid, iid):
    return _GatherConcat.apply(Tu, Ti, uid, iid)
