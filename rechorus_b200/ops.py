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

    def c_struct(self) -> _lib.GradSource:
        return _lib.GradSource(_p(self.src), _p(self.coef), _p(self.src_id), self.n, self.div, 0)


class IndexPlan:
    """Sorted, de-duplicated view of the ids a batch touches in one table (b2r_plan_build)."""

    def __init__(self, ids: torch.Tensor, n_rows: int):
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
        _lib.check(L.b2r_plan_build(_p(ids), n, self.n_rows, _p(self.sorted_key), _p(self.sorted_pos),
                                    _p(self.seg_start), _p(self.n_uniq), _p(ws), nbytes, _p(err_flag(dev)),
                                    _stream()), "b2r_plan_build")
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
                                       _p(self.n_uniq), self.n, d, C.byref(s0),
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
        return uniq[:nu], rows[:nu]

    def add_to_dense(self, dense: torch.Tensor, sources: Sequence[Source]) -> None:
        self._apply(dense.shape[1], sources, 1, dense=dense)

    def apply_optimizer(self, W: torch.Tensor, m, v, opt: _lib.Optim, sources: Sequence[Source]) -> None:
        self._apply(W.shape[1], sources, 2, W=W, m=m, v=v, opt=opt)


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


def _table_grad(table: torch.Tensor, ids: torch.Tensor, source: Source) -> Optional[torch.Tensor]:
    mode = table_mode(table)
    if mode == "fused":
        table._b2r_pending.append((ids.reshape(-1), source))
        return None
    plan = IndexPlan(ids, table.shape[0])
    if mode == "sparse":
        uniq, rows = plan.reduce_rows(table.shape[1], [source])
        return torch.sparse_coo_tensor(uniq.unsqueeze(0), rows, size=tuple(table.shape), is_coalesced=True)
    dense = torch.zeros_like(table)
    plan.add_to_dense(dense, [source])
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


def embedding(table: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
    return _Gather.apply(table, ids)


def score(q: torch.Tensor, table: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
    return _Score.apply(q, table, ids)


def bpr_loss(pred: torch.Tensor) -> torch.Tensor:
    return _BprLoss.apply(pred)


# --------------------------------------------------------------------------------------------------
# whole-step entry points (one C call per training step)
# --------------------------------------------------------------------------------------------------

_step_ws = {}


def bprmf_train_step(U: torch.nn.Parameter, I: torch.nn.Parameter, optimizer, uid: torch.Tensor,
                     iid: torch.Tensor) -> torch.Tensor:
    """b2r_bprmf_train_step: forward + BPR loss + backward + fused row-sparse optimizer, in place."""
    from .optim import RowSparseOptimizer
    if not isinstance(optimizer, RowSparseOptimizer):
        raise _lib.B200RecError("train_step needs model.optimizer to be a rechorus_b200.optim.RowSparseOptimizer")
    _need_cuda(U, I, uid, iid)
    uid, iid = _i64c(uid, "user_id"), _i64c(iid, "item_id")
    B, Cn = iid.shape
    d = U.shape[1]
    eu, ei = optimizer.entry(U), optimizer.entry(I)
    if eu["wd"] != ei["wd"]:
        raise _lib.B200RecError("fused step expects one weight decay for both tables")
    L = _lib.load()
    key = (U.device.index, B, Cn, d, U.shape[0], I.shape[0])
    ws = _step_ws.get(key)
    if ws is None:
        nbytes = L.b2r_bprmf_step_workspace_bytes(B, Cn, d, U.shape[0], I.shape[0])
        if nbytes == 0:
            raise _lib.B200RecError("b2r_bprmf_step_workspace_bytes returned 0")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=U.device)
        _step_ws[key] = ws
    optimizer.advance()
    opt = optimizer._opt(ei["wd"])
    tables = _lib.BprmfTables(_p(U.data), _p(I.data), _p(eu["m"]), _p(eu["v"]), _p(ei["m"]), _p(ei["v"]),
                              U.shape[0], I.shape[0], d, 0)
    loss = torch.empty((), dtype=torch.float32, device=U.device)
    _lib.check(L.b2r_bprmf_train_step(C.byref(tables), _p(uid), _p(iid), B, Cn, C.byref(opt), _p(loss), _p(ws),
                                      ws.numel(), _p(err_flag(U.device)), _stream()), "b2r_bprmf_train_step")
    return loss
