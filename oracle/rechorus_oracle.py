"""CPU oracle for the ReChorus training hot path.  TEST INFRASTRUCTURE -- NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module, and only as the *checker* (or as the timed CPU baseline).  Nothing under
``rechorus_b200/`` imports it; the product path raises if the CUDA library is missing.

What it is: a functional (weights-in, scores-out) restatement, on plain CPU PyTorch ops, of the arithmetic
the reference performs on this path.  Every function cites the reference file:line it restates
(paths relative to the reference checkout's ``src/``).

Parity pin: the reference has no tests or golden vectors of its own.  The oracle is
pinned against outputs of the *reference itself*, executed in the build container by
``tests/golden/make_golden.py`` (imports the unmodified reference classes, seeds them, dumps
inputs/weights/outputs as ``tests/golden/*.npz``).  ``tests/test_oracle_golden.py`` replays those
fixtures through this file on every CPU test run.

All arithmetic in the reference is third-party ATen (``requirements.txt:2`` pins torch==1.12.1); the
oracle runs on the installed torch, dtype-generic so an fp64 replay can adjudicate 1e-5 disputes.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Params = Dict[str, Tensor]

# --------------------------------------------------------------------------------------------------
# parameter construction (models/BaseModel.py:29-35 init_weights; BPRMF.py:30-32; NeuMF.py:42-54;
# SASRec.py:41-49; utils/layers.py:26-28,103-109)
# --------------------------------------------------------------------------------------------------

INIT_STD = 0.01  # models/BaseModel.py:31-35: every Linear weight, Linear bias and Embedding ~ N(0, 0.01^2)


def _normal(shape: Sequence[int], gen: torch.Generator, dtype=torch.float32) -> Tensor:
    return (torch.randn(*shape, generator=gen, dtype=torch.float64) * INIT_STD).to(dtype)


def bprmf_init(n_users: int, n_items: int, d: int, gen: torch.Generator, dtype=torch.float32) -> Params:
    """State-dict keys of models/general/BPRMF.py:30-32."""
    return {
        "u_embeddings.weight": _normal((n_users, d), gen, dtype),
        "i_embeddings.weight": _normal((n_items, d), gen, dtype),
    }


def neumf_init(n_users: int, n_items: int, d: int, layers: Sequence[int], gen: torch.Generator,
               dtype=torch.float32) -> Params:
    """State-dict keys of models/general/NeuMF.py:42-54 (``prediction`` has no bias, :54)."""
    p: Params = {
        "mf_u_embeddings.weight": _normal((n_users, d), gen, dtype),
        "mf_i_embeddings.weight": _normal((n_items, d), gen, dtype),
        "mlp_u_embeddings.weight": _normal((n_users, d), gen, dtype),
        "mlp_i_embeddings.weight": _normal((n_items, d), gen, dtype),
    }
    fan_in = 2 * d
    for l, width in enumerate(layers):
        p[f"mlp.{l}.weight"] = _normal((width, fan_in), gen, dtype)   # nn.Linear stores [out, in]
        p[f"mlp.{l}.bias"] = _normal((width,), gen, dtype)
        fan_in = width
    p["prediction.weight"] = _normal((1, fan_in + d), gen, dtype)
    return p


def sasrec_init(n_items: int, d: int, history_max: int, num_layers: int, gen: torch.Generator,
                dtype=torch.float32) -> Params:
    """State-dict keys of models/sequential/SASRec.py:41-49 + utils/layers.py:26-28,103-109.
    LayerNorm keeps its (1, 0) default affine (init_weights only touches Linear/Embedding)."""
    p: Params = {
        "i_embeddings.weight": _normal((n_items, d), gen, dtype),
        "p_embeddings.weight": _normal((history_max + 1, d), gen, dtype),
    }
    for b in range(num_layers):
        pre = f"transformer_block.{b}."
        for name in ("q_linear", "k_linear", "v_linear"):
            p[pre + f"masked_attn_head.{name}.weight"] = _normal((d, d), gen, dtype)
            p[pre + f"masked_attn_head.{name}.bias"] = _normal((d,), gen, dtype)
        for ln in ("layer_norm1", "layer_norm2"):
            p[pre + ln + ".weight"] = torch.ones(d, dtype=dtype)
            p[pre + ln + ".bias"] = torch.zeros(d, dtype=dtype)
        for lin in ("linear1", "linear2"):   # d_ff == d_model (SASRec.py:46)
            p[pre + lin + ".weight"] = _normal((d, d), gen, dtype)
            p[pre + lin + ".bias"] = _normal((d,), gen, dtype)
    return p


def param_groups(named: Iterable[Tuple[str, Tensor]], l2: float) -> List[dict]:
    """models/BaseModel.py:64-73 + helpers/BaseRunner.py:110-114: names containing 'bias' get
    weight_decay 0, everything else the runner's --l2."""
    decay, no_decay = [], []
    for name, t in named:
        (no_decay if "bias" in name else decay).append(t)
    return [{"params": decay, "weight_decay": l2}, {"params": no_decay, "weight_decay": 0.0}]


# --------------------------------------------------------------------------------------------------
# forward passes
# --------------------------------------------------------------------------------------------------

def bprmf_scores(p: Params, user_id: Tensor, item_id: Tensor) -> Tensor:
    """models/general/BPRMF.py:34-45: pred[b,c] = <U[user_id[b]], I[item_id[b,c]]>."""
    u = F.embedding(user_id, p["u_embeddings.weight"])            # [B, d]
    i = F.embedding(item_id, p["i_embeddings.weight"])            # [B, C, d]
    return torch.einsum("bd,bcd->bc", u, i)


def neumf_scores(p: Params, user_id: Tensor, item_id: Tensor) -> Tensor:
    """models/general/NeuMF.py:56-76 with dropout 0 (BaseModel.py:161 default):
    GMF branch u*i, MLP tower on [u ; i] (user first), final bias-free Linear on [mf ; h] (mf first)."""
    B, C = item_id.shape
    uid = user_id.view(B, 1).expand(B, C)
    mf = F.embedding(uid, p["mf_u_embeddings.weight"]) * F.embedding(item_id, p["mf_i_embeddings.weight"])
    h = torch.cat([F.embedding(uid, p["mlp_u_embeddings.weight"]),
                   F.embedding(item_id, p["mlp_i_embeddings.weight"])], dim=-1)
    l = 0
    while f"mlp.{l}.weight" in p:
        h = torch.relu(F.linear(h, p[f"mlp.{l}.weight"], p[f"mlp.{l}.bias"]))
        l += 1
    out = F.linear(torch.cat([mf, h], dim=-1), p["prediction.weight"])
    return out.reshape(B, C)


def _attention_block(x: Tensor, p: Params, pre: str, num_heads: int) -> Tensor:
    """utils/layers.py:112-118 (TransformerLayer, post-LN, d_ff = d, ReLU) around
    utils/layers.py:34-63 (MultiHeadAttention: q/k/v Linear with bias, contiguous d_k head chunks,
    causal mask -> -inf, softmax shifted by the max over the WHOLE score tensor, NaN -> 0,
    heads concatenated, NO output projection)."""
    B, L, d = x.shape
    dk = d // num_heads

    def heads(name: str) -> Tensor:
        y = F.linear(x, p[pre + f"masked_attn_head.{name}.weight"], p[pre + f"masked_attn_head.{name}.bias"])
        return y.view(B, L, num_heads, dk).permute(0, 2, 1, 3)    # [B, h, L, dk]

    q, k, v = heads("q_linear"), heads("k_linear"), heads("v_linear")
    s = (q @ k.transpose(-1, -2)) / dk ** 0.5                      # layers.py:57
    keep = torch.ones(L, L, dtype=torch.bool).tril()               # SASRec.py:69-70
    s = s.masked_fill(~keep, float("-inf"))
    a = torch.softmax(s - s.max(), dim=-1)                         # layers.py:60 (global max)
    a = torch.where(torch.isnan(a), torch.zeros_like(a), a)        # layers.py:61
    ctx = (a @ v).permute(0, 2, 1, 3).reshape(B, L, d)             # layers.py:49 (no W_o)
    c = F.layer_norm(ctx + x, (d,), p[pre + "layer_norm1.weight"], p[pre + "layer_norm1.bias"], 1e-5)
    o = F.linear(torch.relu(F.linear(c, p[pre + "linear1.weight"], p[pre + "linear1.bias"])),
                 p[pre + "linear2.weight"], p[pre + "linear2.bias"])
    return F.layer_norm(o + c, (d,), p[pre + "layer_norm2.weight"], p[pre + "layer_norm2.bias"], 1e-5)


def sasrec_user_state(p: Params, history: Tensor, lengths: Tensor, num_heads: int) -> Tensor:
    """models/sequential/SASRec.py:51-76: item + reversed-position embedding, causal blocks,
    zero the padded positions, take the state at lengths-1.  Returns [B, d]."""
    B, L = history.shape
    valid = (history > 0).to(torch.long)
    pos = (lengths.view(B, 1) - torch.arange(L).view(1, L)) * valid        # SASRec.py:64
    x = F.embedding(history, p["i_embeddings.weight"]) + F.embedding(pos, p["p_embeddings.weight"])
    b = 0
    while f"transformer_block.{b}.linear1.weight" in p:
        x = _attention_block(x, p, f"transformer_block.{b}.", num_heads)
        b += 1
    x = x * valid.unsqueeze(-1).to(x.dtype)                                # SASRec.py:74
    return x[torch.arange(B), lengths - 1]                                 # SASRec.py:76


def sasrec_scores(p: Params, history: Tensor, lengths: Tensor, item_id: Tensor, num_heads: int) -> Tensor:
    """models/sequential/SASRec.py:80-81: candidates scored by a dot with the user state."""
    h = sasrec_user_state(p, history, lengths, num_heads)
    return torch.einsum("bd,bcd->bc", h, F.embedding(item_id, p["i_embeddings.weight"]))


# --------------------------------------------------------------------------------------------------
# loss (models/BaseModel.py:175-189) and its closed-form gradient
# --------------------------------------------------------------------------------------------------

def bpr_loss(pred: Tensor) -> Tensor:
    """models/BaseModel.py:182-185.  Column 0 is the positive.  The negatives' softmax weights are NOT
    detached; the shift is by the max over the whole negative block; clamp to [1e-8, 1-1e-8] before log."""
    pos, neg = pred[:, :1], pred[:, 1:]
    w = torch.softmax(neg - neg.max(), dim=1)
    s = (torch.sigmoid(pos - neg) * w).sum(dim=1)
    return -torch.log(s.clamp(min=1e-8, max=1 - 1e-8)).mean()


def bpr_loss_and_grad_fp64(pred: np.ndarray) -> Tuple[float, np.ndarray]:
    """Closed form of d loss / d pred in float64 numpy (no autograd), used to check the fused loss
    kernel independently of torch: dS/dp = sum_j w_j s_j (1-s_j);  dS/dn_j = -w_j s_j (1-s_j) + w_j (s_j - S);
    dloss/dS = -1/(B S) inside the clamp window, 0 outside."""
    x = np.asarray(pred, dtype=np.float64)
    B = x.shape[0]
    p, n = x[:, :1], x[:, 1:]
    e = np.exp(n - n.max(axis=1, keepdims=True))
    w = e / e.sum(axis=1, keepdims=True)
    s = 1.0 / (1.0 + np.exp(-(p - n)))
    S = (s * w).sum(axis=1, keepdims=True)
    inside = (S >= 1e-8) & (S <= 1 - 1e-8)
    Sc = np.clip(S, 1e-8, 1 - 1e-8)
    loss = float(-np.log(Sc).mean())
    dS = np.where(inside, -1.0 / (B * Sc), 0.0)
    g = np.empty_like(x)
    g[:, :1] = dS * (w * s * (1 - s)).sum(axis=1, keepdims=True)
    g[:, 1:] = dS * (-w * s * (1 - s) + w * (s - S))
    return loss, g


# --------------------------------------------------------------------------------------------------
# runner-side arithmetic: candidate shuffle, rank metrics, negative sampling, optimizer step
# --------------------------------------------------------------------------------------------------

def shuffle_candidates(item_id: Tensor, gen: torch.Generator | None = None) -> Tuple[Tensor, Tensor]:
    """helpers/BaseRunner.py:187-191: per-row random permutation of the candidate columns."""
    perm = torch.argsort(torch.rand(*item_id.shape, generator=gen), dim=-1)
    return torch.gather(item_id, 1, perm), perm


def unshuffle_scores(pred: Tensor, perm: Tensor) -> Tensor:
    """helpers/BaseRunner.py:196-202: restored[b, perm[b,c]] = pred[b,c] (autograd flows through)."""
    return torch.zeros_like(pred).scatter(1, perm, pred)


def gt_rank(pred: np.ndarray) -> np.ndarray:
    """helpers/BaseRunner.py:63: rank of column 0 = #candidates scoring >= it (ties count against it)."""
    pred = np.asarray(pred)
    return (pred >= pred[:, :1]).sum(axis=-1)


def rank_metrics(pred: np.ndarray, topk: Sequence[int], metrics: Sequence[str]) -> Dict[str, float]:
    """helpers/BaseRunner.py:52-78: HR@k = mean(rank<=k); NDCG@k = mean(hit / log2(rank+1))."""
    r = gt_rank(pred)
    out: Dict[str, float] = {}
    for k in topk:
        hit = r <= k
        for m in metrics:
            if m == "HR":
                out[f"HR@{k}"] = float(hit.mean())
            elif m == "NDCG":
                out[f"NDCG@{k}"] = float((hit / np.log2(r + 1)).mean())
            else:
                raise ValueError(f"Undefined evaluation metric: {m}.")
    return out


def test_all_predictions(q: Tensor, item_table: Tensor, target: Tensor,
                         clicked: Sequence[Sequence[int]] | None = None) -> np.ndarray:
    """The test_all protocol for a dot-product model: candidates = [target] + arange(1, n_items)
    (models/BaseModel.py:194-198), scores = (q[:, None, :] * I[candidates]).sum(-1) (BPRMF.py:42, SASRec.py:81),
    then preds[row, clicked item id] = -inf (helpers/BaseRunner.py:244-251; the column index IS the item id)."""
    n_items = item_table.shape[0]
    cand = torch.cat([target.view(-1, 1), torch.arange(1, n_items).view(1, -1).expand(target.numel(), -1)], dim=1)
    pred = (q[:, None, :] * F.embedding(cand, item_table)).sum(dim=-1).numpy().copy()
    if clicked is not None:
        rows, cols = [], []
        for i, items in enumerate(clicked):
            items = list(items)
            rows.extend([i] * len(items))
            cols.extend(items)
        pred[rows, cols] = -np.inf
    return pred


def sample_negatives(user_ids: Sequence[int], clicked: Dict[int, set], n_items: int, num_neg: int,
                     rng: np.random.RandomState) -> np.ndarray:
    """models/BaseModel.py:206-214: uniform over [1, n_items) with rejection against the user's
    training-set clicks, drawn from NumPy's (global in the reference) RandomState in this exact order."""
    neg = rng.randint(1, n_items, size=(len(user_ids), num_neg))
    for r, u in enumerate(user_ids):
        seen = clicked[u]
        for c in range(num_neg):
            while neg[r][c] in seen:
                neg[r][c] = rng.randint(1, n_items)
    return neg


def scores(model: str, p: Params, batch: Dict[str, Tensor], num_heads: int = 4) -> Tensor:
    if model == "BPRMF":
        return bprmf_scores(p, batch["user_id"], batch["item_id"])
    if model == "NeuMF":
        return neumf_scores(p, batch["user_id"], batch["item_id"])
    if model == "SASRec":
        return sasrec_scores(p, batch["history_items"], batch["lengths"], batch["item_id"], num_heads)
    raise ValueError(model)


def loss_and_grads(model: str, p: Params, batch: Dict[str, Tensor], num_heads: int = 4
                   ) -> Tuple[Tensor, Tensor, Params]:
    """forward + loss + autograd backward (helpers/BaseRunner.py:193-205 without the optimizer):
    returns (pred, loss, dense grads keyed like the state dict) -- dense, as ATen's
    embedding_dense_backward produces for nn.Embedding(sparse=False)."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
    pred = scores(model, leaves, batch, num_heads)
    loss = bpr_loss(pred)
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
    return pred.detach(), loss.detach(), grads


# --------------------------------------------------------------------------------------------------
# list-wise losses of the impression models (models/BaseImpressionModel.py:44-128)
# --------------------------------------------------------------------------------------------------

def listwise_loss(pred: Tensor, target: Tensor, loss_n: str, max_pos: int) -> Tensor:
    """ImpressionModel.loss restated row by row (BaseImpressionModel.py:44-128).  pred [B, Cn], target [B, Cn] in
    {1, 0, -1 (padding)}; columns < max_pos are the positive slots (:55-58).  Supported names: every 'BPR*' form except
    'simple' (which returns a vector in the reference, :79-81), 'listnet', 'softmaxCE', 'attention_rank'.  The batch-wide
    maxima the reference subtracts before its softmaxes (:62,65,68,90,91,114,117) cancel and are replaced by row maxima."""
    B, Cn = pred.shape
    valid = target != -1
    col = torch.arange(Cn).unsqueeze(0).expand(B, Cn)
    P, N = valid & (col < max_pos), valid & (col >= max_pos)
    ninf = torch.full_like(pred, float("-inf"))

    def masked_softmax(x, m):
        return torch.softmax(torch.where(m, x, ninf), dim=1)

    if "BPR" in loss_n:
        if "simple" in loss_n:
            raise ValueError("BPR...simple is not back-propagatable in the reference (returns a [B] vector)")
        nw = masked_softmax(pred, N)                                                  # :60-62
        pw = masked_softmax(-pred if "hard" in loss_n else pred, P)                   # :63-68
        diff = pred.unsqueeze(2) - pred.unsqueeze(1)                                  # [B, i, j] = x_i - x_j
        pair = (P.unsqueeze(2) & N.unsqueeze(1)).to(pred.dtype)
        w2 = pw.unsqueeze(2) * nw.unsqueeze(1)
        if "after" in loss_n:                                                         # :73-75
            return (F.softplus(-diff) * w2 * pair).sum(dim=(1, 2)).mean()
        if "before" in loss_n:                                                        # :76-78
            z = (diff * pair * nw.unsqueeze(1)).sum(dim=2) * pw                       # 0 outside P -> softplus(0) = ln 2
            return F.softplus(-z).sum(dim=1).mean()
        return (-(torch.sigmoid(diff) * w2 * pair).sum(dim=(1, 2)).log()).mean()      # :82-85
    have_neg = valid[:, max_pos].to(pred.dtype)                                       # :53
    scale = have_neg / have_neg.sum() * B                                             # :96,109,126
    tw = masked_softmax(target.to(pred.dtype), valid)                                 # :89-90,113-114
    if loss_n == "listnet":                                                           # :88-97: softmax over ALL columns
        logp = torch.log_softmax(pred, dim=1)
        return (-(tw * torch.where(valid, logp, torch.zeros_like(logp))).sum(dim=1) * scale).mean()
    p = masked_softmax(pred, valid)
    if loss_n == "softmaxCE":                                                         # :99-110
        pos_len = (target == 1).sum(dim=1).to(pred.dtype)
        lp = torch.where(P, p, torch.ones_like(p)).log()
        return (-(lp.sum(dim=1) / pos_len) * scale).mean()
    if loss_n == "attention_rank":                                                    # :112-128
        l1 = -(tw * torch.where(valid, p, torch.ones_like(p)).log()).sum(dim=1)
        p2 = torch.where(valid & (p != 1), p, torch.zeros_like(p))
        l2 = -((1 - tw) * (1 - p2).log()).sum(dim=1)
        return ((l1 + l2) * scale).mean()
    raise ValueError("Undefined loss function: {}".format(loss_n))


class ReferenceStyleTrainer:
    """The reference's per-batch training step as helpers/BaseRunner.py:184-207 performs it on CPU:
    candidate shuffle, zero_grad, forward, un-shuffle, loss, backward with DENSE embedding grads, and a
    stock torch.optim step over every parameter row (helpers/BaseRunner.py:110-114).  This is the leg
    ``bench.py`` times as ``cpu_baseline`` / ``--impl reference`` (kind "port")."""

    def __init__(self, model: str, p: Params, lr: float = 1e-3, l2: float = 0.0,
                 optimizer: str = "Adam", num_heads: int = 4):
        self.model, self.num_heads = model, num_heads
        self.p = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
        self.opt = getattr(torch.optim, optimizer)(param_groups(self.p.items(), l2), lr=lr)

    def step(self, batch: Dict[str, Tensor], shuffle: bool = True) -> float:
        batch = dict(batch)
        if shuffle:
            batch["item_id"], perm = shuffle_candidates(batch["item_id"])
        self.opt.zero_grad()
        pred = scores(self.model, self.p, batch, self.num_heads)
        if shuffle:
            pred = unshuffle_scores(pred, perm)
        loss = bpr_loss(pred)
        loss.backward()
        self.opt.step()
        return float(loss.detach())


class LazyExactAdam:
    """Reference for the NEXT optimizer mode of the kernels: row-sparse bookkeeping with the RESULTS of
    the reference's dense ``torch.optim.Adam`` (helpers/BaseRunner.py:110-114,206).  Dense Adam keeps moving a row it
    has no gradient for (momentum, and g = wd * w under weight decay); the row-sparse kernels of this round skip such
    rows (SparseAdam semantics).  Here every row remembers the step it was last brought up to date and is advanced
    through the skipped steps right before it is read by a forward pass or updated -- after ``flush()`` the table
    equals dense Adam's (tests/test_oracle_golden.py::test_lazy_exact_adam_equals_dense_adam).  Pure-Python loops:
    small cases only."""

    def __init__(self, W: Tensor, lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0):
        self.W = W
        self.m, self.v = torch.zeros_like(W), torch.zeros_like(W)
        self.last = torch.zeros(W.shape[0], dtype=torch.long)
        self.lr, (self.b1, self.b2), self.eps, self.wd = lr, betas, eps, weight_decay
        self.t = 0

    def _one(self, r: int, g: Tensor, t: int) -> None:
        self.m[r] = self.b1 * self.m[r] + (1 - self.b1) * g
        self.v[r] = self.b2 * self.v[r] + (1 - self.b2) * g * g
        m_hat, v_hat = self.m[r] / (1 - self.b1 ** t), self.v[r] / (1 - self.b2 ** t)
        self.W[r] -= self.lr * m_hat / (v_hat.sqrt() + self.eps)

    def _advance(self, rows: Tensor, upto: int) -> None:
        for r in rows.tolist():
            for t in range(int(self.last[r]) + 1, upto + 1):
                self._one(r, self.wd * self.W[r], t)
            self.last[r] = max(int(self.last[r]), upto)

    def read(self, rows: Tensor) -> Tensor:
        """rows as the forward of step t+1 must see them"""
        self._advance(torch.unique(rows), self.t)
        return self.W[rows]

    def step(self, rows: Tensor, grads: Tensor) -> None:
        """rows unique, grads [len(rows), d]: the data gradient (weight decay is added here, like torch.optim.Adam)"""
        self.t += 1
        self._advance(rows, self.t - 1)
        for i, r in enumerate(rows.tolist()):
            self._one(r, grads[i] + self.wd * self.W[r], self.t)
            self.last[r] = self.t

    def flush(self) -> None:
        self._advance(torch.arange(self.W.shape[0]), self.t)


# ------------------------------------------------------------------------------------------------------
# device negative sampler (row f3 groundwork): the definition csrc/sampler.cu implements, restated
# ------------------------------------------------------------------------------------------------------

def philox4x32_10(ctr: Sequence[int], key: Sequence[int]) -> List[int]:
    """Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11; Random123 reference
    constants).  Pinned by the Random123 known-answer vectors in tests/test_oracle_golden.py."""
    M0, M1, W0, W1, MASK = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85, 0xFFFFFFFF
    c, k = list(ctr), list(key)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> 32) & MASK) ^ c[1] ^ k[0], p1 & MASK, ((p0 >> 32) & MASK) ^ c[3] ^ k[1], p0 & MASK]
        k = [(k[0] + W0) & MASK, (k[1] + W1) & MASK]
    return c


def device_sampler_reference(user_ids: Sequence[int], clicked: Dict[int, set], n_items: int, num_neg: int, seed: int,
                             epoch: int) -> np.ndarray:
    """What b2r_sample_negatives must return, bit for bit: uniform over the non-clicked items of [1, n_items)
    (the distribution of models/BaseModel.py:206-214), drawn from the counter-based stream described in
    include/b200rec.h instead of NumPy's global generator: one uniform r over the user's allowed items, mapped to
    the (r+1)-th item of [1, n_items) that is not clicked."""
    out = np.zeros((len(user_ids), num_neg), dtype=np.int64)
    key = [seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF]
    allowed_of: Dict[int, List[int]] = {}
    for i, u in enumerate(user_ids):
        u = int(u)
        if u not in allowed_of:
            seen = clicked.get(u, set())
            assert all(1 <= c < n_items for c in seen)
            allowed_of[u] = [c for c in range(1, n_items) if c not in seen]
        allowed = allowed_of[u]
        for j in range(num_neg):
            idx = i * num_neg + j
            x = philox4x32_10([idx & 0xFFFFFFFF, (idx >> 32) & 0xFFFFFFFF, 0, epoch], key)[0]
            out[i, j] = allowed[(x * len(allowed)) >> 32]
    return out
