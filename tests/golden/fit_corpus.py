"""The tiny synthetic corpus shared by tests/golden/make_fit_golden.py (which runs the REFERENCE's BaseRunner on it)
and tests/test_gpu_fit_golden.py (which runs this repository's runner on it).  Attribute names are the ones
helpers/BaseReader.py:32-60 and helpers/SeqReader.py:21-36 produce."""
import types

import numpy as np
import pandas as pd

N_USERS, N_ITEMS, ROWS, N_EVAL, N_NEG_EVAL = 30, 50, 400, 48, 19


def build():
    rng = np.random.RandomState(20260923)
    df = pd.DataFrame({"user_id": rng.randint(1, N_USERS, ROWS), "item_id": rng.randint(1, N_ITEMS, ROWS),
                       "time": np.arange(ROWS)})
    # SeqReader.py:21-36: position = number of earlier interactions of the user; user_his = [(item, time), ...]
    user_his, position = {}, []
    for u, i, t in zip(df.user_id, df.item_id, df.time):
        h = user_his.setdefault(int(u), [])
        position.append(len(h))
        h.append((int(i), int(t)))
    df["position"] = position
    ev = df.iloc[ROWS - N_EVAL:].copy().reset_index(drop=True)
    ev["neg_items"] = [list(rng.randint(1, N_ITEMS, N_NEG_EVAL)) for _ in range(len(ev))]
    train = df.iloc[:ROWS - N_EVAL].copy().reset_index(drop=True)
    clicked = {u: set() for u in range(N_USERS)}
    for u, i in zip(train.user_id, train.item_id):
        clicked[int(u)].add(int(i))
    residual = {u: set() for u in range(N_USERS)}
    for u, i in zip(ev.user_id, ev.item_id):
        residual[int(u)].add(int(i))
    return types.SimpleNamespace(n_users=N_USERS, n_items=N_ITEMS, data_df={"train": train, "dev": ev, "test": ev},
                                 train_clicked_set=clicked, residual_clicked_set=residual, user_his=user_his)


CASES = {
    "fit_bprmf": ("BPRMF", ["--emb_size", "64", "--num_neg", "3", "--lr", "0.01", "--l2", "1e-5"]),
    "fit_neumf": ("NeuMF", ["--emb_size", "32", "--layers", "[32, 16]", "--num_neg", "2", "--lr", "0.01", "--l2", "1e-5"]),
    "fit_sasrec": ("SASRec", ["--emb_size", "32", "--num_layers", "2", "--num_heads", "2", "--history_max", "8",
                              "--num_neg", "2", "--lr", "0.005", "--l2", "1e-6"]),
}
COMMON = ["--batch_size", "64", "--eval_batch_size", "32", "--num_workers", "0", "--topk", "5,10", "--metric", "NDCG,HR"]
EPOCHS = 2
