"""Load a tests/golden/*.npz fixture (made by tests/golden/make_golden.py from the reference itself)."""
import ast
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

MODEL_FIXTURES = {
    "BPRMF": ["bprmf_k1", "bprmf_k9_trained", "bprmf_d128", "bprmf_d20"],
    "NeuMF": ["neumf_l64", "neumf_l64_32_16", "neumf_d32"],
    "SASRec": ["sasrec_l1h1", "sasrec_l2h4", "sasrec_d32"],
}


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    meta = ast.literal_eval(str(z["meta"]))
    weights = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w:")}
    batch = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("in:")}
    grads = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("g:")}
    return meta, weights, batch, torch.from_numpy(z["prediction"]), float(z["loss"]), grads
