"""diagnostic: where does the exact dense-Adam mode (l2 = 0) deviate from dense torch.optim.Adam?"""
import argparse, os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import rechorus_oracle as O
from rechorus_b200 import plugin
from rechorus_b200.optim import RowSparseOptimizer
l2 = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0
p = plugin.BPRMF.parse_model_args(argparse.ArgumentParser())
a = p.parse_args(["--emb_size", "64", "--num_neg", "4", "--table_mode", "fused"])
a.device, a.model_path = torch.device("cuda", 0), "/tmp/_b2r_exact.pt"
torch.manual_seed(21)
model = plugin.BPRMF(a, types.SimpleNamespace(n_users=60, n_items=90)).to(a.device)
with torch.no_grad():
    for prm in model.parameters():
        prm.mul_(30.0)
w0 = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
model.optimizer = RowSparseOptimizer(model, "Adam", lr=1e-2, l2=l2, exact_dense=True)
ref = O.ReferenceStyleTrainer("BPRMF", w0, lr=1e-2, l2=l2, optimizer="Adam")
g = torch.Generator().manual_seed(22)
model.train()
hist_i = {}
gmin = {}
track = {}     # (row, col) of the user table -> list of (step, g_ref, w_ref_after, w_gpu_after_flush)
for step in range(25):
    uid = torch.randint(1, 60, (8,), generator=g)
    iid = torch.randint(1, 90, (8, 5), generator=g)
    model.optimizer.zero_grad()
    loss = model.loss(model({"user_id": uid.cuda(), "item_id": iid.cuda(), "batch_size": 8, "phase": "train"}))
    loss.backward()
    model.optimizer.step()
    ref.step({"user_id": uid, "item_id": iid}, shuffle=False)
    if os.environ.get("DIAG_FLUSH"):
        model.optimizer.flush()                     # exact mode must be invariant to extra flushes
        e_i = (model.i_embeddings.weight.detach().cpu() - ref.p["i_embeddings.weight"].detach()).abs().max(dim=1).values
        bad = (e_i > 1e-5).nonzero().reshape(-1).tolist()
        if bad:
            print("step", step + 1, "rows off", [(r, float(e_i[r]), sorted(set(hist_i.get(r, [])))[-4:]) for r in bad[:5]],
                  "ids this step", sorted(set(iid.reshape(-1).tolist()))[:0])
    gu = ref.p["u_embeddings.weight"].grad
    for (r, c) in ((46, 44), (46, 54), (46, 0)):
        wg = float(model.u_embeddings.weight[r, c]) if os.environ.get("DIAG_FLUSH") else float("nan")
        track.setdefault((r, c), []).append((step + 1, float(gu[r, c]), float(ref.p["u_embeddings.weight"][r, c]), wg,
                                             46 in uid.tolist()))
    for r in iid.reshape(-1).tolist():
        hist_i.setdefault(r, []).append(step + 1)
    gi = ref.p["i_embeddings.weight"].grad
    for r in set(iid.reshape(-1).tolist()):
        gmin[r] = min(gmin.get(r, 1e9), float(gi[r].abs().min()))
model.optimizer.flush()
for k, v in model.state_dict().items():
    err = (v.cpu() - ref.p[k].detach()).abs()
    print(k, "max", float(err.max()), "q99.5", float(err.reshape(-1).quantile(0.995)), "entries > 2e-5:", int((err > 2e-5).sum()))
    rows = (err > 2e-5).any(dim=1).nonzero().reshape(-1).tolist()
    for r in rows[:8]:
        cols = (err[r] > 2e-5).nonzero().reshape(-1).tolist()
        print("   row", r, "ncols", len(cols), "cols", cols[:6], "err", [float(err[r, c]) for c in cols[:4]],
              "touched at", sorted(set(hist_i.get(r, []))) if k.startswith("i_") else "-", "min|g| seen", gmin.get(r) if k.startswith("i_") else "-")
        if k.startswith("i_"):
            opt = ref.opt.state[ref.p[k]]
            c = cols[0]
            print("      ref m,v at col", float(opt["exp_avg"][r, c]), float(opt["exp_avg_sq"][r, c]))

for key, rows in track.items():
    print("user entry", key)
    for t in rows:
        print("   step %2d g_ref % .3e w_ref % .8f w_gpu % .8f diff % .2e touched %s" % (t[0], t[1], t[2], t[3], t[3] - t[2], t[4]))
