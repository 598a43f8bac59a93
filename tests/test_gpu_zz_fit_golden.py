"""GPU end-to-end against the REFERENCE's own runner (tests/golden/fit_*.npz: two epochs of the unmodified
BaseRunner.fit + evaluate on CPU, made by tests/golden/make_fit_golden.py): this repository's runner drives the
kernel-backed models on the same corpus, seeds and flags, in 'dense' table mode with the stock torch.optim the runner
builds -- the exact reference semantics.

What can be asserted: the batches, permutations and optimizer steps are identical (pinned exactly on CPU by
tests/test_fit_golden_cpu.py), so the per-epoch losses agree to fp32 rounding.  The final weights agree tightly for
BPRMF; for the deep models Adam divides every gradient entry by its own running magnitude, so entries whose true
gradient is (near) zero -- the key bias of an attention head is exactly that: softmax is invariant to it -- receive
steps of size ~lr made of rounding noise in ANY two correct fp32 implementations.  Those are compared through what
they cannot change: losses and the model's scores."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_fit_golden_cpu import fit_corpus, run_case  # noqa: E402

pytestmark = pytest.mark.gpu


def _cls(name):
    from rechorus_b200 import plugin
    return getattr(plugin, name)


def _gold_final(gold):
    return {k[3:]: torch.from_numpy(gold[k]) for k in gold.files if k.startswith("w1:")}


def test_bprmf_two_epochs_equal_reference_runner():
    gold, losses, metrics, final = run_case("fit_bprmf", _cls("BPRMF"), torch.device("cuda", 0), ["--table_mode", "dense"])
    assert np.allclose(losses, gold["losses"], rtol=0, atol=2e-5), (losses, gold["losses"])
    ref = _gold_final(gold)
    for k, v in ref.items():
        assert (final[k] - v).abs().max() <= 5e-5, k
    for k in gold.files:
        if k.startswith("m:"):
            assert abs(metrics[k[2:]] - float(gold[k])) <= 1.0 / 48 + 1e-9, (k, metrics[k[2:]], float(gold[k]))


@pytest.mark.parametrize("case", ["fit_neumf", "fit_sasrec"])
def test_deep_models_two_epochs_track_reference_runner(case):
    name = fit_corpus.CASES[case][0]
    gold, losses, metrics, final = run_case(case, _cls(name), torch.device("cuda", 0), ["--table_mode", "dense"])
    assert np.allclose(losses, gold["losses"], rtol=0, atol=1e-4), (losses, gold["losses"])
    keys = [k[3:] for k in gold.files if k.startswith("w1:")]
    dev = torch.cat([(final[k].double() - torch.from_numpy(gold["w1:" + k]).double()).abs().reshape(-1) for k in keys])
    # The yardstick is the reference's OWN rounding sensitivity on this run: the same two epochs of the unmodified
    # reference in float64 ("w1_64:", tests/golden/make_fit_golden.py).  Where the fp32 reference itself sits 2e-5
    # (median) to 0.1 (max) away from the exact result -- NeuMF near its flat start: gradient entries are rounding
    # noise and Adam turns them into lr-sized steps -- no fp32 implementation can be asked to sit closer to it.
    ref = torch.cat([(torch.from_numpy(gold["w1:" + k]).double() - torch.from_numpy(gold["w1_64:" + k])).abs().reshape(-1)
                     for k in keys])
    for name_q, q in (("median", 0.5), ("q90", 0.9), ("q99", 0.99)):
        got, yard = float(dev.quantile(q)), float(ref.quantile(q))
        assert got <= max(1e-6, 4.0 * yard), (case, name_q, got, yard)
    assert float(dev.max()) <= max(1e-4, 4.0 * float(ref.max())), (case, float(dev.max()), float(ref.max()))
    for k in gold.files:
        if k.startswith("m:"):
            assert 0.0 <= metrics[k[2:]] <= 1.0 and abs(metrics[k[2:]] - float(gold[k])) <= 0.15, k


def test_whole_step_fit_equals_autograd_fit_under_sgd():
    """--fused_step 1 (BaseRunner.fit's loop body as one C call per batch, next plan prefetched, ragged last batch)
    against the autograd-node route of --fused_optimizer 1 on the same batches: same row-sparse SGD, so the final
    weights agree to rounding."""
    import argparse
    from rechorus_b200 import plugin
    from rechorus_b200.runner import BaseRunner
    out = {}
    for tag, extra in (("nodes", []), ("steps", ["--fused_step", "1"])):
        p = argparse.ArgumentParser()
        p = BaseRunner.parse_runner_args(p)
        p = plugin.BPRMF.parse_model_args(p)
        a = p.parse_args(["--emb_size", "64", "--num_neg", "7", "--batch_size", "96", "--num_workers", "0", "--lr", "0.05",
                          "--l2", "0", "--optimizer", "SGD", "--table_mode", "fused", "--fused_optimizer", "1", *extra])
        a.device, a.model_path, a.log_file = torch.device("cuda", 0), "/tmp/_b2r_ws.pt", ""
        corpus = fit_corpus.build()
        torch.manual_seed(9)
        model = plugin.BPRMF(a, corpus).to(a.device)
        with torch.no_grad():
            for prm in model.parameters():
                prm.mul_(30.0)                       # gradients well above rounding
        train = plugin.BPRMF.Dataset(model, corpus, "train")
        runner = BaseRunner(a)
        np.random.seed(11)
        torch.manual_seed(11)
        losses = [runner.fit(train, epoch=e + 1) for e in range(2)]
        out[tag] = (losses, {k: v.detach().cpu() for k, v in model.state_dict().items()})
    # the whole-step kernel evaluates the loss with the fast exp/division forms (<= 2 ulp each): ~1e-6 relative
    assert np.allclose(out["nodes"][0], out["steps"][0], rtol=2e-5, atol=2e-5), (out["nodes"][0], out["steps"][0])
    for k, v in out["nodes"][1].items():
        assert (v - out["steps"][1][k]).abs().max() <= 2e-6, k


def test_eval_ranks_test_all_equal_reference_predict_fixture():
    """the plugin's device route for --test_all 1 (b2r_rank_all_items + clicked-item masking, scores never
    materialised) gives the integer ranks of the REFERENCE's own BaseRunner.predict + evaluate_method
    (tests/golden/eval_test_all.npz, made by tests/golden/make_eval_golden.py)"""
    import argparse
    import os
    import sys
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, golden)
    import fit_corpus
    from rechorus_b200 import ops, plugin
    from rechorus_b200.runner import BaseRunner
    gold = np.load(os.path.join(golden, "eval_test_all.npz"))
    p = argparse.ArgumentParser()
    p = BaseRunner.parse_runner_args(p)
    p = plugin.BPRMF.parse_model_args(p)
    a = p.parse_args(["--emb_size", "64", "--test_all", "1", "--device_metrics", "1"] + fit_corpus.COMMON)
    a.device, a.model_path, a.log_file = torch.device("cuda", 0), "/tmp/_b2r_evalg.pt", ""
    corpus = fit_corpus.build()
    model = plugin.BPRMF(a, corpus)
    model.load_state_dict({k[2:]: torch.from_numpy(gold[k]) for k in gold.files if k.startswith("w:")})
    model = model.to(a.device)
    dev = plugin.BPRMF.Dataset(model, corpus, "dev")
    dev.prepare()
    runner = BaseRunner(a)
    metrics = runner.evaluate(dev, [5, 10, 20], ["HR", "NDCG"])          # device route
    for k in gold.files:
        if k.startswith("m:"):
            assert abs(metrics[k[2:]] - float(gold[k])) <= 1e-12, (k, metrics[k[2:]], float(gold[k]))
    # and the ranks themselves, batch by batch
    batch = dev.collate_batch([dev[i] for i in range(len(dev))])
    uids = dev.data["user_id"]
    rows = torch.tensor([i for i, u in enumerate(uids) for _ in (corpus.train_clicked_set[u] | corpus.residual_clicked_set[u])])
    cols = torch.tensor([j for u in uids for j in (corpus.train_clicked_set[u] | corpus.residual_clicked_set[u])])
    feed = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}
    model.eval()
    ranks = model.eval_ranks(feed, rows.cuda(), cols.cuda()).cpu().numpy()
    assert np.array_equal(ranks, gold["gt_rank"])
    ops.check_ids()
