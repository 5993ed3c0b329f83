"""Trainer shell (shadow_gnn_b200/main.py): Metrics against scikit-learn (what shaDow/metric.py:58-82 calls), the logger's
epoch bookkeeping / best-model rule (shaDow/logging_base.py:108-124,274-310; metric.py:106-128), and -- on the GPU -- a few epochs of
`train()` (shaDow/main.py:172-201) on a learnable synthetic task, eager and through the whole-step CUDA graph."""
import csv
import os

import numpy as np
import pytest
import torch

from shadow_gnn_b200.main import LoggerBase, Metrics, TRAIN, VALID, TEST, one_epoch, train


def test_metrics_match_sklearn():
    from sklearn import metrics as skm
    rng = np.random.default_rng(0)
    y_true = np.eye(7)[rng.integers(0, 5, 400)]                      # classes 5, 6 never occur in y_true
    y_pred = rng.random((400, 7))
    m = Metrics("toy", False, "f1")
    got = m.calc(y_true, y_pred)
    assert np.isclose(got["f1mic"], skm.f1_score(y_true.argmax(1), y_pred.argmax(1), average="micro"))
    assert np.isclose(got["f1mac"], skm.f1_score(y_true.argmax(1), y_pred.argmax(1), average="macro"))
    assert np.isclose(Metrics("toy", False, "accuracy").calc(y_true, y_pred)["accuracy"], skm.accuracy_score(y_true.argmax(1), y_pred.argmax(1)))
    yt = (rng.random((300, 6)) > 0.7).astype(np.float32)
    yp = rng.random((300, 6)).astype(np.float32)
    got = Metrics("toy", True, "f1").calc(yt, yp.copy())
    yb = (yp > 0.5).astype(np.int64)
    assert np.isclose(got["f1mic"], skm.f1_score(yt, yb, average="micro")) and np.isclose(got["f1mac"], skm.f1_score(yt, yb, average="macro"))


def test_is_better_window_rule():
    m = Metrics("toy", False, "accuracy", metric_win_size=2)
    assert m.is_better([1.0, 0.8], float("inf"), [0.5, 0.7], -float("inf")) == (True, 0.9, 0.6)
    assert m.is_better([1.0, 0.8, 0.9], 0.9, [0.5, 0.7, 0.4], 0.6) == (False, 0.9, 0.6)
    f = Metrics("toy", False, "f1")
    assert f.is_better([0.3], 1.0, [0.8], 0.7, [0.2], 0.6) == (True, 0.3, 0.8, 0.2)          # decided by micro F1 only (metric.py:117-128)


class _Toy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(3))
        self.optimizer = None


def test_logger_epoch_summary_best_model_and_csv(tmp_path):
    lg = LoggerBase("toy", False, metric="accuracy", dir_log=str(tmp_path))
    lg.init_log2file("running")
    model = _Toy()
    accs = []
    for ep, (acc_target, wval) in enumerate([(0.5, 1.0), (0.75, 2.0), (0.25, 3.0)]):
        with torch.no_grad():
            model.w.fill_(wval)
        lg.epoch_start_reset(ep, VALID, 8)
        labels = torch.eye(2)[torch.tensor([0, 1, 0, 1, 0, 1, 0, 1])]
        right = int(8 * acc_target)
        preds = labels.clone(); preds[right:] = 1 - preds[right:]
        for b, (lo, hi, loss) in enumerate([(0, 6, 2.0), (6, 8, 4.0)]):
            lg.update_batch(VALID, b, {"batch_size": hi - lo, "loss": torch.tensor(loss), "labels": labels[lo:hi], "preds": preds[lo:hi]})
        lg.update_epoch(ep, VALID)
        row = lg.log_key_step(VALID, time=0.1)
        assert np.isclose(row["loss"], (6 * 2.0 + 2 * 4.0) / 8) and np.isclose(row["accuracy"], acc_target)      # entity-weighted loss (logging_base.py:110)
        accs.append(lg.update_best_model(ep, model))
    assert accs == [True, True, False] and lg.info_epoch[VALID].epoch_best == 1
    with torch.no_grad():
        model.w.fill_(9.0)
    lg.restore_model(model)
    assert torch.equal(model.w.detach(), torch.full((3,), 2.0))
    rows = list(csv.reader(open(os.path.join(str(tmp_path), "running.csv"))))
    assert rows[0] == ["epoch", "mode", "loss", "accuracy", "time"] and len(rows) == 4 and rows[2][1] == "valid"
    with pytest.raises(AssertionError):                                   # out-of-sync batch index (logging_base.py:64)
        lg.epoch_start_reset(3, VALID, 8)
        lg.update_batch(VALID, 1, {"batch_size": 1, "loss": 0.0, "labels": labels[:1], "preds": preds[:1]})


@pytest.mark.gpu
@pytest.mark.parametrize("graphed", [False, True])
def test_train_shell_learns_a_planted_task(tmp_path, graphed):
    """main.train for 4 epochs on a planted-partition graph whose labels are the communities and whose features are noisy community codes:
    validation accuracy must leave chance level far behind, the best epoch is restored for the final pass, the CSV logs hold every epoch"""
    from shadow_gnn_b200 import minibatch as MB
    from shadow_gnn_b200.models import DeepGNN
    from shadow_gnn_b200.train import GraphedTrainer
    rng = np.random.default_rng(3)
    N, C, Fd = 3000, 6, 16
    comm = rng.integers(0, C, N)
    src = np.repeat(np.arange(N), 8)
    same = rng.random(src.size) < 0.8
    # endpoints: inside the community with probability 0.8 (drawn per community), anywhere otherwise
    members = [np.flatnonzero(comm == c) for c in range(C)]
    dst = np.array([members[comm[u]][rng.integers(0, members[comm[u]].size)] if s else rng.integers(0, N) for u, s in zip(src, same)])
    keep = src != dst
    key = np.unique(np.concatenate([src[keep] * N + dst[keep], dst[keep] * N + src[keep]]))
    rows, indices = key // N, (key % N).astype(np.uint32)
    indptr = np.zeros(N + 1, np.int64); np.cumsum(np.bincount(rows, minlength=N), out=indptr[1:])
    indptr = indptr.astype(np.uint32)
    codes = rng.standard_normal((C, Fd)).astype(np.float32)
    feat = torch.from_numpy(codes[comm] * 0.5 + rng.standard_normal((N, Fd)).astype(np.float32))
    label = torch.from_numpy(comm.astype(np.int64))
    perm = rng.permutation(N)
    sets = {0: perm[:1024].astype(np.int64), 1: perm[1024:1536].astype(np.int64), 2: perm[1536:2048].astype(np.int64)}
    cfg = {"batch_size": 32, "configs": [{"method": "ppr", "k": [20], "threshold": [0.0], "epsilon": [1e-4]}]}
    arch = dict(num_layers=2, num_cls_layers=1, heads=1, branch_sharing=False, dim=32, act="relu", layer_norm="norm_feat", feature_augment_ops="sum",
                aggr="sage", residue="none", pooling="center", loss="softmax", ensemble_act="leakyrelu")
    torch.manual_seed(1); np.random.seed(1)
    mb = MB.MinibatchShallowExtractor("toy", None, {m: (indptr, indices) for m in range(3)}, sets, cfg, set(), None, feat, label, Fd, True, 1,
                                      seed_cpp=1, num_subg_per_batch=256)
    model = DeepGNN(Fd, Fd, C, 0, arch, [], 1, dict(dropout=0.1, dropedge=0.0, lr=0.01, ensemble_dropout="none"), "node").cuda()
    lg = LoggerBase("toy", False, metric="accuracy", dir_log=str(tmp_path), log_test_convergence=2)
    tr = GraphedTrainer(model, mb, row_cap=32 * 21, edge_cap=32 * 21 * 21) if graphed else None
    trainers = {TRAIN: tr, VALID: GraphedTrainer(model, mb, row_cap=32 * 21, edge_cap=32 * 21 * 21, mode=VALID),
                TEST: GraphedTrainer(model, mb, row_cap=32 * 21, edge_cap=32 * 21 * 21, mode=TEST)} if graphed else None
    final = train(model, mb, 4, lg, trainer=trainers)
    val = lg.info_epoch[VALID].acc["accuracy"]
    assert len(val) == 5 and max(val[:4]) > 0.6, val                      # chance = 1/6
    assert lg.info_epoch[TRAIN].loss[3] < lg.info_epoch[TRAIN].loss[0]
    assert np.isclose(final[VALID]["accuracy"], val[lg.info_epoch[VALID].epoch_best], atol=0.02)      # best epoch restored (eval is deterministic up to fp order)
    assert final[TEST]["accuracy"] > 0.6
    running = list(csv.reader(open(os.path.join(str(tmp_path), "running.csv"))))
    assert len(running) == 1 + 4 * 2 + 2                                  # header, 4 x (train, valid), test at epochs 0 and 2
    assert len(list(csv.reader(open(os.path.join(str(tmp_path), "final.csv"))))) == 1 + 3
    if graphed:
        assert tr.graph_steps == 4 * 32 and tr.eager_steps == 0
        assert trainers[VALID].graph_steps == 5 * 16 and trainers[VALID].eager_steps == 0
        # the captured forward pass evaluates exactly like the eager DeepGNN.step
        lg2 = LoggerBase("toy", False, metric="accuracy")
        a = one_epoch(0, VALID, model, mb, lg2, status="final", trainer=trainers)
        b = one_epoch(1, VALID, model, mb, lg2, status="final")
        assert a["accuracy"] == b["accuracy"] and abs(a["loss"] - b["loss"]) <= 1e-5 * max(1.0, abs(b["loss"]))


@pytest.mark.gpu
def test_yaml_config_to_training_run(tmp_path):
    """a training config in the reference's YAML schema -> config.load_config -> config.instantiate -> main.train on the files of a dataset in
    shaDow's on-disk format read straight into HBM (loader.load_data_device): the whole f-2 / f-4 chain in one run"""
    from shadow_gnn_b200.config import instantiate, load_config
    from shadow_gnn_b200.loader import load_data_device
    from tests.golden.make_loader_golden import NAME, write_dataset
    write_dataset(str(tmp_path))
    yml = tmp_path / "gcn_2_khop.yml"
    yml.write_text("""
data: {transductive: True, to_undirected: True, norm_feat: True}
architecture: {dim: 32, aggr: gcn, loss: softmax, num_layers: 2, act: relu, feature_augment: hops, residue: none, pooling: center}
hyperparameter: {end: 2, lr: 0.01, dropout: 0.1, dropedge: 0.1, batch_size: 16}
sampler:
  - {method: khop, phase: train, depth: [2], budget: [5]}
""")
    params, pre, cfg_train, cfg_data, arch = load_config(str(yml))
    adjs, feat, label, node_set = load_data_device({"local": str(tmp_path)}, NAME, cfg_data, torch.device("cuda:0"))
    model, mb = instantiate(NAME, adjs, feat, label, node_set, params, arch, cfg_train, config_sampler_preproc=pre, seed_cpp=3, num_subg_per_batch=64)
    assert mb.num_ensemble == 1 and cfg_train["configs"][0]["add_self_edge"] == [True]
    is_sigmoid = label.dim() == 2 and arch["loss"] == "sigmoid"
    lg = LoggerBase(NAME, is_sigmoid, metric="accuracy", dir_log=str(tmp_path / "log"))
    final = train(model, mb, params["end"], lg)
    assert set(final) == {TRAIN, VALID, TEST} and all(np.isfinite(final[m]["loss"]) for m in final)
    assert len(lg.info_epoch[TRAIN].loss) == params["end"] + 1
    assert os.path.isfile(os.path.join(str(tmp_path / "log"), "running.csv"))


@pytest.mark.gpu
def test_cli_entry_runs_a_config(tmp_path, capsys):
    """`python -m shadow_gnn_b200.main --configs ... --dataset ...` (shaDow/main.py:344-449): YAML + on-disk dataset -> captured train / eval epochs"""
    from shadow_gnn_b200.main import main
    from tests.golden.make_loader_golden import NAME, write_dataset
    write_dataset(str(tmp_path))
    yml = tmp_path / "sage_2_ppr.yml"
    yml.write_text("""
data: {transductive: True, to_undirected: True}
architecture: {dim: 32, aggr: sage, loss: softmax, num_layers: 2, act: relu, feature_augment: hops, residue: none, pooling: center}
hyperparameter: {end: 2, lr: 0.01, dropout: 0.1, batch_size: 20}
sampler:
  - {method: ppr, phase: train, k: [15], epsilon: [1e-4]}
""")
    final = main(["--configs", str(yml), "--dataset", NAME, "--dir_data", str(tmp_path), "--dir_log", str(tmp_path / "log"), "--seed", "5"])
    assert all(np.isfinite(final[m]["loss"]) and 0.0 <= final[m]["accuracy"] <= 1.0 for m in (TRAIN, VALID, TEST))
    assert "valid" in capsys.readouterr().out
