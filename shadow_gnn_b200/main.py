"""Trainer shell of the reference: `one_epoch` / `train` with the call contract of shaDow/main.py:136-201, a compact
`LoggerBase` (shaDow/logging_base.py: per-batch bookkeeping, entity-weighted epoch loss, best-model tracking on the validation metric, the
running / final CSV logs) and `Metrics` (shaDow/metric.py: accuracy, micro / macro F1, windowed `is_better`).

What differs from the reference is where things live: a batch's loss / labels / predictions stay on the GPU until the epoch is summarised
(the reference pulls every batch to numpy, logging_base.py:54-72 -- a device synchronisation per 32-target step), the best weights are kept
as a device copy of the flat parameter buffer (plus `torch.save` when a directory is given), and TRAIN epochs can run through the
whole-step CUDA graph (`train.GraphedTrainer`).  YAML / CLI parsing, post-processing (C&S, ensemble training) and the OGB evaluators stay
out of scope.
"""
import os
import time

import numpy as np
import torch

from .minibatch import TRAIN, VALID, TEST

MODE2STR = {TRAIN: "train", VALID: "valid", TEST: "test"}
STR2MODE = {v: k for k, v in MODE2STR.items()}
METRICS = {"accuracy": ["accuracy"], "f1": ["f1mic", "f1mac"]}         # logging_base.py:22-29 (node tasks)


def _f1(y_true, y_pred, num_classes):
    """micro / macro F1 of single-label predictions (what sklearn.metrics.f1_score returns for 1-D class vectors: macro averages over the
    classes present in y_true or y_pred, an absent class counts as 0/0 -> skipped by `labels=None`)"""
    cm = np.zeros((num_classes, num_classes), np.int64)
    np.add.at(cm, (y_true, y_pred), 1)
    tp = np.diag(cm).astype(np.float64)
    fp, fn = cm.sum(0) - tp, cm.sum(1) - tp
    present = (cm.sum(0) + cm.sum(1)) > 0
    denom = 2 * tp + fp + fn
    f1 = np.where(denom > 0, 2 * tp / np.maximum(denom, 1), 0.0)
    mic = tp.sum() / max(cm.sum(), 1)
    return float(mic), float(f1[present].mean()) if present.any() else 0.0


def _f1_multilabel(y_true, y_pred):
    """micro / macro F1 of multi-hot predictions (sklearn's label-indicator case)"""
    tp = (y_true * y_pred).sum(0).astype(np.float64)
    fp, fn = y_pred.sum(0) - tp, y_true.sum(0) - tp
    denom = 2 * tp + fp + fn
    f1 = np.where(denom > 0, 2 * tp / np.maximum(denom, 1), 0.0)
    mic_d = 2 * tp.sum() + fp.sum() + fn.sum()
    return float(2 * tp.sum() / mic_d) if mic_d > 0 else 0.0, float(f1.mean())


class Metrics:
    """metric.py:23-147 without the OGB evaluators (`accuracy_ogb` is numerically `accuracy`, metric.py:84-93)"""

    def __init__(self, name_data, is_sigmoid, metric, metric_win_size=1):
        self.window_size, self.name_data, self.is_sigmoid = metric_win_size, name_data, is_sigmoid
        if metric == "accuracy_ogb":
            metric = "accuracy"
        if metric not in METRICS:
            raise NotImplementedError(f"metric {metric}: link-prediction hits@K need the OGB evaluator (out of scope)")
        self.name = metric
        self.metric_term = ("f1mic", "max") if metric == "f1" else ("accuracy", "max")

    def calc(self, y_true, y_pred):
        y_true, y_pred = np.asarray(y_true), np.asarray(y_pred)
        if self.name == "f1" and self.is_sigmoid:                       # metric.py:67-69
            mic, mac = _f1_multilabel((y_true > 0.5).astype(np.int64), (y_pred > 0.5).astype(np.int64))
            return {"f1mic": mic, "f1mac": mac}
        c = y_true.shape[1]
        mic, mac = _f1(y_true.argmax(1), y_pred.argmax(1), c)
        return {"f1mic": mic, "f1mac": mac} if self.name == "f1" else {"accuracy": mic}      # single label: accuracy == micro F1 (metric.py:76-82)

    def is_better(self, loss_all, loss_min_hist, *acc):
        """metric.py:106-128: window-averaged first metric must beat its history; returns (better, loss, metrics...)"""
        alls, hists = acc[0::2], acc[1::2]
        w = self.window_size
        avgs = [sum(a[-w:]) / len(a[-w:]) for a in alls]
        loss_avg = sum(loss_all[-w:]) / len(loss_all[-w:])
        if avgs[0] > hists[0]:
            return (True, loss_avg, *avgs)
        return (False, loss_min_hist, *hists)


class _InfoEpoch:
    def __init__(self, names):
        self.names = names
        self.loss = []
        self.acc = {n: [] for n in names}
        self.idx_epoch = self.epoch_best = -1
        self.loss_min_hist = float("inf")
        self.max_hist = {n: -float("inf") for n in names}


class LoggerBase:
    """logging_base.py:163-533, the part `one_epoch` / `train` call.  CSV rows: `epoch,mode,loss,<metrics...>,time` in
    `<dir_log>/{running,final}.csv` (the reference's per-epoch rows, logging_base.py:375-423)."""

    def __init__(self, name_data, is_sigmoid, metric="accuracy", metric_win_size=1, dir_log=None, log_test_convergence=-1, no_log=False,
                 printf=None):
        self.metrics = Metrics(name_data, is_sigmoid, metric, metric_win_size)
        self.names = METRICS[self.metrics.name]
        self.dir_log = None if no_log else dir_log
        self.log_test_convergence = log_test_convergence
        self._printf = printf
        self.info_epoch = {m: _InfoEpoch(self.names) for m in (TRAIN, VALID, TEST)}
        self._batches = {m: [] for m in (TRAIN, VALID, TEST)}
        self._total_entity = {m: -1 for m in (TRAIN, VALID, TEST)}
        self._idx_batch = -1
        self._best_flat = None
        self.path_saver = None
        if self.dir_log:
            os.makedirs(self.dir_log, exist_ok=True)
            self.path_saver = os.path.join(self.dir_log, "saved_model.pkl")

    def printf(self, msg, style=None):
        if self._printf is not None:
            self._printf(msg, style=style)

    # ---- epoch / batch bookkeeping (logging_base.py:364-373,473-483) ----
    def epoch_start_reset(self, ep, mode, total_entity):
        self._batches[mode] = []
        self._total_entity[mode] = total_entity
        self._idx_batch = -1
        self.info_epoch[mode].idx_epoch = ep

    def update_batch(self, mode, idx_batch, info_batch):
        self._idx_batch += 1
        assert self._idx_batch == idx_batch, "Out of sync between minibatch and logger!!"        # logging_base.py:64
        loss = info_batch["loss"]
        self._batches[mode].append((int(info_batch["batch_size"]), loss.detach() if torch.is_tensor(loss) else loss,
                                    info_batch["labels"].detach(), info_batch["preds"].detach()))

    def update_epoch(self, ep, mode):
        ie, b = self.info_epoch[mode], self._batches[mode]
        assert ep == ie.idx_epoch, "Out of sync between minibatch and logger!!"
        sizes = np.array([x[0] for x in b], np.float64)
        assert int(sizes.sum()) == self._total_entity[mode], "an epoch must cover every entity once"
        losses = torch.stack([torch.as_tensor(x[1], dtype=torch.float32, device=b[0][3].device) for x in b]).double().cpu().numpy()      # ONE sync per epoch
        ie.loss.append(float((losses * sizes).sum() / sizes.sum()))                          # logging_base.py:110-111
        y_true = torch.cat([x[2] for x in b]).cpu().numpy()
        y_pred = torch.cat([x[3] for x in b]).float().cpu().numpy()
        for k, v in self.metrics.calc(y_true, y_pred).items():
            ie.acc[k].append(v)
        self._batches[mode] = []

    # ---- best model (logging_base.py:274-338) ----
    def update_best_model(self, ep, model, optimizer=None):
        ie = self.info_epoch[VALID]
        args = [ie.loss, ie.loss_min_hist]
        for n in self.names:
            args += [ie.acc[n], ie.max_hist[n]]
        ret = self.metrics.is_better(*args)
        if ret[0]:
            ie.epoch_best = ep
            ie.loss_min_hist = ret[1]
            for n, v in zip(self.names, ret[2:]):
                ie.max_hist[n] = v
            self._best_flat = {k: v.detach().clone() for k, v in model.state_dict().items()}
            if self.path_saver:
                torch.save(self._best_flat, self.path_saver)
            self.printf(f"  Saving model {ep:4d} ...", style="yellow")
        return ret[0]

    def restore_model(self, model, optimizer=None, force_reload=False):
        if self._best_flat is None and self.path_saver and os.path.isfile(self.path_saver):
            self._best_flat = torch.load(self.path_saver)
        if self._best_flat is not None:
            model.load_state_dict(self._best_flat)
            self.printf(f"  Restoring model from epoch {self.info_epoch[VALID].epoch_best} ...", style="yellow")

    # ---- CSV (logging_base.py:375-467) ----
    def init_log2file(self, status="running", meta_info=None):
        if self.dir_log:
            with open(os.path.join(self.dir_log, f"{status}.csv"), "w") as f:
                f.write(",".join(["epoch", "mode", "loss"] + self.names + ["time"]) + "\n")

    def log_key_step(self, mode, time=-1, status="running"):
        ie = self.info_epoch[mode]
        row = {"epoch": ie.idx_epoch, "mode": MODE2STR[mode], "loss": ie.loss[-1], **{n: ie.acc[n][-1] for n in self.names}, "time": time}
        if self.dir_log:
            with open(os.path.join(self.dir_log, f"{status}.csv"), "a") as f:
                f.write(",".join(str(row[k]) for k in ["epoch", "mode", "loss"] + self.names + ["time"]) + "\n")
        self.printf("  ".join(f"{k} {v:.4f}" if isinstance(v, float) else f"{k} {v}" for k, v in row.items()))
        return row


def one_epoch(ep, mode, model, minibatch, logger, status="running", pred_mat=None, emb_ens=None, trainer=None):
    """shaDow/main.py:136-169.  `trainer` (a `train.GraphedTrainer`, or a dict mode -> GraphedTrainer) replaces `one_batch` + `model.step`: running
    TRAIN epochs through the captured training step, VALID / TEST epochs through a captured forward pass; its
    batches are logged through the same `update_batch` (loss kept on the device, labels / predictions of the captured step are static
    buffers, so they are cloned)."""
    assert status in ["running", "final"] and mode in [TRAIN, VALID, TEST]
    assert pred_mat is None and emb_ens is None, "post-processing (C&S / ensemble training) is out of scope"
    minibatch.epoch_start_reset(ep, mode)
    minibatch.shuffle_entity(mode)
    logger.epoch_start_reset(ep, mode, minibatch.entity_epoch[mode].shape[0])
    t1 = time.time()
    if isinstance(trainer, dict):                          # {mode: GraphedTrainer}: evaluation epochs through a captured forward pass too
        trainer = trainer.get(mode)
    graphed = trainer is not None and trainer.mode == mode and (mode != TRAIN or status == "running")
    while not minibatch.is_end_epoch(mode):
        if graphed:
            output_batch = trainer.step_logged()
        else:
            input_batch = minibatch.one_batch(mode=mode, ret_raw_idx=False)
            output_batch = model.step(mode, status, input_batch)
        logger.update_batch(mode, minibatch.batch_num, output_batch)
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    t2 = time.time()
    minibatch.epoch_end_reset(mode)
    logger.update_epoch(ep, mode)
    return logger.log_key_step(mode, status=status, time=t2 - t1)


def train(model, minibatch, max_epoch, logger, nocache=None, trainer=None):
    """shaDow/main.py:172-201"""
    logger.init_log2file(status="running")
    logger.init_log2file(status="final")
    if type(nocache) == str and len(nocache) > 0:
        modes = [TRAIN, VALID, TEST] if nocache == "all" else [STR2MODE[nocache]]
        for mode in modes:
            minibatch.disable_cache(mode)
    e = -1
    for e in range(max_epoch):
        one_epoch(e, TRAIN, model, minibatch, logger, trainer=trainer)
        one_epoch(e, VALID, model, minibatch, logger, trainer=trainer)
        if logger.log_test_convergence > 0 and e % logger.log_test_convergence == 0:
            one_epoch(int(e / logger.log_test_convergence), TEST, model, minibatch, logger, trainer=trainer)
        logger.update_best_model(e, model, model.optimizer)
    logger.printf("======================\nOptimization Finished!\n======================\n", style="red")
    logger.restore_model(model, optimizer=None)
    ep_final_test = 0 if logger.log_test_convergence <= 0 else int(e / logger.log_test_convergence) + 1
    ep_final = {TRAIN: e + 1, VALID: e + 1, TEST: ep_final_test}
    return {md: one_epoch(ep_final[md], md, model, minibatch, logger, status="final", trainer=trainer) for md in [TRAIN, VALID, TEST]}


def main(argv=None):
    """`python -m shadow_gnn_b200.main --configs <yml> --dataset <name> --dir_data <root>`: the training entry of shaDow/main.py:344-449 for a
    dataset in shaDow's on-disk format (`<root>/<name>/{adj_full_raw.npz, feat_full.npy, label_full.npy, split.npy, ...}`)."""
    import argparse
    ap = argparse.ArgumentParser(description="shaDow-GNN training on the B200 path")
    ap.add_argument("--configs", required=True, help="training YAML (config_train/... of the reference)")
    ap.add_argument("--dataset", required=True)
    ap.add_argument("--dir_data", default="./data", help="root that holds <dataset>/ in shaDow's on-disk format")
    ap.add_argument("--dir_log", default=None)
    ap.add_argument("--gpu", type=int, default=0)
    ap.add_argument("--epochs", type=int, default=None, help="override hyperparameter.end")
    ap.add_argument("--seed", type=int, default=-1, help="sampler seed (-1: time-seeded like the reference)")
    ap.add_argument("--eager", action="store_true", help="DeepGNN.step per batch instead of the whole-step CUDA graph")
    ap.add_argument("--nocache", default=None)
    ap.add_argument("--log_test_convergence", type=int, default=-1)
    args = ap.parse_args(argv)
    from .config import instantiate, load_config
    from .loader import load_data_device
    from .train import GraphedTrainer
    torch.cuda.set_device(args.gpu)
    dev = torch.device("cuda", args.gpu)
    params, pre, cfg_train, cfg_data, arch = load_config(args.configs)
    adjs, feat, label, node_set = load_data_device({"local": args.dir_data}, args.dataset, cfg_data, dev, printf=lambda t, style=None: print(t))
    model, mb = instantiate(args.dataset, adjs, feat, label, node_set, params, arch, cfg_train, config_sampler_preproc=pre, seed_cpp=args.seed, device=dev,
                            num_subg_per_batch=cfg_train["batch_size"] * 64, rng="philox" if args.seed < 0 else "glibc")
    metric = "f1" if arch["loss"] == "sigmoid" else "accuracy"
    logger = LoggerBase(args.dataset, arch["loss"] == "sigmoid", metric=metric, metric_win_size=params["term_window_size"], dir_log=args.dir_log,
                        log_test_convergence=args.log_test_convergence, printf=lambda t, style=None: print(t))
    trainers = None
    pools = {rp.type_pool for rp in model.res_pool_layers}
    if not args.eager and mb.prediction_task == "node" and "sort" not in pools:
        def scope(c):                                       # upper bound of a subgraph's node count (a batch beyond the capacity takes the eager step)
            if c["method"].startswith("ppr"):
                return max(c.get("k", [0])) + 1
            if c["method"] == "khop" and min(c.get("budget", [-1])) > 0:
                return max(sum(b ** d for d in range(dep + 1)) for b, dep in zip(c["budget"], c["depth"]))
            return 2048
        k = min(4096, max(scope(c) for c in cfg_train["configs"]))
        B = cfg_train["batch_size"]
        trainers = {m: GraphedTrainer(model, mb, row_cap=B * k, edge_cap=B * k * 64, mode=m) for m in (TRAIN, VALID, TEST)}
    final = train(model, mb, params["end"] if args.epochs is None else args.epochs, logger, nocache=args.nocache, trainer=trainers)
    for m in (TRAIN, VALID, TEST):
        print(MODE2STR[m], {k: v for k, v in final[m].items() if k not in ("mode",)})
    return final


if __name__ == "__main__":
    main()
