"""YAML front end of the trainer shell: the reference's training configs (`config_train/<data>/<...>.yml`) parsed with the
defaults and checks of `shaDow/utils.py:20-137` (`parse_n_prepare`: data / architecture / hyperparameter / sampler sections), and `instantiate`
(`shaDow/main.py:33-133`) for the device-resident minibatch + model of this repository.  Out of scope like in DESIGN.md 7: `phase: preprocess`
samplers (SGC / SIGN smoothing, label propagation), post-processing configs, the CLI.
"""
from copy import deepcopy

import yaml

from .minibatch import TRAIN, VALID, TEST

DATA_DEFAULTS = {"to_undirected": False, "transductive": False, "norm_feat": True, "valedges_as_input": False}
ARCH_DEFAULTS = {"dim": -1, "aggr": "sage", "residue": "none", "pooling": "center", "loss": "softmax", "num_layers": -1, "num_cls_layers": 1, "act": "I",
                 "layer_norm": "norm_feat", "heads": -1, "feature_augment": "hops", "feature_augment_ops": "sum", "feature_smoothen": "none",
                 "label_smoothen": "none", "ensemble_act": "leakyrelu", "branch_sharing": False, "use_label": "none"}
TRAIN_DEFAULTS = {"lr": 0.01, "dropedge": 0.0, "ensemble_dropout": "none", "term_window_size": 1, "term_window_aggr": "center",
                  "percent_per_epoch": {"train": 1.0, "valid": 1.0, "test": 1.0}}


def parse_config(config_train):
    """dict (a loaded YAML) -> (params_train, config_sampler_preproc, config_sampler_train, config_data, arch_gnn), utils.py:53-131"""
    config_train = deepcopy(config_train)
    config_data = dict(DATA_DEFAULTS)
    config_data.update(config_train.get("data") or {})
    arch_gnn = dict(ARCH_DEFAULTS)
    arch_gnn.update(config_train["architecture"])
    for k, v in arch_gnn.items():
        if type(v) == str:
            arch_gnn[k] = v.lower()
    assert arch_gnn["aggr"] in ["sage", "gat", "gatscat", "gcn", "mlp", "gin", "sgc", "sign"]
    assert arch_gnn["use_label"] in ["all", "none", "no_valid"]
    assert arch_gnn["pooling"].split("-")[0] in ["mean", "max", "sum", "center", "sort"]
    assert arch_gnn["residue"] in ["sum", "concat", "max", "none"]
    assert arch_gnn["feature_augment"] in ["hops", "pprs", "none", "hops-pprs", "drnls"]
    assert arch_gnn["feature_augment_ops"] in ["concat", "sum"]
    assert arch_gnn["layer_norm"] in ["norm_feat", "pairnorm"]
    fa = arch_gnn["feature_augment"]
    arch_gnn["feature_augment"] = set(fa.split("-")) if fa and fa != "none" else set()
    params_train = deepcopy(TRAIN_DEFAULTS)
    params_train.update(config_train["hyperparameter"])
    params_train["lr"] = float(params_train["lr"])
    for m in ("train", "valid", "test"):
        params_train["percent_per_epoch"].setdefault(m, 1.0)          # (the reference indexes with a stale loop variable here, utils.py:104)
        assert 0 <= params_train["percent_per_epoch"][m] <= 1.0
    sampler_preproc, sampler_train = [], []
    for s in config_train["sampler"]:
        phase = s.pop("phase")
        if phase == "preprocess":
            sampler_preproc.append(s)
        elif phase == "train":
            sampler_train.append(s)
        else:
            raise NotImplementedError(phase)
    batch_size = config_train["hyperparameter"]["batch_size"]
    cfg_pre = {"batch_size": batch_size, "configs": sampler_preproc}
    cfg_train = {"batch_size": batch_size, "configs": sampler_train}
    if arch_gnn["aggr"] in ["gcn", "gat", "gatscat"]:                  # self edges: GAT's softmax needs a non-empty row (utils.py:120-125)
        for sc in cfg_train["configs"]:
            num_ens = [len(v) for k, v in sc.items() if k != "method"]
            assert max(num_ens) == min(num_ens)
            sc["add_self_edge"] = [True] * num_ens[0]
    return params_train, cfg_pre, cfg_train, config_data, arch_gnn


def load_config(path):
    with open(path) as f:
        return parse_config(yaml.load(f, Loader=yaml.FullLoader))


def instantiate(name_data, adjs, feat_full, label_full, entity_set, params_train, arch_gnn, config_sampler_train, *, config_sampler_preproc=None,
                seed_cpp=-1, device=None, num_subg_per_batch=None, rng="glibc"):
    """main.py:33-133 for node / link data that already sits in the form `loader.load_data_device` returns: adjs[mode] = (indptr, indices)
    device tensors (or host arrays), features, labels, entity sets.  Returns (model, minibatch)."""
    import torch
    from .minibatch import MinibatchShallowExtractor
    from .models import DeepGNN
    if config_sampler_preproc is not None and len(config_sampler_preproc["configs"]) > 0:
        raise NotImplementedError("phase: preprocess samplers (SGC / SIGN / label smoothing) are out of scope")
    if arch_gnn["aggr"] in ("sgc", "sign") or arch_gnn["use_label"] != "none":
        raise NotImplementedError("SGC / SIGN architectures and label reuse need the preprocessing pass (out of scope)")
    dim_feat_raw = feat_full.shape[1]
    if label_full is not None:
        if label_full.dim() == 1:
            dim_label_raw = int(label_full[label_full == label_full].max().item()) + 1       # main.py:70-71 (NaN labels skipped)
        else:
            dim_label_raw = label_full.shape[1]
    else:
        dim_label_raw = 1
    ip_tr, ip_full = adjs[TRAIN][0], adjs[VALID][0]
    is_transductive = ip_tr is ip_full or (ip_tr.shape == ip_full.shape and adjs[TRAIN][1].shape == adjs[VALID][1].shape)
    bs = config_sampler_train["batch_size"]
    kw = {} if num_subg_per_batch is None else {"num_subg_per_batch": num_subg_per_batch}
    minibatch = MinibatchShallowExtractor(name_data, None, adjs, entity_set, config_sampler_train, arch_gnn["feature_augment"], params_train["percent_per_epoch"],
                                          feat_full, label_full, dim_feat_raw, is_transductive, 1, seed_cpp=seed_cpp, device=device, rng=rng, **kw)
    aug_feat = [(k, minibatch.get_aug_dim(k)) for k in sorted(arch_gnn["feature_augment"])]
    model = DeepGNN(dim_feat_raw, dim_feat_raw, dim_label_raw, 0, arch_gnn, aug_feat, minibatch.num_ensemble, params_train, minibatch.prediction_task)
    dev = minibatch.dev_torch if hasattr(minibatch, "dev_torch") else (device or torch.device("cuda"))
    return model.to(dev), minibatch
