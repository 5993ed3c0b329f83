"""YAML front end (shadow_gnn_b200/config.py) against the rules of shaDow/utils.py:53-131: defaults, checks, sampler phases, the self-edge rule
for GCN / GAT; every training config of the reference tree parses when that tree is present (it is not on the GPU box)."""
import glob
import os

import pytest
import yaml

from shadow_gnn_b200.config import load_config, parse_config

YML = """
data:
  transductive: True
architecture:
  dim: 64
  aggr: GAT
  heads: 4
  loss: softmax
  num_layers: 3
  act: relu
  feature_augment: hops-pprs
  residue: max
  pooling: sort-5
hyperparameter:
  end: 3
  lr: 2e-3
  dropout: 0.3
  batch_size: 32
  percent_per_epoch: {train: 0.5}
sampler:
  - method: ppr
    phase: train
    k: [20, 40]
    epsilon: [1e-4, 1e-4]
  - method: khop
    phase: train
    depth: [2, 1]
    budget: [5, 8]
"""


def test_parse_defaults_checks_and_self_edge_rule(tmp_path):
    p = tmp_path / "c.yml"
    p.write_text(YML)
    params, pre, train, data, arch = load_config(str(p))
    assert data == {"to_undirected": False, "transductive": True, "norm_feat": True, "valedges_as_input": False}
    assert arch["aggr"] == "gat" and arch["feature_augment"] == {"hops", "pprs"} and arch["num_cls_layers"] == 1 and arch["layer_norm"] == "norm_feat"
    assert arch["ensemble_act"] == "leakyrelu" and arch["branch_sharing"] is False and arch["pooling"] == "sort-5"
    assert params["lr"] == 0.002 and params["dropedge"] == 0.0 and params["ensemble_dropout"] == "none"
    assert params["percent_per_epoch"] == {"train": 0.5, "valid": 1.0, "test": 1.0}
    assert pre == {"batch_size": 32, "configs": []}
    assert [c["method"] for c in train["configs"]] == ["ppr", "khop"] and all("phase" not in c for c in train["configs"])
    assert train["configs"][0]["add_self_edge"] == [True, True] and train["configs"][1]["add_self_edge"] == [True, True]      # utils.py:120-125
    bad = yaml.safe_load(YML)
    bad["architecture"]["residue"] = "mean"
    with pytest.raises(AssertionError):
        parse_config(bad)
    bad = yaml.safe_load(YML)
    bad["sampler"][0]["phase"] = "inference"
    with pytest.raises(NotImplementedError):
        parse_config(bad)
    sage = yaml.safe_load(YML)
    sage["architecture"]["aggr"] = "sage"
    assert "add_self_edge" not in parse_config(sage)[2]["configs"][0]


REF = "/root/reference/config_train"


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is not present on this box")
def test_every_reference_training_config_parses():
    files = sorted(glob.glob(f"{REF}/**/*.yml", recursive=True))
    assert len(files) >= 60
    n_hops = 0
    for f in files:
        params, pre, train, data, arch = load_config(f)
        assert arch["dim"] > 0 and arch["num_layers"] > 0 and train["batch_size"] > 0 and (train["configs"] or pre["configs"])
        n_hops += "hops" in arch["feature_augment"]
        for c in train["configs"]:
            assert c["method"] in ("ppr", "khop", "ppr_st", "nodeIID", "full"), (f, c["method"])
    assert n_hops >= 40                       # `feature_augment: hops` is the common case (it stays on the fast path)
    params, pre, train, data, arch = load_config(f"{REF}/products/vanilla/sage_5_ppr.yml")
    assert arch["dim"] == 256 and arch["num_layers"] == 5 and train["configs"][0]["k"] == [150] and params["dropedge"] == 0.05
