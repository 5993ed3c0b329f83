"""CPU tests: our `ParallelSampler` module exposes the surface of the reference's pybind module (PS.cpp:707-746).
Compared against the compiled, unmodified reference (oracle/_ref) when it is present, else against the list in SURVEY.md 8b."""
import inspect

import pytest

REF_SAMPLER = ["num_nodes", "num_edges", "num_nodes_target", "shuffle_targets", "get_idx_root", "is_seq_root_traversal",
               "preproc_ppr_approximate", "parallel_sampler_ensemble", "drop_full_graph_info"]
REF_VEC = ["get_num_valid_subg"] + [f"get_subgraph_{n}" for n in ("indptr", "indices", "data", "node", "edge_index", "target", "hop", "ppr", "drnl")]


def test_surface_matches_reference_module():
    import shadow_gnn_b200.ParallelSampler as ours
    from oracle import oracle as O
    ref = O.load_ref()
    names_s, names_v = REF_SAMPLER, REF_VEC
    if ref is not None:
        names_s = [n for n in dir(ref.ParallelSampler) if not n.startswith("_")]
        names_v = [n for n in dir(ref.SubgraphStructVec) if not n.startswith("_")]
        assert sorted(names_s) == sorted(REF_SAMPLER) and sorted(names_v) == sorted(REF_VEC)
    for n in names_s:
        assert callable(getattr(ours.ParallelSampler, n)), n
    for n in names_v:
        assert callable(getattr(ours.SubgraphStructVec, n)), n
    # 13 positional constructor arguments, in the reference's order (PS.cpp:710-724)
    params = [p.name for p in inspect.signature(ours.ParallelSampler.__init__).parameters.values() if p.kind == p.POSITIONAL_OR_KEYWORD][1:]
    assert params == ["indptr", "indices", "data", "num_sampler_per_batch", "max_num_threads", "fix_target", "sequential_traversal",
                      "edge_reweighted", "num_subgraphs_ensemble", "path_indptr", "path_indices", "path_data", "seed"]


def test_config_parsing_mirrors_stoi_stod():
    from shadow_gnn_b200.ParallelSampler import parse_cfg
    c = parse_cfg({"method": "khop", "depth": " 2", "budget": "-1", "num_roots": "1", "add_self_edge": "True"}, {"hops", "unknown"})
    assert (c.method, c.depth, c.budget, c.add_self_edge, c.aug) == (0, 2, -1, 1, 1)
    c = parse_cfg({"method": "ppr", "k": "150abc", "threshold": "1e-2", "num_roots": "2", "include_target_conn": "1"})       # stoi stops at garbage
    assert (c.k, c.num_roots, c.include_target_conn) == (150, 2, 1) and abs(c.threshold - 0.01) < 1e-9
    with pytest.raises(IndexError):
        parse_cfg({"method": "ppr", "k": "3", "num_roots": "1"})
    with pytest.raises(ValueError):
        parse_cfg({"method": "ppr", "k": "x", "threshold": "0", "num_roots": "1"})


def test_layer_and_model_surface():
    """class names, constructor signatures and parameter names of the reference (shaDow/layers.py, models.py:17-26)"""
    import torch
    from shadow_gnn_b200 import layers as L
    from shadow_gnn_b200.models import DeepGNN
    assert set(DeepGNN.NAME2CLS) == {"mlp", "gcn", "gin", "sage", "gat", "gatscat", "sgc", "sign"}
    keys = lambda m: sorted(m.state_dict().keys())
    assert keys(L.GraphSAGE(4, 8)) == ["f_lin_neigh.bias", "f_lin_neigh.weight", "f_lin_self.bias", "f_lin_self.weight", "offset", "scale"]
    assert keys(L.GCN(4, 8)) == ["f_lin.bias", "f_lin.weight", "offset", "scale"]
    assert keys(L.GIN(4, 8)) == ["eps", "mlp.0.bias", "mlp.0.weight", "mlp.2.bias", "mlp.2.weight", "offset", "scale"]
    assert keys(L.GAT(4, 8, mulhead=2)) == ["attention", "f_lin.0.bias", "f_lin.0.weight", "f_lin.1.bias", "f_lin.1.weight", "offset", "scale"]
    assert tuple(L.GAT(4, 8, mulhead=2).scale.shape) == (2, 2, 4) and tuple(L.GraphSAGE(4, 8).scale.shape) == (2, 8)
    arch = dict(num_layers=2, num_cls_layers=1, heads=1, branch_sharing=False, dim=8, act="relu", layer_norm="norm_feat", feature_augment_ops="sum",
                aggr="sage", residue="cat", pooling="max", loss="softmax", ensemble_act="leakyrelu")
    m = DeepGNN(4, 4, 3, 0, arch, [("hops", 7)], 1, dict(dropout=0.0, dropedge=0.0, lr=0.01, ensemble_dropout="none"), "node")
    k = keys(m)
    assert "conv_layers.0.1.f_lin_neigh.weight" in k and "res_pool_layers.0.nn.1.weight" in k and "classifier.0.f_lin.weight" in k and "aug_layers.0.0.weight" in k
