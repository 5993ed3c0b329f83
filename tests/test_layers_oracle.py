"""CPU tests: the torch fp32 restatement of the layer math (oracle/layers_ref.py) against golden outputs of the
UNMODIFIED reference layers (tests/golden/layers_golden.npz, made by tests/golden/make_layer_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import layers_ref as R

Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "layers_golden.npz"))


def params(name, grad=True):
    pre = f"{name}_p_"
    return {k[len(pre):]: torch.tensor(Z[k], requires_grad=grad) for k in Z.files if k.startswith(pre)}


def adjacency(self_edge):
    ip, ix = Z[f"adj{int(self_edge)}_indptr"], Z[f"adj{int(self_edge)}_indices"]
    return R.dense_counts(ip, ix, ip.size - 1), torch.as_tensor(Z[f"adj{int(self_edge)}_target"]), torch.as_tensor(Z[f"adj{int(self_edge)}_size_subg"])


def check(name, fn, tol=2e-5):
    p = params(name)
    x = torch.tensor(Z[f"{name}_x"], requires_grad=True)
    out = fn(p, x)
    want = torch.tensor(Z[f"{name}_out"])
    assert out.shape == want.shape
    assert torch.allclose(out, want, rtol=tol, atol=tol), (name, (out - want).abs().max())
    (out * torch.tensor(Z[f"{name}_w"])).sum().backward()
    assert torch.allclose(x.grad, torch.tensor(Z[f"{name}_dx"]), rtol=1e-4, atol=1e-4), name
    for k, v in p.items():
        g = v.grad if v.grad is not None else torch.zeros_like(v)
        assert torch.allclose(g, torch.tensor(Z[f"{name}_g_{k}"]), rtol=1e-4, atol=1e-4), (name, k)


def test_adjacency_has_bug_edges():
    A, _, _ = adjacency(False)
    assert (A > 1).any() or not torch.equal(A, A.T), "fixture should exercise duplicate / asymmetric (PS.cpp:401) edges"


@pytest.mark.parametrize("name,se,fn", [
    ("sage", False, lambda p, x, A: R.sage(p, x, A, "relu")), ("sage_elu", True, lambda p, x, A: R.sage(p, x, A, "elu")),
    ("gcn", True, lambda p, x, A: R.gcn(p, x, A, "elu")), ("gin", False, lambda p, x, A: R.gin(p, x, A, "relu")),
    ("gat", True, lambda p, x, A: R.gat(p, x, A, "relu", 4)), ("gatscat", True, lambda p, x, A: R.gatscat(p, x, A, "relu", 2))])
def test_layer_restatement(name, se, fn):
    A, _, _ = adjacency(se)
    check(name, lambda p, x: fn(p, x, A))


def test_stacked_layers():
    A, _, _ = adjacency(False)
    sub = lambda p, i: {k[2:]: v for k, v in p.items() if k.startswith(f"{i}.")}
    check("sage2", lambda p, x: R.sage(sub(p, 1), R.sage(sub(p, 0), x, A, "relu"), A, "relu"))
    check("gin2", lambda p, x: R.gin(sub(p, 1), R.gin(sub(p, 0), x, A, "relu"), A, "relu"))


@pytest.mark.parametrize("res,pool", [("none", "center"), ("cat", "center"), ("max", "max"), ("sum", "mean"), ("cat", "sum"), ("none", "sort")])
def test_respool_restatement(res, pool):
    _, tgt, sizes = adjacency(False)
    f = lambda x: [x[:, :16], x[:, 16:32] * 0.5 + 0.1, x[:, 32:48] - 0.2]
    check(f"respool_{res}_{pool}", lambda p, x: R.respool(p, f(x), tgt, sizes, res, pool, "relu", k=5))


def test_deepgnn_restatement():
    for name, aggr, se, act, heads in (("model_sage", "sage", False, "relu", 1), ("model_gat", "gat", True, "elu", 2)):
        A, tgt, _ = adjacency(se)
        check(name, lambda p, x: R.deepgnn(p, x, A, tgt, aggr, act, heads, 3), tol=5e-5)
