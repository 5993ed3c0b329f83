"""CPU tests: the fp32 torch restatement (oracle/layers_ref.py) against the round-2 golden outputs of the UNMODIFIED reference
(tests/golden/r2_golden.npz, made by tests/golden/make_r2_golden.py): 5-layer SAGE-256 (C3 width), 3-layer GCN-256 with the hops
augmentation (C2), EnsembleAggregator, dropedge with an injected draw; and the reference-written PPR cache files against the golden tables."""
import os

import numpy as np
import torch

from oracle import layers_ref as R
from tests.common import Golden, det_fill, grad_signature

HERE = os.path.dirname(os.path.abspath(__file__))
Z = np.load(os.path.join(HERE, "golden", "r2_golden.npz"))


class _Shape(torch.nn.Module):
    """parameter container with the reference's names / shapes, filled by det_fill (no reference code needed)"""

    def __init__(self, shapes):
        super().__init__()
        self.names = list(shapes)
        for i, (k, sh) in enumerate(shapes.items()):
            self.register_parameter(f"p{i}", torch.nn.Parameter(torch.zeros(sh)))

    def named_parameters(self, *a, **k):
        for i, n in enumerate(self.names):
            yield n, getattr(self, f"p{i}")


def model_shapes(F, D, C, L, aggr, aug_dims=()):
    sh = {}
    for ia, d in enumerate(aug_dims):
        sh[f"aug_layers.0.{ia}.weight"] = (F, d); sh[f"aug_layers.0.{ia}.bias"] = (F,)
    for l in range(L):
        din = F if l == 0 else D
        pre = f"conv_layers.0.{l}."
        if aggr == "sage":
            sh[pre + "offset"] = (2, D); sh[pre + "scale"] = (2, D)
            for b in ("self", "neigh"):
                sh[pre + f"f_lin_{b}.weight"] = (D, din); sh[pre + f"f_lin_{b}.bias"] = (D,)
        else:
            sh[pre + "offset"] = (1, D); sh[pre + "scale"] = (1, D)
            sh[pre + "f_lin.weight"] = (D, din); sh[pre + "f_lin.bias"] = (D,)
    sh["classifier.0.offset"] = (1, C); sh["classifier.0.scale"] = (1, C)
    sh["classifier.0.f_lin.weight"] = (C, D); sh["classifier.0.f_lin.bias"] = (C,)
    return sh


def _features(n, F, seed):
    return torch.from_numpy(np.random.default_rng(seed).standard_normal((n, F), dtype=np.float32))


def _check_model(tag, F, D, C, L, aggr, act, seed, aug=None, x_grad=True):
    ip, ix, tgt = Z[f"{tag}_indptr"], Z[f"{tag}_indices"], torch.as_tensor(Z[f"{tag}_target"])
    n = ip.size - 1
    A = R.dense_counts(ip, ix, n)
    holder = det_fill(_Shape(model_shapes(F, D, C, L, aggr, [a.shape[1] for a in (aug or [])])))
    p = dict(holder.named_parameters())
    x = _features(n, F, seed).requires_grad_(x_grad)
    preds = R.deepgnn(p, x, A, tgt, aggr, act, 1, L, aug=aug)
    want = torch.as_tensor(Z[f"{tag}_preds"])
    err = (preds - want).abs().max().item() / (want.abs().max().item() + 1e-6)
    assert err < 1e-4, (tag, err)
    loss = torch.nn.functional.cross_entropy(preds, torch.as_tensor(Z[f"{tag}_labels"]).long())
    assert abs(loss.item() - float(Z[f"{tag}_loss"])) < 1e-4 * abs(float(Z[f"{tag}_loss"]))
    loss.backward()
    sig = grad_signature([(k, v.grad.numpy() if v.grad is not None else np.zeros(tuple(v.shape), np.float32)) for k, v in p.items()] +
                         ([("input_x", x.grad.numpy())] if x_grad else []))
    for k, v in sig.items():
        w = Z[f"{tag}_g|{k}"]
        scale = float(Z[f"{tag}_g|{k.split('|')[0]}|norm"]) + 1e-12
        assert np.abs(np.asarray(v, np.float64) - w).max() <= 2e-3 * scale + 1e-7, (tag, k, v, w)


def test_sage5x256_restatement_vs_reference_golden():
    _check_model("sage5", 100, 256, 47, 5, "sage", "relu", 21)


def test_gcn3_hops_restatement_vs_reference_golden():
    _check_model("gcn3", 128, 256, 40, 3, "gcn", "elu", 22, aug=[torch.as_tensor(Z["gcn3_hop_onehot"])], x_grad=False)


def test_hop_onehot_matches_reference_encoding():
    """columns of EntityEncoding.hop2onehot_vec as the reference's one_batch builds them (graph.py:134-147; unreachable = 0xFFFFFFFF)"""
    hop = Z["gcn3_hop"]
    want = Z["gcn3_hop_onehot"]
    got = np.zeros_like(want)
    ok = hop <= 5
    got[np.nonzero(ok)[0], hop[ok] + 1] = 1
    got[hop >= 255, 0] = 1
    assert np.array_equal(got, want)


def test_ensemble_aggregator_restatement_vs_reference_golden():
    holder = det_fill(_Shape({"f_lin.weight": (16, 16), "f_lin.bias": (16,), "q": (16,)}))
    p = dict(holder.named_parameters())
    Xs = [_features(9, 16, 30 + i).requires_grad_(True) for i in range(3)]
    y = R.ensemble_aggregator(p, Xs)
    assert torch.allclose(y, torch.as_tensor(Z["ens_out"]), rtol=1e-5, atol=1e-6)
    (y * _features(9, 16, 40)).sum().backward()
    for i, X in enumerate(Xs):
        assert torch.allclose(X.grad, torch.as_tensor(Z[f"ens_dx{i}"]), rtol=1e-4, atol=1e-6)
    for k, v in p.items():
        assert torch.allclose(v.grad, torch.as_tensor(Z[f"ens_g_{k}"]), rtol=1e-4, atol=1e-6), k


def test_dropedge_restatement_vs_reference_with_injected_draw():
    ip, ix, drop = Z["drop_indptr"], Z["drop_indices"], Z["drop_idx"]
    assert np.allclose(R.rw_vals_dropedge(ip, drop), Z["drop_rw_vals"], rtol=1e-6, atol=0)
    assert np.allclose(R.sym_vals_dropedge(ip, ix, drop), Z["drop_sym_vals"], rtol=1e-6, atol=1e-8)
    assert (Z["drop_sym_vals"] == 0).sum() >= np.unique(drop).size          # symmetric survival drops at least the drawn edges


def test_reference_written_ppr_cache_matches_golden_tables():
    """the committed files are what oracle/_ref wrote (PS.cpp:94-139): header + rows == the golden tables"""
    G = Golden()
    for fn, dt, flat in (("ppr_ref_neighs.bin", np.uint32, G.ppr_neighs), ("ppr_ref_scores.bin", np.float32, G.ppr_scores)):
        raw = open(os.path.join(HERE, "golden", fn), "rb").read()
        alpha, eps = np.frombuffer(raw, np.float32, 2, 0)
        k, cnt = int(np.frombuffer(raw, np.int32, 1, 8)[0]), int(np.frombuffer(raw, np.uint32, 1, 12)[0])
        assert alpha == np.float32(1) - np.float32(G.meta["ppr"]["alpha"]) and k == G.meta["ppr"]["k"] and cnt == G.indptr.size - 1
        body, pos = np.frombuffer(raw, np.uint32, -1, 16), 0
        for v in range(cnt):
            ln = int(body[pos]); pos += 1
            a, b = int(G.ppr_ptr[v]), int(G.ppr_ptr[v + 1])
            assert ln == b - a and body[pos:pos + ln].tobytes() == flat[a:b].astype(dt).tobytes(), (fn, v)
            pos += ln
        assert pos == body.size
