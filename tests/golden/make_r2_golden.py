"""Round-2 golden fixtures from the UNMODIFIED reference (needs /root/reference; run here: python tests/golden/make_r2_golden.py).

Writes tests/golden/r2_golden.npz + the reference-written PPR cache files ppr_ref_{neighs,scores}.bin:
  * sage5     BASELINE config C3 at its real width: 5-layer GraphSAGE-256, F=100, C=47 on a 32-root PPR(k=150) batch -- predictions,
              loss and a signature of every gradient (tests.common.grad_signature), weights = tests.common.det_fill
  * gcn3      BASELINE config C2: 3-layer GCN-256 (elu, hops augmentation summed into the features), F=128, C=40 on a 32-root
              khop(2,10)+hops batch of the S-arxiv stand-in
  * ens       EnsembleAggregator (shaDow/layers.py:236-296) forward + gradients, and a 2-branch DeepGNN
  * sym_drop  adj_norm_sym / adj_norm_rw with an injected dropedge mask (graph_utils.py:81-95,112-123): the reference functions run with
              their random draw replaced by a fixed index list
  * ppr bins  the files `preproc_ppr_approximate(..., fname_neighs, fname_scores)` of oracle/_ref writes for the sampler_golden graph
The batches themselves come from the reference-pinned C oracle (oracle/oracle_sampler.c); the GPU tests re-sample them with the CUDA sampler.
"""
import os
import sys

import numpy as np
import scipy.sparse as sp
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O                                   # noqa: E402
from tests.common import Golden, det_fill, grad_signature        # noqa: E402
from tests.golden import ref_shim                                # noqa: E402
from tests.golden.make_r2_inputs import sage5_inputs, gcn3_inputs, SAGE5_CFG, GCN3_CFG   # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "r2_golden.npz")

ARCH_SAGE5 = dict(num_layers=5, num_cls_layers=1, heads=1, branch_sharing=False, dim=256, act="relu", layer_norm="norm_feat",
                  feature_augment_ops="sum", aggr="sage", residue="none", pooling="center", loss="softmax", ensemble_act="leakyrelu")
ARCH_GCN3 = dict(ARCH_SAGE5, num_layers=3, aggr="gcn", act="elu")


def features(n, F, seed):
    return np.random.default_rng(seed).standard_normal((n, F), dtype=np.float32)


def hop_onehot_ref(graph_mod, hop, dim):
    """what one_batch builds for `hops` (shaDow/minibatch.py:473-477): EntityEncoding.hop2onehot_vec on the int64 list->array column"""
    enc = graph_mod.EntityEncoding(hop=np.asarray(hop.tolist()), validate=False)
    return enc.hop2onehot_vec(7, return_type="tensor").type(torch.float32)


def main():
    ref = ref_shim.load()
    L, M, GU, GR = ref["layers"], ref["models"], ref["graph_utils"], ref["graph"]
    store = {}

    def run_model(tag, model, col, x, labels, C, aug=None, x_grad=True):
        n = col["indptr"].size - 1
        adj = sp.csr_matrix((np.ones(col["indices"].size, np.float32), col["indices"], col["indptr"]), shape=(n, n))
        det_fill(model)
        model.eval()
        xt = torch.tensor(x, requires_grad=x_grad)       # (the reference adds the aug embedding into the features IN PLACE, models.py:189)
        preds, _ = model(0, [xt], [adj], [np.asarray(col["target"])], torch.as_tensor(col["size_subg"]).view(1, -1), [aug or {}], 0.0)
        onehot = torch.nn.functional.one_hot(torch.as_tensor(labels).long(), C)
        loss = model._loss(preds, onehot)
        loss.backward()
        for f in ("indptr", "indices", "target", "size_subg", "node"):
            store[f"{tag}_{f}"] = np.asarray(col[f], np.int64)
        store[f"{tag}_preds"] = preds.detach().numpy()
        store[f"{tag}_loss"] = np.float64(loss.item())
        sig = grad_signature([(pn, p.grad.numpy() if p.grad is not None else np.zeros(tuple(p.shape), np.float32)) for pn, p in model.named_parameters()] +
                             ([("input_x", xt.grad.numpy())] if x_grad else []))
        for k, v in sig.items():
            store[f"{tag}_g|{k}"] = v
        print(tag, "n", n, "e", col["indices"].size, "loss", loss.item(), "params", sum(p.numel() for p in model.parameters()))

    # ---- sage5 (C3 width) ----
    indptr, indices, targets, tables = sage5_inputs(O)
    o = O.OracleSampler(indptr, indices, 32, 1, 1)
    o.set_ppr(*tables); o.shuffle_targets(targets)
    col = O.cat_to_block_diagonal(o.sample(O.cfg_from_cpp_config(SAGE5_CFG)).subgraphs())
    n = col["indptr"].size - 1
    labels = np.random.default_rng(5).integers(0, 47, 32)
    store["sage5_labels"] = labels
    tp = dict(dropout=0.0, dropedge=0.0, lr=0.002, ensemble_dropout="none")
    run_model("sage5", M.DeepGNN(100, 100, 47, 0, ARCH_SAGE5, [], 1, tp, "node"), col, features(n, 100, 21), labels, 47)

    # ---- gcn3 (C2) ----
    indptr, indices, targets = gcn3_inputs()
    o = O.OracleSampler(indptr, indices, 32, 1, 1)
    o.shuffle_targets(targets)
    subs = o.sample(O.cfg_from_cpp_config(GCN3_CFG, aug=("hops",))).subgraphs()
    col = O.cat_to_block_diagonal(subs)
    hop = np.concatenate([s["hop"] for s in subs])
    store["gcn3_hop"] = hop.astype(np.int64)
    n = col["indptr"].size - 1
    labels = np.random.default_rng(6).integers(0, 40, 32)
    store["gcn3_labels"] = labels
    aug = {"hops": hop_onehot_ref(GR, hop, 7)}
    store["gcn3_hop_onehot"] = aug["hops"].numpy()
    run_model("gcn3", M.DeepGNN(128, 128, 40, 0, ARCH_GCN3, [("hops", 7)], 1, tp, "node"), col, features(n, 128, 22), labels, 40, aug=aug, x_grad=False)

    # ---- EnsembleAggregator alone + a 2-branch model ----
    ens = det_fill(L.EnsembleAggregator(16, 16, 3, dropout=0.0, act="leakyrelu", type_dropout="none")).eval()
    Xs = [torch.tensor(features(9, 16, 30 + i), requires_grad=True) for i in range(3)]
    y = ens(list(Xs))
    w = torch.tensor(features(9, 16, 40))
    (y * w).sum().backward()
    store["ens_out"] = y.detach().numpy()
    for i, X in enumerate(Xs):
        store[f"ens_dx{i}"] = X.grad.numpy()
    for pn, p in ens.named_parameters():
        store[f"ens_g_{pn}"] = p.grad.numpy()
    # 2 branches (khop without / with self edges of the same roots), sage, dim 16
    from shadow_gnn_b200.synth import small_parity_graph
    ip2, ix2 = small_parity_graph(400, 8, 5, self_loops=10)
    t2 = np.random.default_rng(1).permutation(ip2.size - 3)[:6].astype(np.uint32)
    cols = []
    for se in (False, True):
        o = O.OracleSampler(ip2, ix2, 6, 1, 3)
        o.shuffle_targets(t2)
        cols.append(O.cat_to_block_diagonal(o.sample(O.make_cfg("khop", depth=2, budget=4, add_self_edge=se)).subgraphs()))
    arch2 = dict(ARCH_SAGE5, num_layers=2, dim=16)
    m2 = det_fill(M.DeepGNN(12, 12, 5, 0, arch2, [], 2, dict(dropout=0.0, dropedge=0.0, lr=0.01, ensemble_dropout="none"), "node")).eval()
    xs, adjs = [], []
    for bi, c in enumerate(cols):
        nb_ = c["indptr"].size - 1
        xs.append(torch.tensor(features(nb_, 12, 50 + bi), requires_grad=True))
        adjs.append(sp.csr_matrix((np.ones(c["indices"].size, np.float32), c["indices"], c["indptr"]), shape=(nb_, nb_)))
        for f in ("indptr", "indices", "target", "size_subg"):
            store[f"ens2_b{bi}_{f}"] = np.asarray(c[f], np.int64)
    sizes = torch.stack([torch.as_tensor(c["size_subg"]) for c in cols], 0)
    preds, _ = m2(0, xs, adjs, [np.asarray(c["target"]) for c in cols], sizes, [{}, {}], 0.0)
    wl = torch.tensor(features(6, 5, 60))
    (preds * wl).sum().backward()
    store["ens2_preds"] = preds.detach().numpy()
    for k, v in grad_signature([(pn, p.grad.numpy() if p.grad is not None else np.zeros(tuple(p.shape), np.float32)) for pn, p in m2.named_parameters()]).items():
        store[f"ens2_g|{k}"] = v

    # ---- dropedge with an injected mask: the reference's own adj_norm_sym / adj_norm_rw, random draw replaced ----
    rng = np.random.default_rng(8)
    nA = 300
    src = rng.integers(0, nA, 1500); dst = rng.integers(0, nA, 1500)
    A = sp.coo_matrix((np.ones(3000 + nA), (np.concatenate([src, dst, np.arange(nA)]), np.concatenate([dst, src, np.arange(nA)]))), shape=(nA, nA)).tocsr()
    A.data[:] = 1.0                       # symmetric, self loops, binary, sorted rows (what the sampler hands to GCN)
    A.sort_indices()
    drop_idx = rng.integers(0, A.nnz, int(A.nnz * 0.3))
    store["drop_indptr"], store["drop_indices"], store["drop_idx"] = A.indptr.astype(np.int64), A.indices.astype(np.int64), drop_idx.astype(np.int64)
    real_choice = np.random.choice
    np.random.choice = lambda size, cnt: drop_idx if (size == A.nnz and cnt == drop_idx.size) else real_choice(size, cnt)
    try:
        S = GU.adj_norm_sym(A.copy(), sort_indices=True, dropedge=0.3)
    finally:
        np.random.choice = real_choice
    S = S.tocsr(); S.sort_indices()
    dense = np.asarray(S.todense(), np.float32)
    rows = np.repeat(np.arange(nA), np.diff(A.indptr))
    store["drop_sym_vals"] = dense[rows, A.indices]                     # value at every structural position (0 where dropped)
    coo = GU.coo_scipy2torch(A.tocoo())
    real_rand = torch.rand
    u = (torch.as_tensor(drop_idx, dtype=torch.float64) + 0.5) / A.nnz     # floor(u * nnz) == drop_idx
    torch.rand = lambda cnt: u.to(torch.float32) if cnt == drop_idx.size else real_rand(cnt)
    try:
        assert torch.equal(torch.floor(u.to(torch.float32) * A.nnz).long(), torch.as_tensor(drop_idx))
        R = GU.adj_norm_rw(coo, dropedge=0.3)
    finally:
        torch.rand = real_rand
    store["drop_rw_vals"] = R._values().numpy()
    assert np.array_equal(R._indices()[0].numpy(), rows) and np.array_equal(R._indices()[1].numpy(), A.indices)

    # ---- reference-written PPR cache files for the sampler_golden graph ----
    G = Golden()
    p = G.meta["ppr"]
    mod = O.load_ref()
    fnn, fns = os.path.join(HERE, "ppr_ref_neighs.bin"), os.path.join(HERE, "ppr_ref_scores.bin")
    for f in (fnn, fns):
        if os.path.exists(f):
            os.remove(f)
    r = mod.ParallelSampler(G.indptr.tolist(), G.indices.tolist(), [], 4, 1, True, True, [], 1, "", "", "", 1)
    r.preproc_ppr_approximate(list(range(G.indptr.size - 1)), p["k"], p["alpha"], p["epsilon"], fnn, fns)
    print("ppr bins:", os.path.getsize(fnn), os.path.getsize(fns))

    np.savez_compressed(OUT, **store)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(store), "arrays")


if __name__ == "__main__":
    main()
