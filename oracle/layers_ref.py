"""Plain PyTorch fp32 restatement of the reference's layer math (shaDow/layers.py, fe/graph_utils.py) on DENSE adjacency.
TEST INFRASTRUCTURE ONLY (the checker of the floating-point kernels; tolerance 1e-3 relative, BASELINE.json north_star).

Pinned against the unmodified reference layers through tests/golden/layers_golden.npz (tests/test_layers_oracle.py).
Duplicated edges (uncoalesced COO) count with multiplicity: summed in the products and in the degrees,
separate terms in the GAT softmax -- which a dense matrix of edge COUNTS reproduces exactly.
"""
import torch
import torch.nn.functional as F

ACT = {"relu": F.relu, "I": lambda x: x, "elu": F.elu, "tanh": torch.tanh, "leakyrelu": lambda x: F.leaky_relu(x, 0.2)}


def dense_counts(indptr, indices, n):
    """A[i,j] = number of (i,j) entries in the CSR"""
    A = torch.zeros(n, n)
    rows = torch.repeat_interleave(torch.arange(n), torch.as_tensor(indptr[1:] - indptr[:-1]))
    A.index_put_((rows, torch.as_tensor(indices).long()), torch.ones(len(indices)), accumulate=True)
    return A


def norm_feat(h, scale, offset):
    """layers.py:329-338"""
    mean = h.mean(1, keepdim=True)
    var = h.var(1, unbiased=False, keepdim=True) + 1e-9
    return (h - mean) * scale * torch.rsqrt(var) + offset


def adj_rw(A):          # graph_utils.py:81-95 without dropedge
    return A / A.sum(1, keepdim=True).clamp(min=1)


def adj_sym(A):         # graph_utils.py:139-142
    d = A.sum(1).clamp(min=1).pow(-0.5)
    return d[:, None] * A * d[None, :]


def sage(p, x, A, act):                                     # layers.py:454-484 ; slot 0 = self, 1 = neighbour
    hs = ACT[act](x @ p["f_lin_self.weight"].T + p["f_lin_self.bias"])
    hn = ACT[act]((adj_rw(A) @ x) @ p["f_lin_neigh.weight"].T + p["f_lin_neigh.bias"])
    return norm_feat(hs, p["scale"][0], p["offset"][0]) + norm_feat(hn, p["scale"][1], p["offset"][1])


def gcn(p, x, A, act):                                      # layers.py:423-436
    return norm_feat(ACT[act]((adj_sym(A) @ x) @ p["f_lin.weight"].T + p["f_lin.bias"]), p["scale"][0], p["offset"][0])


def gin(p, x, A, act):                                      # layers.py:508-527
    h = A @ x + (1 + p["eps"]) * x
    h = F.relu(h @ p["mlp.0.weight"].T + p["mlp.0.bias"]) @ p["mlp.2.weight"].T + p["mlp.2.bias"]
    return norm_feat(ACT[act](h), p["scale"][0], p["offset"][0])


def _attend(A, e_self, e_neigh, h):
    """softmax over the row's edges of e_self[i] + e_neigh[j], multiplicity-aware; clamp(sum, 1e-10) (layers.py:568-582)"""
    e = e_self[:, None] + e_neigh[None, :]
    mx = torch.where(A > 0, e, torch.full_like(e, -float("inf"))).max(1, keepdim=True).values
    mx = torch.where(torch.isfinite(mx), mx, torch.zeros_like(mx))
    u = torch.exp(e - mx) * A
    return (u @ h) / u.sum(1, keepdim=True).clamp(min=1e-10)


def gat(p, x, A, act, heads):                               # layers.py:602-626 ; norm slot [0,k] = neighbour, [1,k] = self
    n = x.shape[0]
    hs = ACT[act](x @ p["f_lin.0.weight"].T + p["f_lin.0.bias"]).view(n, heads, -1)
    hn = ACT[act](x @ p["f_lin.1.weight"].T + p["f_lin.1.bias"]).view(n, heads, -1)
    outs = []
    for k in range(heads):
        es = F.leaky_relu(hs[:, k] @ p["attention"][0, k], 0.2)
        en = F.leaky_relu(hn[:, k] @ p["attention"][1, k], 0.2)
        agg = _attend(A, es, en, hn[:, k])
        outs.append(norm_feat(hs[:, k], p["scale"][1, k], p["offset"][1, k]) + norm_feat(agg, p["scale"][0, k], p["offset"][0, k]))
    return torch.cat(outs, 1) / 2


def gatscat(p, x, A, act, heads):                           # layers.py:711-743
    n = x.shape[0]
    src = (x @ p["f_lin.0.weight"].T + p["f_lin.0.bias"]).view(n, heads, -1)
    el = F.leaky_relu((src * p["attention"]).sum(-1), 0.2)
    agg = torch.cat([_attend(A, torch.zeros(n), el[:, k], src[:, k]) for k in range(heads)], 1)
    return norm_feat(ACT[act](agg + x @ p["f_lin.1.weight"].T + p["f_lin.1.bias"]), p["scale"][0], p["offset"][0])


def segment_pool(x, sizes, mode):
    out, off = [], 0
    for s in sizes.tolist():
        seg = x[off:off + s]
        out.append({"max": lambda t: t.max(0).values, "mean": lambda t: t.mean(0), "sum": lambda t: t.sum(0)}[mode](seg))
        off += s
    return torch.stack(out)


def sort_pool(x, sizes, k):
    out, off = [], 0
    for s in sizes.tolist():
        seg = x[off:off + s]
        seg = seg[torch.argsort(seg[:, -1], descending=True, stable=True)][:k]
        if seg.shape[0] < k:
            seg = torch.cat([seg, seg.new_zeros(k - seg.shape[0], seg.shape[1])])
        out.append(seg.reshape(-1)); off += s
    return torch.stack(out)


def respool(p, feats, targets, sizes, type_res, type_pool, act, k=None):      # layers.py:154-199
    def res(fl):
        if type_res in ("cat", "concat"):
            return torch.cat(fl, 1)
        st = torch.stack(fl, 0)
        return st.sum(0) if type_res == "sum" else st.max(0).values
    if type_pool == "center":
        if type_res == "none":
            return feats[-1][targets]
        x = res([f[targets] for f in feats])
    else:
        if type_res == "none":
            pin, root = feats[-1], feats[-1][targets]
            pool = segment_pool(pin, sizes, type_pool) if type_pool != "sort" else None
        else:
            root = res([f[targets] for f in feats])
            pin = res(feats) if type_pool == "sort" else None
            pool = res([segment_pool(f, sizes, type_pool) for f in feats]) if type_pool != "sort" else None
        if type_pool == "sort":
            pool = ACT[act](sort_pool(pin, sizes, k) @ p["nn_pool.1.weight"].T + p["nn_pool.1.bias"])
        x = torch.cat([root, pool], 1)
    return norm_feat(ACT[act](x @ p["nn.1.weight"].T + p["nn.1.bias"]), p["scale"], p["offset"])


def branch_embedding(p, x, A, targets, aggr, act, heads, num_layers, branch=0, aug=None):
    """one branch of DeepGNN.forward up to the L2-normalised root embedding (models.py:181-201): aug embeddings summed into the
    features (feature_augment_ops == 'sum', models.py:186-189), conv stack, center pooling without residue"""
    h = x
    for ia, onehot in enumerate(aug or []):
        h = h + onehot @ p[f"aug_layers.{branch}.{ia}.weight"].T + p[f"aug_layers.{branch}.{ia}.bias"]
    for l in range(num_layers):
        pre = f"conv_layers.{branch}.{l}."
        pl = {k[len(pre):]: v for k, v in p.items() if k.startswith(pre)}
        h = {"sage": lambda: sage(pl, h, A, act), "gcn": lambda: gcn(pl, h, A, act), "gat": lambda: gat(pl, h, A, act, heads)}[aggr]()
    return F.normalize(h[targets], p=2, dim=1)


def classifier(p, emb):
    pc = {k[len("classifier.0."):]: v for k, v in p.items() if k.startswith("classifier.0.")}
    return norm_feat(emb @ pc["f_lin.weight"].T + pc["f_lin.bias"], pc["scale"][0], pc["offset"][0])


def deepgnn(p, x, A, targets, aggr, act, heads, num_layers, aug=None):         # models.py:169-204, 1 branch, center pooling, no residue
    return classifier(p, branch_embedding(p, x, A, targets, aggr, act, heads, num_layers, 0, aug))


def ensemble_aggregator(p, Xs, act="leakyrelu"):                             # layers.py:276-289 (type_dropout none / eval)
    omega = torch.cat([(ACT[act](X @ p["f_lin.weight"].T + p["f_lin.bias"]) @ p["q"]).view(-1, 1) for X in Xs], 1)
    w = F.softmax(omega, dim=1)
    return sum(w[:, i:i + 1] * X for i, X in enumerate(Xs))


def deepgnn_ensemble(p, xs, As, targets, aggr, act, heads, num_layers):        # models.py:179-203 with >= 2 branches
    embs = [branch_embedding(p, x, A, t, aggr, act, heads, num_layers, b) for b, (x, A, t) in enumerate(zip(xs, As, targets))]
    pe = {k[len("ensembler."):]: v for k, v in p.items() if k.startswith("ensembler.")}
    return classifier(p, ensemble_aggregator(pe, embs))


def rw_vals_dropedge(indptr, drop_idx):
    """adj_norm_rw on the torch path with a given draw (graph_utils.py:81-95): value of every CSR position"""
    import numpy as np
    e = int(indptr[-1])
    v = np.ones(e, np.float32)
    v[np.asarray(drop_idx)] = 0
    rows = np.repeat(np.arange(len(indptr) - 1), np.diff(indptr))
    deg = np.bincount(rows, weights=v, minlength=len(indptr) - 1).astype(np.float32)
    return v / np.maximum(deg, 1)[rows]


def sym_vals_dropedge(indptr, indices, drop_idx):
    """adj_norm_sym with a given draw (graph_utils.py:112-123,139-142) on a structurally symmetric CSR with sorted rows: an edge survives
    iff both directions survived (csr data + csc data == 2), then D^-1/2 A D^-1/2 with D = clip(rowsum, 1)"""
    import numpy as np
    import scipy.sparse as sp
    n, e = len(indptr) - 1, int(indptr[-1])
    m = np.ones(e, np.float32)
    m[np.asarray(drop_idx)] = 0
    adj_m = sp.csr_matrix((m, np.asarray(indices), np.asarray(indptr)), shape=(n, n))
    both = adj_m.data + adj_m.tocsc().data
    v = (both == 2).astype(np.float32)
    rows = np.repeat(np.arange(n), np.diff(indptr))
    d = np.maximum(np.bincount(rows, weights=v, minlength=n), 1).astype(np.float64) ** -0.5
    return (d[rows] * v * d[np.asarray(indices)]).astype(np.float32)
