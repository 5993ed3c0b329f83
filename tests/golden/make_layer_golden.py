"""Generate tests/golden/layers_golden.npz from the UNMODIFIED reference layers (shaDow/layers.py, shaDow/models.py).

Run here (needs /root/reference):  python tests/golden/make_layer_golden.py
Inputs: a block-diagonal batch sampled by the reference-pinned oracle on a small graph, WITHOUT self edges for
sage/gin (so the PS.cpp:401 bug edges / duplicate columns are in the adjacency) and WITH self edges for gcn/gat.
Outputs: eval-mode (dropout = dropedge = 0) forward values and the gradients of sum(out * W) w.r.t. the input features and
every parameter, for each layer type, each ResPool mode and two full DeepGNN models.
"""
import os
import sys

import numpy as np
import scipy.sparse as sp
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O                         # noqa: E402
from shadow_gnn_b200.synth import small_parity_graph    # noqa: E402
from tests.golden import ref_shim                      # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "layers_golden.npz")


def sample_batch(add_self_edge, B=6, seed=3):
    indptr, indices = small_parity_graph(400, 8, 5, self_loops=10)
    o = O.OracleSampler(indptr, indices, B, 1, seed)
    o.shuffle_targets(np.random.default_rng(1).permutation(indptr.size - 3)[:B].astype(np.uint32))
    sub = o.sample(O.make_cfg("khop", depth=2, budget=4, add_self_edge=add_self_edge)).subgraphs()
    return O.cat_to_block_diagonal(sub)


def main():
    ref = ref_shim.load()
    L, M = ref["layers"], ref["models"]
    torch.manual_seed(0)
    store = {}
    batches = {False: sample_batch(False), True: sample_batch(True)}
    for key, c in batches.items():
        for f in ("indptr", "indices", "target", "size_subg"):
            store[f"adj{int(key)}_{f}"] = np.asarray(c[f], np.int64)

    def run(name, module, self_edge, dim_in, post=None):
        c = batches[self_edge]
        n = c["indptr"].size - 1
        adj = sp.csr_matrix((np.ones(c["indices"].size, np.float32), c["indices"], c["indptr"]), shape=(n, n))
        module.eval()
        with torch.no_grad():           # non-trivial norm parameters
            for pn, p in module.named_parameters():
                if pn.endswith("scale") or pn.endswith("offset"):
                    p.add_(0.3 * torch.randn_like(p))
        x = torch.randn(n, dim_in, requires_grad=True)
        sizes = torch.as_tensor(c["size_subg"])
        out = module((x, adj, False, 0.0), sizes)[0] if post is None else post(module, x, adj, sizes, c)
        w = torch.randn_like(out)
        (out * w).sum().backward()
        store[f"{name}_x"] = x.detach().numpy(); store[f"{name}_w"] = w.numpy(); store[f"{name}_out"] = out.detach().numpy()
        store[f"{name}_dx"] = x.grad.numpy()
        for pn, p in module.named_parameters():
            store[f"{name}_p_{pn}"] = p.detach().numpy()
            store[f"{name}_g_{pn}"] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy()

    run("sage", L.GraphSAGE(12, 16, act="relu"), False, 12)
    run("sage_elu", L.GraphSAGE(12, 16, act="elu"), True, 12)
    run("gcn", L.GCN(12, 16, act="elu"), True, 12)
    run("gin", L.GIN(12, 16, act="relu", eps=0.1), False, 12)
    run("gat", L.GAT(12, 16, act="relu", mulhead=4), True, 12)
    run("gatscat", L.GATScatter(12, 16, act="relu", mulhead=2), True, 12)

    def two(mod, x, adj, sizes, c):          # two stacked layers (is_normed path)
        t = mod[0]((x, adj, False, 0.0), sizes)
        return mod[1](t, sizes)[0]
    run("sage2", torch.nn.Sequential(L.GraphSAGE(12, 16), L.GraphSAGE(16, 16)), False, 12, post=two)
    run("gin2", torch.nn.Sequential(L.GIN(12, 16), L.GIN(16, 16)), False, 12, post=two)
    for res, pool in (("none", "center"), ("cat", "center"), ("max", "max"), ("sum", "mean"), ("cat", "sum"), ("none", "sort-5")):
        pname = pool.split("-")[0]
        args_pool = {"k": 5} if pname == "sort" else {}
        rp = L.ResPool(16, 16, 3, res, pname, dropout=0.0, act="relu", args_pool=args_pool)

        def pool_post(mod, x, adj, sizes, c):
            feats = [x[:, :16], x[:, 16:32] * 0.5 + 0.1, x[:, 32:48] - 0.2]
            return mod(feats, torch.as_tensor(c["target"]), sizes)
        run(f"respool_{res}_{pname}", rp, False, 48, post=pool_post)
    for name, aggr, self_edge, act, heads in (("model_sage", "sage", False, "relu", 1), ("model_gat", "gat", True, "elu", 2)):
        arch = dict(num_layers=3, num_cls_layers=1, heads=heads, branch_sharing=False, dim=16, act=act, layer_norm="norm_feat",
                    feature_augment_ops="sum", aggr=aggr, residue="none", pooling="center", loss="softmax", ensemble_act="leakyrelu")
        model = M.DeepGNN(12, 12, 5, 0, arch, [], 1, dict(dropout=0.0, dropedge=0.0, lr=0.01, ensemble_dropout="none"), "node")

        def model_post(mod, x, adj, sizes, c):
            preds, _ = mod(1, [x], [adj], [np.asarray(c["target"])], sizes.view(1, -1), [{}], 0.0)
            return preds
        run(name, model, self_edge, 12, post=model_post)
    np.savez_compressed(OUT, **store)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(store), "arrays")


if __name__ == "__main__":
    main()
