"""Generate tests/golden/loader_golden.npz: outputs of the UNMODIFIED reference loader
(para_graph_sampler/graph_engine/frontend/loader.py: load_data) on a small synthetic dataset in shaDow's on-disk format, for the
configurations the node-task YAMLs use.  Needs /root/reference (imported through tests/golden/ref_shim.py); the test
(tests/test_loader.py) rebuilds the same dataset files with `write_dataset` and compares shadow_gnn_b200.loader.load_data with this fixture.

    python tests/golden/make_loader_golden.py
"""
import os
import pickle
import sys
import tempfile

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

NAME = "arxiv"          # any node dataset known to the reference's DATA_ZOO
CONFIGS = [dict(transductive=True, to_undirected=True, norm_feat=True), dict(transductive=True, to_undirected=False, norm_feat=False),
           dict(transductive=False, to_undirected=True, norm_feat=True), dict(transductive=False, to_undirected=False, norm_feat=True)]


def write_dataset(root, seed=3, n=300, f=8, m=2400, with_bin=False, prestored_undirected=False):
    """a directed random graph (a few duplicate-direction pairs, isolated nodes, a zero-variance feature column) + split/labels/features"""
    rng = np.random.default_rng(seed)
    d = os.path.join(root, NAME)
    os.makedirs(d, exist_ok=True)
    src, dst = rng.integers(0, n - 5, m), rng.integers(0, n - 5, m)           # the last 5 nodes stay isolated
    keep = src != dst
    adj = sp.csr_matrix((np.ones(keep.sum(), dtype=bool), (src[keep], dst[keep])), shape=(n, n))
    adj.data[:] = True
    perm = rng.permutation(n)
    split = {0: np.sort(perm[:180]), 1: np.sort(perm[180:240]), 2: np.sort(perm[240:])}
    tr = np.zeros(n, dtype=bool)
    tr[split[0]] = True
    coo = adj.tocoo()
    kt = tr[coo.row] & tr[coo.col]
    adj_train = sp.csr_matrix((np.ones(kt.sum(), dtype=bool), (coo.row[kt], coo.col[kt])), shape=(n, n))
    feats = rng.normal(2.0, 3.0, (n, f)).astype(np.float32)
    feats[:, 3] = 1.5                                                       # zero variance
    labels = rng.integers(0, 7, n).astype(np.int64)
    sp.save_npz(os.path.join(d, "adj_full_raw.npz"), adj)
    sp.save_npz(os.path.join(d, "adj_train_raw.npz"), adj_train)
    np.save(os.path.join(d, "feat_full.npy"), feats)
    np.save(os.path.join(d, "label_full.npy"), labels)
    np.save(os.path.join(d, "split.npy"), split, allow_pickle=True)
    if prestored_undirected or with_bin:
        und = ((adj + adj.T) > 0).tocsr()
        und.sort_indices()
        if prestored_undirected:
            with open(os.path.join(d, "adj_full_undirected.npy"), "wb") as fh:      # data_converter.py:456-459: a pickle with a .npy name
                pickle.dump({"indptr": und.indptr.astype(np.uint32), "indices": und.indices.astype(np.uint32)}, fh, protocol=4)
        if with_bin:
            os.makedirs(os.path.join(d, "cpp"), exist_ok=True)
            und.indptr.astype(np.uint32).tofile(os.path.join(d, "cpp", "adj_full_undirected_indptr.bin"))
            und.indices.astype(np.uint32).tofile(os.path.join(d, "cpp", "adj_full_undirected_indices.bin"))
    return d


def main():
    from tests.golden import ref_shim
    ref_shim.load()
    import torch
    torch.set_default_dtype(torch.float32)
    cwd = os.getcwd()
    os.chdir(ref_shim.REF)
    try:
        from graph_engine.frontend.loader import load_data
    finally:
        os.chdir(cwd)
    out = {}
    with tempfile.TemporaryDirectory() as root:
        write_dataset(root)
        for ci, cfg in enumerate(CONFIGS):
            g = load_data({"local": root}, NAME, dict(cfg), printf=lambda *a, **k: None)
            for nm, a in (("full", g.adj_full), ("train", g.adj_train)):
                out[f"c{ci}_{nm}_indptr"] = np.asarray(a.indptr).astype(np.int64)
                out[f"c{ci}_{nm}_indices"] = np.asarray(a.indices).astype(np.int64)
                out[f"c{ci}_{nm}_data_all_one"] = np.array([bool(np.all(np.asarray(a.data) == 1))])
            out[f"c{ci}_feat"] = g.feat_full.numpy()
            out[f"c{ci}_label"] = g.label_full.numpy()
            for k in (0, 1, 2):
                out[f"c{ci}_nodes{k}"] = np.asarray(g.node_set[k])
            out[f"c{ci}_bin_none"] = np.array([all(v is None for v in g.bin_adj_files.values())])
    np.savez_compressed(os.path.join(HERE, "loader_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "loader_golden.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
