"""Data ingest for the hot path: shaDow's on-disk dataset format -> what `MinibatchShallowExtractor` consumes.

Drop-in for `graph_engine.frontend.loader.load_data` (para_graph_sampler/graph_engine/frontend/loader.py:18-121) for node tasks:
same arguments, same files (`split.npy`, `label_full.npy`, `feat_full.npy` / `feat_full_norm_{all,train}.npy`,
`adj_{full,train}_{raw,undirected}.{npz,npy}`, `cpp/adj_*_{indptr,indices,data}.bin`), same decisions (transductive vs inductive
adjacency, prestored undirected adjacency preferred, feature standardisation fitted on all / on the training rows), same result
fields (`RawGraph`: adj_full, adj_train, feat_full, label_full, node_set, edge_set, bin_adj_files).

What is different:
  * `to_undirected_csr` (frontend/graph_utils.py:19-45) is a per-row Python loop over `np.union1d` in the reference (minutes for
    ogbn-products); here it is one sparse pattern union with the same result: rows sorted, unique, all-ones boolean data, uint32
    index arrays when they fit (`get_adj_dtype`, graph_utils.py:11-16);
  * the StandardScaler fit/transform (loader.py:104-112) is restated in numpy (float64 statistics, population variance, zero-variance
    columns left unscaled -- sklearn's documented behaviour), so scikit-learn is not needed on the training box;
  * the dataset download / conversion step (`convert2shaDow`, needs `ogb` and the network) is out of scope: missing files raise.
Link-prediction datasets (edge sets) are not handled here (the device minibatch covers node tasks).
"""
import os
from dataclasses import dataclass
from typing import Any, Dict, Optional

import numpy as np
import scipy.sparse as sp

TRAIN, VALID, TEST = 0, 1, 2          # graph_engine/frontend/__init__.py


@dataclass
class RawGraph:
    """frontend/graph.py:14-64"""
    adj_full: sp.csr_matrix
    adj_train: Optional[sp.csr_matrix]
    feat_full: Any
    label_full: Any
    node_set: Optional[Dict[int, np.ndarray]]
    edge_set: Optional[Dict[Any, Any]]
    bin_adj_files: Optional[Dict[int, Optional[Dict[str, str]]]]

    def __post_init__(self):
        if self.feat_full is not None:
            assert self.feat_full.shape[0] == self.num_nodes, \
                f"[RawGraph]: unmatched feature size ({self.feat_full.shape[0]}) and graph size ({self.num_nodes})"
        if self.label_full is not None:
            assert self.label_full.shape[0] == self.num_nodes, \
                f"[RawGraph]: unmatched label size ({self.label_full.shape[0]}) and graph size ({self.num_nodes})"

    @property
    def entity_set(self):
        return self.edge_set if self.node_set is None else self.node_set

    @property
    def num_nodes(self):
        return self.adj_full.indptr.size - 1

    @property
    def num_edges(self):
        return self.adj_full.indices.size


def get_adj_dtype(adj=None, num_nodes=-1, num_edges=-1):
    """graph_utils.py:11-16"""
    if adj is not None:
        num_nodes, num_edges = adj.shape[0], adj.size
    return np.uint32 if max(num_nodes, num_edges) < 2 ** 32 else np.int64


def to_undirected_csr(adj):
    """Pattern of A + A^T as CSR with sorted, unique rows and broadcast all-ones boolean data (graph_utils.py:19-45)."""
    adj = adj.tocsr()
    n = adj.shape[0]
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(adj.indptr.astype(np.int64)))
    cols = adj.indices.astype(np.int64)
    key = np.unique(np.concatenate([rows * n + cols, cols * n + rows]))
    r, c = key // n, key % n
    indptr = np.zeros(n + 1, np.int64)
    np.cumsum(np.bincount(r, minlength=n), out=indptr[1:])
    data = np.broadcast_to(np.ones(1, dtype=bool), c.size)
    und = sp.csr_matrix((data, c, indptr), shape=adj.shape)
    dt = get_adj_dtype(adj=und)
    und.indptr = und.indptr.astype(dt, copy=False)
    und.indices = und.indices.astype(dt, copy=False)
    return und


def _load_adj(prefix, dataset, type_, split_, surfix=""):
    """loader.py:124-149: .npz (scipy) or a pickled {'indptr','indices'[,'data']} dict stored as .npy; None when absent"""
    assert split_ in ("full", "train") and type_ in ("raw", "undirected")
    base = f"{prefix}/{dataset}/adj_{split_}_{type_}{surfix}."
    if os.path.isfile(base + "npz"):
        return sp.load_npz(base + "npz")
    if os.path.isfile(base + "npy"):
        d = np.load(base + "npy", allow_pickle=True)
        if isinstance(d, np.ndarray):
            d = d[()]
        assert isinstance(d, dict)
        indptr, indices = d["indptr"], d["indices"]
        data = d["data"] if "data" in d else np.broadcast_to(np.ones(1, dtype=bool), indices.size)
        n = indptr.size - 1
        return sp.csr_matrix((data, indices, indptr), shape=(n, n))
    return None


def validate_bin_file(bin_adj_files):
    """loader.py:152-160: all-or-nothing; a missing data file becomes ''"""
    for _md, df in bin_adj_files.items():
        assert set(df.keys()) == {"indptr", "indices", "data"}
        if not os.path.isfile(df["indptr"]) or not os.path.isfile(df["indices"]):
            return {m: None for m in bin_adj_files}
        if not os.path.isfile(df["data"]):
            df["data"] = ""
    return bin_adj_files


def standardize(feats, fit_rows=None, chunk=1 << 18):
    """StandardScaler().fit(feats[fit_rows]).transform(feats) (loader.py:104-112): mean / population variance per column accumulated in
    float64 over row chunks (no float64 copy of the matrix: papers100M-scale features would not fit twice), columns with zero variance are
    only centred; the transform runs in the INPUT dtype like sklearn's (float32 in, float32 out)"""
    feats = np.asarray(feats)
    dt = feats.dtype if feats.dtype in (np.float32, np.float64) else np.float64
    rows = np.arange(feats.shape[0]) if fit_rows is None else np.asarray(fit_rows)
    s1 = np.zeros(feats.shape[1], np.float64)
    for a in range(0, rows.size, chunk):
        s1 += feats[rows[a:a + chunk]].astype(np.float64).sum(axis=0)
    mean = s1 / max(rows.size, 1)
    s2 = np.zeros(feats.shape[1], np.float64)
    for a in range(0, rows.size, chunk):
        d = feats[rows[a:a + chunk]].astype(np.float64) - mean
        s2 += (d * d).sum(axis=0)
    scale = np.sqrt(s2 / max(rows.size, 1))
    scale[scale < 10 * np.finfo(np.float64).eps] = 1.0           # sklearn's _handle_zeros_in_scale
    out = np.empty(feats.shape, dt)
    m, sc = mean.astype(dt), scale.astype(dt)
    for a in range(0, feats.shape[0], chunk):
        out[a:a + chunk] = (feats[a:a + chunk].astype(dt) - m) / sc
    return out


def load_data(prefix, dataset, config_data, printf=lambda txt, style=None: print(txt)):
    """loader.py:18-121 for node-classification datasets.  `config_data`: transductive, to_undirected, norm_feat (and coalesce)."""
    import torch
    printf("Loading training data..")
    d = f"{prefix['local']}/{dataset}"
    for f in ("split.npy", "label_full.npy", "feat_full.npy"):
        if not os.path.isfile(f"{d}/{f}"):
            raise FileNotFoundError(f"{d}/{f} is missing: convert the dataset to the shaDow format first (graph_engine.frontend.data_converter; "
                                    "needs ogb + network, out of scope here)")
    role = np.load(f"{d}/split.npy", allow_pickle=True)
    role = role[()] if isinstance(role, np.ndarray) else role
    assert isinstance(role, dict)
    if not all(k in role for k in (TRAIN, VALID, TEST)):
        raise NotImplementedError("link-prediction splits (edge sets) are not handled by this loader")
    node_set = {k: np.asarray(role[k], dtype=np.int64) for k in (TRAIN, VALID, TEST)}
    label_full = torch.from_numpy(np.load(f"{d}/label_full.npy"))
    bin_adj_files = {md: {"indptr": None, "indices": None, "data": None} for md in (TRAIN, VALID, TEST)}

    def fill(mode_, split_, type_):
        for x in ("indptr", "indices", "data"):
            bin_adj_files[mode_][x] = f"{d}/cpp/adj_{split_}_{type_}_{x}.bin"
    if "coalesce" in config_data and not config_data["coalesce"]:
        raise NotImplementedError
    kind = "undirected" if config_data["to_undirected"] else "raw"

    def adj_of(split_):
        a = _load_adj(prefix["local"], dataset, kind, split_)
        if a is None and kind == "undirected":            # no prestored undirected adjacency: convert the raw one
            raw = _load_adj(prefix["local"], dataset, "raw", split_)
            if raw is None:
                raise FileNotFoundError(f"{d}/adj_{split_}_raw.np[yz] is missing")
            a = to_undirected_csr(raw)
        if a is None:
            raise FileNotFoundError(f"{d}/adj_{split_}_{kind}.np[yz] is missing")
        return a
    adj_full = adj_of("full")
    for md in (VALID, TEST):
        fill(md, "full", kind)
    if config_data["transductive"]:
        adj_train = adj_full
        fill(TRAIN, "full", kind)
    else:
        adj_train = adj_of("train")
        fill(TRAIN, "train", kind)
        train_set = set(node_set[TRAIN].tolist())
        touched = set(adj_train.indices.tolist()) if kind == "undirected" else set(adj_train.nonzero()[0].tolist())
        assert touched.issubset(train_set), "the training adjacency touches non-training nodes"
    bin_adj_files = validate_bin_file(bin_adj_files)
    printf(f"SETTING TO {'TRANS' if config_data['transductive'] else 'IN'}DUCTIVE LEARNING", style="red")

    mode_norm = "all" if config_data["transductive"] else "train"
    if config_data["norm_feat"] and os.path.isfile(f"{d}/feat_full_norm_{mode_norm}.npy"):
        feats = np.load(f"{d}/feat_full_norm_{mode_norm}.npy")
        printf(f"Loading '{mode_norm}'-normalized features", style="yellow")
    else:
        feats = np.load(f"{d}/feat_full.npy")
        if config_data["norm_feat"]:
            feats = standardize(feats, None if config_data["transductive"] else node_set[TRAIN])
            printf(f"Normalizing node features (mode = {mode_norm})", style="yellow")
        else:
            printf("Not normalizing node features", style="yellow")
    feats = torch.from_numpy(np.ascontiguousarray(feats)).type(torch.get_default_dtype())
    printf("Done loading training data..")
    return RawGraph(adj_full, adj_train, feats, label_full, node_set, None, bin_adj_files)


def to_undirected_csr_torch(indptr, indices, device=None):
    """`to_undirected_csr` with torch ops on `device` (one sort/unique of 2*nnz 64-bit keys; seconds for ogbn-products on a B200 instead
    of minutes of per-row Python): returns (indptr int64[N+1], indices int32[nnz_und]) on that device -- the uint32 bit patterns the
    sampler borrows through `ParallelSampler.from_device_csr`.  Setup code, not part of any timed region."""
    import torch
    indptr = torch.as_tensor(np.asarray(indptr).astype(np.int64) if not isinstance(indptr, torch.Tensor) else indptr, device=device).long()
    cols = torch.as_tensor(np.asarray(indices).astype(np.int64) if not isinstance(indices, torch.Tensor) else indices, device=device).long()
    n = indptr.numel() - 1
    rows = torch.repeat_interleave(torch.arange(n, device=indptr.device), torch.diff(indptr))
    key = torch.unique(torch.cat([rows * n + cols, cols * n + rows]))
    r = torch.div(key, n, rounding_mode="floor")
    out_idx = (key - r * n).to(torch.int32)
    out_ptr = torch.zeros(n + 1, dtype=torch.int64, device=indptr.device)
    torch.cumsum(torch.bincount(r, minlength=n), 0, out=out_ptr[1:])
    return out_ptr, out_idx


def graph_to_device(raw_graph, device, mode_adj=None):
    """CSR of a `RawGraph` as device tensors for the borrowed-CSR path of `MinibatchShallowExtractor` (adjs[mode] = (indptr, indices)):
    int32 tensors holding the uint32 bit patterns.  `mode_adj` maps mode -> 'train' | 'full' (default: TRAIN -> train, others -> full)."""
    import torch
    mode_adj = mode_adj or {TRAIN: "train", VALID: "full", TEST: "full"}
    cache, out = {}, {}
    for md, which in mode_adj.items():
        adj = raw_graph.adj_train if which == "train" else raw_graph.adj_full
        if id(adj) not in cache:
            assert adj.indices.size < 2 ** 32 and adj.shape[0] < 2 ** 32
            ip = torch.from_numpy(np.ascontiguousarray(adj.indptr).astype(np.uint32).view(np.int32)).to(device)
            ix = torch.from_numpy(np.ascontiguousarray(adj.indices).astype(np.uint32).view(np.int32)).to(device)
            cache[id(adj)] = (ip, ix)
        out[md] = cache[id(adj)]
    return out


def standardize_torch(feats, fit_rows=None, chunk=1 << 20):
    """`standardize` on the device that holds `feats` (float32 [N, F] tensor), in place: column mean / population variance of the fit rows
    accumulated in float64 over row chunks, zero-variance columns only centred, transform in float32 (loader.py:104-112).  The float64
    statistics are the ones sklearn computes; the float32 transform matches its float32-in / float32-out behaviour."""
    import torch
    n, dev = feats.shape[0], feats.device
    rows = None if fit_rows is None else torch.as_tensor(np.asarray(fit_rows), device=dev).long()
    cnt = n if rows is None else int(rows.numel())
    s1 = torch.zeros(feats.shape[1], dtype=torch.float64, device=dev)
    for a in range(0, cnt, chunk):
        blk = feats[a:a + chunk] if rows is None else feats[rows[a:a + chunk]]
        s1 += blk.double().sum(0)
    mean = s1 / max(cnt, 1)
    s2 = torch.zeros_like(s1)
    for a in range(0, cnt, chunk):
        blk = feats[a:a + chunk] if rows is None else feats[rows[a:a + chunk]]
        d = blk.double() - mean
        s2 += (d * d).sum(0)
    scale = torch.sqrt(s2 / max(cnt, 1))
    scale[scale < 10 * np.finfo(np.float64).eps] = 1.0
    m32, s32 = mean.float(), scale.float()
    for a in range(0, n, chunk):
        feats[a:a + chunk] = (feats[a:a + chunk] - m32) / s32
    return feats


def load_data_device(prefix, dataset, config_data, device, printf=lambda txt, style=None: None):
    """f-2, direct-to-HBM half: the files `load_data` reads (loader.py:18-121) go straight to `device` -- the raw adjacency's index arrays
    through pinned staging, the undirected conversion (graph_utils.py:19-45) as ONE device sort/unique instead of the reference's per-row
    Python loop, the StandardScaler on the device -- and come back in the form the device minibatch takes:
        adjs[mode] = (indptr int32, indices int32) device tensors holding uint32 bit patterns (shared between modes when transductive),
        feat_full float32 [N, F], label_full, node_set (host int64 arrays).
    Same decisions as `load_data` (prestored undirected / normalised files win; transductive => one adjacency); node datasets only."""
    import torch
    d = f"{prefix['local']}/{dataset}"
    for f in ("split.npy", "label_full.npy", "feat_full.npy"):
        if not os.path.isfile(f"{d}/{f}"):
            raise FileNotFoundError(f"{d}/{f} is missing")
    role = np.load(f"{d}/split.npy", allow_pickle=True)
    role = role[()] if isinstance(role, np.ndarray) else role
    if not all(k in role for k in (TRAIN, VALID, TEST)):
        raise NotImplementedError("link-prediction splits (edge sets) are not handled by this loader")
    node_set = {k: np.asarray(role[k], dtype=np.int64) for k in (TRAIN, VALID, TEST)}
    if "coalesce" in config_data and not config_data["coalesce"]:
        raise NotImplementedError
    dev = torch.device(device)
    pin = dev.type == "cuda"

    def up(a, dtype):
        t = torch.from_numpy(np.ascontiguousarray(a).astype(dtype, copy=False))
        return (t.pin_memory() if pin else t).to(dev, non_blocking=pin)

    def adj_dev(split_):
        und = config_data["to_undirected"]
        a = _load_adj(prefix["local"], dataset, "undirected" if und else "raw", split_)
        if a is None and und:
            raw = _load_adj(prefix["local"], dataset, "raw", split_)
            if raw is None:
                raise FileNotFoundError(f"{d}/adj_{split_}_raw.np[yz] is missing")
            raw = raw.tocsr()
            ip, ix = to_undirected_csr_torch(up(raw.indptr, np.int64), up(raw.indices, np.int64), device=dev)
        elif a is None:
            raise FileNotFoundError(f"{d}/adj_{split_}_raw.np[yz] is missing")
        else:
            a = a.tocsr()
            ip, ix = up(a.indptr, np.int64), up(a.indices, np.int64).to(torch.int32)
        assert int(ip[-1]) < 2 ** 32 and ip.numel() - 1 < 2 ** 32
        ip32 = torch.where(ip >= 2 ** 31, ip - 2 ** 32, ip).to(torch.int32)        # uint32 bit patterns in int32 storage
        return ip32, ix
    full = adj_dev("full")
    adjs = {VALID: full, TEST: full, TRAIN: full if config_data["transductive"] else adj_dev("train")}
    mode_norm = "all" if config_data["transductive"] else "train"
    if config_data["norm_feat"] and os.path.isfile(f"{d}/feat_full_norm_{mode_norm}.npy"):
        feats = up(np.load(f"{d}/feat_full_norm_{mode_norm}.npy"), np.float32)
    else:
        feats = up(np.load(f"{d}/feat_full.npy"), np.float32)
        if config_data["norm_feat"]:
            standardize_torch(feats, None if config_data["transductive"] else node_set[TRAIN])
    lab = np.load(f"{d}/label_full.npy")
    label_full = up(lab, lab.dtype)
    printf("Done loading training data to the device..")
    return adjs, feats, label_full, node_set
