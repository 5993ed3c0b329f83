"""ctypes front-end of the CPU oracle (oracle/oracle_sampler.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; nothing under shadow_gnn_b200/ does.  The result layout mirrors what the reference's
`GraphSamplerEnsemble._extract_subgraph_return` sees from pybind
(para_graph_sampler/graph_engine/frontend/samplers_ensemble.py:250-265): per-subgraph arrays
indptr / indices / node / edge_index / target / hop / ppr / drnl.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle_sampler.so")

METHOD = {"khop": 0, "ppr": 1, "ppr_st": 2, "nodeIID": 3}
AUG = {"hops": 1, "pprs": 2, "drnls": 4}


class _Cfg(C.Structure):
    _fields_ = [("method", C.c_int), ("num_roots", C.c_int), ("depth", C.c_int), ("budget", C.c_int),
                ("k", C.c_int), ("threshold", C.c_float), ("add_self_edge", C.c_int),
                ("include_target_conn", C.c_int), ("return_target_only", C.c_int), ("aug", C.c_int),
                ("fixed_mode", C.c_int)]


class _Batch(C.Structure):
    _fields_ = [("num_valid", C.c_int)] + \
        [(n, C.POINTER(C.c_int64)) for n in ("node_ptr", "edge_ptr", "indptr_ptr", "target_ptr", "hop_ptr", "drnl_ptr")] + \
        [(n, C.POINTER(C.c_uint32)) for n in ("indptr", "indices", "orig_node", "orig_edge", "target", "hop", "drnl")] + \
        [("ppr", C.POINTER(C.c_float)), ("rand_draws", C.c_int64)]


def build(force=False):
    """Compile the C restatement (and, when /root/reference exists, oracle/_ref)."""
    if force or not os.path.exists(_LIB_PATH) or \
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "oracle_sampler.c")):
        subprocess.check_call(["make", "-s", "-C", _HERE, "port", f"PY={sys.executable}"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_shuffle_targets.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.orc_get_idx_root.restype = C.c_uint32
        L.orc_get_idx_root.argtypes = [C.c_void_p]
        L.orc_reseed.argtypes = [C.c_void_p, C.c_int]
        L.orc_set_ppr.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_sample.restype = C.POINTER(_Batch)
        L.orc_sample.argtypes = [C.c_void_p, C.POINTER(_Cfg), C.c_int]
        L.orc_batch_free.argtypes = [C.POINTER(_Batch)]
        L.orc_ppr_push.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_int64, C.c_int, C.c_float,
                                   C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_srand.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_rand.restype = C.c_uint32
        L.orc_rand.argtypes = [C.c_void_p]
        L.orc_rand_fill.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def glibc_rand_stream(seed, n):
    """First n outputs of rand() after srand(seed)."""
    st = (C.c_uint32 * 36)()
    lib().orc_srand(st, seed)
    out = np.empty(n, np.uint32)
    lib().orc_rand_fill(st, _ptr(out), n)
    return out


def make_cfg(method, num_roots=1, depth=0, budget=0, k=0, threshold=0.0, add_self_edge=False,
             include_target_conn=False, return_target_only=False, aug=(), fixed_mode=False):
    return _Cfg(METHOD[method], int(num_roots), int(depth), int(budget), int(k), float(threshold),
                int(bool(add_self_edge)), int(bool(include_target_conn)), int(bool(return_target_only)),
                sum(AUG[a] for a in aug), int(bool(fixed_mode)))


def cfg_from_cpp_config(cpp_config, aug=(), fixed_mode=False):
    """Translate the reference's str->str sampler config (frontend/samplers_cpp.py:73-80,124-131)."""
    t = lambda key: cpp_config.get(key, "false") in ("true", "True", "1")
    return make_cfg(cpp_config["method"], num_roots=int(cpp_config["num_roots"]),
                    depth=int(cpp_config.get("depth", 0)), budget=int(cpp_config.get("budget", 0)),
                    k=int(cpp_config.get("k", 0)), threshold=float(cpp_config.get("threshold", 0)),
                    add_self_edge=t("add_self_edge"), include_target_conn=t("include_target_conn"),
                    return_target_only=t("return_target_only"), aug=aug, fixed_mode=fixed_mode)


class FlatBatch:
    """Concatenated result arrays of one sampler call + per-subgraph views."""
    FIELDS = ("indptr", "indices", "node", "edge_index", "target", "hop", "ppr", "drnl")

    def __init__(self, **kw):
        self.__dict__.update(kw)

    @property
    def num_subg(self):
        return len(self.node_ptr) - 1

    def subgraph(self, p):
        s = lambda arr, ptr: arr[ptr[p]:ptr[p + 1]]
        return dict(indptr=s(self.indptr, self.indptr_ptr), indices=s(self.indices, self.edge_ptr),
                    node=s(self.node, self.node_ptr), edge_index=s(self.edge_index, self.edge_ptr),
                    target=s(self.target, self.target_ptr), hop=s(self.hop, self.hop_ptr),
                    ppr=s(self.ppr, self.node_ptr) if len(self.ppr) else self.ppr[:0],
                    drnl=s(self.drnl, self.drnl_ptr))

    def subgraphs(self):
        return [self.subgraph(p) for p in range(self.num_subg)]


class OracleSampler:
    """CPU restatement of `ParallelSampler` (backend/ParallelSampler.cpp:707-734)."""

    def __init__(self, indptr, indices, num_sampler_per_batch, num_threads=1, seed=0):
        self.indptr = np.ascontiguousarray(indptr, dtype=np.uint32)
        self.indices = np.ascontiguousarray(indices, dtype=np.uint32)
        self._h = lib().orc_create(_ptr(self.indptr), _ptr(self.indices), self.indptr.size - 1, self.indices.size,
                                   num_sampler_per_batch, num_threads, seed)
        self.num_threads = num_threads

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_destroy(self._h)
            self._h = None

    def num_nodes(self):
        return self.indptr.size - 1

    def num_edges(self):
        return self.indices.size

    def shuffle_targets(self, targets):
        t = np.ascontiguousarray(np.asarray(targets).flatten(), dtype=np.uint32)
        lib().orc_shuffle_targets(self._h, _ptr(t), t.size)

    def get_idx_root(self):
        return lib().orc_get_idx_root(self._h)

    def reseed(self, seed):
        lib().orc_reseed(self._h, seed)

    def set_ppr(self, ptr, neighs, scores):
        ptr = np.ascontiguousarray(ptr, dtype=np.uint64)
        neighs = np.ascontiguousarray(neighs, dtype=np.uint32)
        scores = np.ascontiguousarray(scores, dtype=np.float32)
        lib().orc_set_ppr(self._h, _ptr(ptr), _ptr(neighs), _ptr(scores))

    def preproc_ppr_approximate(self, targets, k, alpha, epsilon):
        """PS.cpp:237-344.  Returns and installs (ptr, neighs, scores) indexed by node id."""
        targets = np.ascontiguousarray(targets, dtype=np.uint32)
        nb, sc, ln = ppr_push(self.indptr, self.indices, targets, k, alpha, epsilon, self.num_threads)
        ptr, fn, fs = ppr_rows_to_csr(self.num_nodes(), targets, nb, sc, ln)
        self.set_ppr(ptr, fn, fs)
        return ptr, fn, fs

    def sample(self, cfg, advance_roots=True):
        bp = lib().orc_sample(self._h, C.byref(cfg), int(advance_roots))
        b = bp.contents
        P = b.num_valid

        def arr(p, n, dt):
            return np.ctypeslib.as_array(p, shape=(max(int(n), 1),))[:int(n)].astype(dt, copy=True)
        ptrs = {n: arr(getattr(b, n), P + 1, np.int64) for n in
                ("node_ptr", "edge_ptr", "indptr_ptr", "target_ptr", "hop_ptr", "drnl_ptr")}
        has_csr = ptrs["indptr_ptr"][P] > 0
        out = FlatBatch(
            indptr=arr(b.indptr, ptrs["indptr_ptr"][P], np.uint32), indices=arr(b.indices, ptrs["edge_ptr"][P], np.uint32),
            node=arr(b.orig_node, ptrs["node_ptr"][P], np.uint32), edge_index=arr(b.orig_edge, ptrs["edge_ptr"][P], np.uint32),
            target=arr(b.target, ptrs["target_ptr"][P], np.uint32), hop=arr(b.hop, ptrs["hop_ptr"][P], np.uint32),
            ppr=arr(b.ppr, ptrs["node_ptr"][P] if has_csr else 0, np.float32), drnl=arr(b.drnl, ptrs["drnl_ptr"][P], np.uint32),
            rand_draws=int(b.rand_draws), **ptrs)
        lib().orc_batch_free(bp)
        return out


def ppr_push(indptr, indices, targets, k, alpha, epsilon, num_threads=1):
    indptr = np.ascontiguousarray(indptr, dtype=np.uint32)
    indices = np.ascontiguousarray(indices, dtype=np.uint32)
    targets = np.ascontiguousarray(targets, dtype=np.uint32)
    T = targets.size
    nb = np.zeros((T, k), np.uint32)
    sc = np.zeros((T, k), np.float32)
    ln = np.zeros(T, np.uint32)
    lib().orc_ppr_push(_ptr(indptr), _ptr(indices), indptr.size - 1, _ptr(targets), T, k, alpha, epsilon,
                       num_threads, _ptr(nb), _ptr(sc), _ptr(ln))
    return nb, sc, ln


def ppr_rows_to_csr(num_nodes, targets, nb, sc, ln):
    """[T,k] padded rows -> (ptr[num_nodes+1], neighs, scores) indexed by node id (top_ppr_* layout)."""
    lens = np.zeros(num_nodes, np.uint64)
    lens[targets] = ln
    ptr = np.zeros(num_nodes + 1, np.uint64)
    np.cumsum(lens, out=ptr[1:])
    tot = int(ptr[-1])
    fn = np.zeros(tot, np.uint32)
    fs = np.zeros(tot, np.float32)
    mask = np.arange(nb.shape[1])[None, :] < ln[:, None]
    order = np.argsort(targets, kind="stable")
    # a node that appears twice in `targets` keeps one (identical) row
    _, first = np.unique(targets[order], return_index=True)
    sel = order[first]
    fn[:] = nb[sel][mask[sel]]
    fs[:] = sc[sel][mask[sel]]
    return ptr, fn, fs


def cat_to_block_diagonal(subgs):
    """numpy restatement of Subgraph.cat_to_block_diagonal (frontend/graph.py:280-320)."""
    n_off = np.concatenate([[0], np.cumsum([s["node"].size for s in subgs])[:-1]]).astype(np.int64)
    e_off = np.concatenate([[0], np.cumsum([s["edge_index"].size for s in subgs])[:-1]]).astype(np.int64)
    indptr = np.concatenate([(s["indptr"].astype(np.int64) if i == 0 else s["indptr"][1:].astype(np.int64)) + e_off[i]
                             for i, s in enumerate(subgs)])
    return dict(
        indptr=indptr,
        indices=np.concatenate([s["indices"].astype(np.int64) + n_off[i] for i, s in enumerate(subgs)]),
        node=np.concatenate([s["node"] for s in subgs]),
        edge_index=np.concatenate([s["edge_index"] for s in subgs]),
        target=np.concatenate([s["target"].astype(np.int64) + n_off[i] for i, s in enumerate(subgs)]),
        size_subg=np.array([s["node"].size for s in subgs], np.int64))


# ------------------------------------------------------------------------------------------------
# the compiled, unmodified reference (oracle/_ref) -- only importable where it was built/shipped
# ------------------------------------------------------------------------------------------------
def load_ref():
    """Import oracle/_ref/ParallelSampler*.so (the reference's pybind module, PS.cpp:707-746)."""
    import glob
    import importlib.util
    so = glob.glob(os.path.join(_HERE, "_ref", "ParallelSampler*.so"))
    if not so:
        return None
    spec = importlib.util.spec_from_file_location("ParallelSampler", so[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def ref_subgraphs(vec, aug=()):
    """SubgraphStructVec -> list of dicts of numpy arrays (samplers_ensemble.py:254-265)."""
    clip = vec.get_num_valid_subg()
    names = dict(indptr="indptr", indices="indices", node="node", edge_index="edge_index", target="target",
                 hop="hop", ppr="ppr", drnl="drnl")
    cols = {n: getattr(vec, f"get_subgraph_{g}")()[:clip] for n, g in names.items()}
    out = []
    for p in range(clip):
        d = {}
        for n in names:
            dt = np.float32 if n == "ppr" else np.uint32
            d[n] = np.asarray(cols[n][p], dtype=np.int64).astype(dt) if n != "ppr" else np.asarray(cols[n][p], dtype=dt)
        out.append(d)
    return out
