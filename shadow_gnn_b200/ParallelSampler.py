"""Drop-in for the reference's pybind module `ParallelSampler`
(para_graph_sampler/graph_engine/backend/ParallelSampler.cpp:707-746), backed by the CUDA C ABI.

    import shadow_gnn_b200.ParallelSampler as cpp_para_sampler     # instead of `import ParallelSampler`

Same class names, ctor arity (13 positional arguments), method names, argument meaning and return shapes as the
pybind module, so `GraphSamplerEnsemble` (frontend/samplers_ensemble.py:100-113,179,192,254-265,296,300-301) and
`PPRSamplingCpp.preproc` (frontend/samplers_cpp.py:184) run on it unmodified.  On top of that,
`sample_to_device()` exposes the batch where it already lives -- in HBM, as one block-diagonal CSR.

Error behaviour mirrors what pybind surfaces from the C++ code: a missing config key is an IndexError
(`unordered_map::at` -> std::out_of_range), an unparsable number a ValueError (std::stoi/stod ->
std::invalid_argument).  The one deliberate difference: a config without "method" raises ValueError
instead of killing the process with exit(1) (PS.cpp:676-679).
"""
import ctypes as C
import re

import numpy as np

from . import _lib
from ._lib import lib, check

_INT_RE = re.compile(r"^\s*[+-]?\d+")
_FLT_RE = re.compile(r"^\s*[+-]?(?:(?:\d+\.?\d*|\.\d+)(?:[eE][+-]?\d+)?|inf(?:inity)?|nan)", re.I)


def _at(cfg, key):
    if key not in cfg:
        raise IndexError(f"sampler config has no key '{key}'")       # std::out_of_range via pybind
    return cfg[key]


def _stoi(s):
    m = _INT_RE.match(str(s))
    if not m:
        raise ValueError(f"stoi: cannot parse '{s}'")
    return int(m.group(0))


def _stod(s):
    m = _FLT_RE.match(str(s))
    if not m:
        raise ValueError(f"stod: cannot parse '{s}'")
    return float(m.group(0))


def _bool(cfg, key, default=False):           # _extract_bool_config (PS.cpp:483-496)
    if key in cfg:
        return cfg[key] in ("true", "True", "1")
    return default


def parse_cfg(cfg, aug=(), fixed_mode=False, rng_mode=_lib.RNG_GLIBC):
    """str->str sampler config (frontend/samplers_cpp.py:50-56,73-80,124-131) -> shadow_sampler_cfg."""
    if "method" not in cfg:
        raise ValueError("[parallel sampler]: need to have the 'method' key in the config")
    method = cfg["method"]
    c = _lib.SamplerCfg()
    c.num_roots = _stoi(_at(cfg, "num_roots"))
    c.return_target_only = int(_bool(cfg, "return_target_only"))
    if method not in _lib.METHOD:
        # the reference silently returns empty subgraphs for an unknown method (PS.cpp:685-693); be explicit
        raise ValueError(f"unknown sampler method '{method}'")
    c.method = _lib.METHOD[method]
    if not c.return_target_only:
        if method == "khop":
            c.depth, c.budget = _stoi(_at(cfg, "depth")), _stoi(_at(cfg, "budget"))
        elif method in ("ppr", "ppr_st"):
            c.k = _stoi(_at(cfg, "k"))
            c.threshold = _stod(_at(cfg, "threshold"))
    c.add_self_edge = int(_bool(cfg, "add_self_edge"))
    c.include_target_conn = int(_bool(cfg, "include_target_conn"))
    c.aug = sum(_lib.AUG[a] for a in aug if a in _lib.AUG)
    c.fixed_mode = int(bool(fixed_mode))
    c.rng_mode = int(rng_mode)
    return c


class _DevArray:
    """Zero-copy view of a library-owned device buffer (`__cuda_array_interface__`)."""

    def __init__(self, ptr, count, typestr, owner):
        self._owner = owner
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": typestr, "data": (int(ptr or 0), False),
                                         "version": 2, "strides": None}


class DeviceBatch:
    """One ensemble branch of the latest call, resident in HBM as a block-diagonal batch (layouts: shadow_b200.h).
    Tensors are zero-copy views fetched on first access, valid until `num_ring` further sampler calls.
    RAW edge layout: row_span [N,2], edge_span [P,2], indices_raw, orig_edge_raw (what the kernel wrote);
    CANONICAL CSR: edge_ptr, rowptr, indices, orig_edge (one extra streaming kernel on first access)."""
    _FIELDS = {"node_ptr": (_lib.F_NODE_PTR, "<i4"), "edge_ptr": (_lib.F_EDGE_PTR, "<i4"), "rowptr": (_lib.F_ROWPTR, "<i4"),
               "indices": (_lib.F_INDICES, "<i4"), "orig_node": (_lib.F_ORIG_NODE, "<i4"), "orig_edge": (_lib.F_ORIG_EDGE, "<i4"),
               "target": (_lib.F_TARGET, "<i4"), "ppr": (_lib.F_PPR, "<f4"), "hop": (_lib.F_HOP, "<i4"),
               "drnl": (_lib.F_DRNL, "<i4"), "num_target": (_lib.F_NUM_TARGET, "<i4"),
               "row_span": (_lib.F_ROW_SPAN, "<i4"), "edge_span": (_lib.F_EDGE_SPAN, "<i4"),
               "indices_raw": (_lib.F_INDICES_RAW, "<i4"), "orig_edge_raw": (_lib.F_ORIG_EDGE_RAW, "<i4")}

    def __init__(self, sampler, branch):
        info = _lib.BatchInfo()
        check(lib.shadow_sampler_batch_info(sampler._h, branch, C.byref(info)))
        self._sampler, self._branch = sampler, branch
        self.num_subg, self.num_roots = info.num_subg, info.num_roots
        self.total_nodes, self.total_edges = info.total_nodes, info.total_edges
        self.has_csr = bool(info.has_csr)

    def __getattr__(self, name):
        if name.startswith("_") or name not in self._FIELDS:
            raise AttributeError(name)
        import torch
        fid, ts = self._FIELDS[name]
        sampler = self._sampler
        dev = torch.device("cuda", sampler.device)
        p, n = C.c_void_p(), C.c_int64()
        check(lib.shadow_sampler_batch_field_dev(sampler._h, self._branch, fid, C.byref(p), C.byref(n)))
        if n.value == 0:
            t = torch.empty(0, dtype=torch.float32 if ts == "<f4" else torch.int32, device=dev)
        else:
            t = torch.as_tensor(_DevArray(p.value, n.value, ts, sampler), device=dev)
        if name in ("row_span", "edge_span"):
            t = t.view(-1, 2)
        setattr(self, name, t)
        return t


class SubgraphStructVec:
    """Host view of one ensemble branch: same getters as the pybind class (PS.cpp:735-745; G.h:59-97)."""

    def __init__(self, sampler, branch):
        info = _lib.BatchInfo()
        check(lib.shadow_sampler_batch_info(sampler._h, branch, C.byref(info)))
        self._info = info
        self._num_per_batch = sampler.num_sampler_per_batch

        def fetch(fid, dtype=np.uint32):
            p, n = C.c_void_p(), C.c_int64()
            check(lib.shadow_sampler_batch_field_dev(sampler._h, branch, fid, C.byref(p), C.byref(n)))
            a = np.empty(n.value, dtype)
            check(lib.shadow_sampler_batch_field_host(sampler._h, branch, fid, a.ctypes.data_as(C.c_void_p), n.value))
            return a
        P = info.num_subg
        self._np = {}
        node = fetch(_lib.F_ORIG_NODE)
        if info.has_csr:
            node_ptr = fetch(_lib.F_NODE_PTR, np.int32).astype(np.int64)
            edge_ptr = fetch(_lib.F_EDGE_PTR, np.int32).astype(np.int64)
            rowptr = fetch(_lib.F_ROWPTR, np.int32).astype(np.int64)
            indices = fetch(_lib.F_INDICES, np.int32).astype(np.int64)
            target = fetch(_lib.F_TARGET, np.int32).astype(np.int64).reshape(P, info.num_roots)
            ntarget = fetch(_lib.F_NUM_TARGET, np.int32)
            eidx = fetch(_lib.F_ORIG_EDGE)
            ppr = fetch(_lib.F_PPR, np.float32)
            hop = fetch(_lib.F_HOP) if info.has_hop else None
            drnl = fetch(_lib.F_DRNL) if info.has_drnl else None
            sp = lambda a, ptr, p: a[ptr[p]:ptr[p + 1]]
            self._np["indptr"] = [(rowptr[node_ptr[p]:node_ptr[p + 1] + 1] - edge_ptr[p]).astype(np.uint32) for p in range(P)]
            self._np["indices"] = [(sp(indices, edge_ptr, p) - node_ptr[p]).astype(np.uint32) for p in range(P)]
            self._np["data"] = [np.ones(edge_ptr[p + 1] - edge_ptr[p], np.float32) for p in range(P)]     # PS.cpp:411,423
            self._np["node"] = [sp(node, node_ptr, p) for p in range(P)]
            self._np["edge_index"] = [sp(eidx, edge_ptr, p) for p in range(P)]
            self._np["target"] = [(target[p, :ntarget[p]] - node_ptr[p]).astype(np.uint32) for p in range(P)]
            self._np["ppr"] = [sp(ppr, node_ptr, p) for p in range(P)]
            empty = np.zeros(0, np.uint32)
            self._np["hop"] = [sp(hop, node_ptr, p) for p in range(P)] if hop is not None else [empty] * P
            self._np["drnl"] = [sp(drnl, node_ptr, p) for p in range(P)] if drnl is not None else [empty] * P
        else:                                     # dummy_sampler: origNodeID = roots, everything else empty
            nr = info.num_roots
            self._np["node"] = [node[p * nr:(p + 1) * nr] for p in range(P)]
            for k in ("indptr", "indices", "edge_index", "target", "hop", "drnl"):
                self._np[k] = [np.zeros(0, np.uint32)] * P
            self._np["data"] = [np.zeros(0, np.float32)] * P
            self._np["ppr"] = [np.zeros(0, np.float32)] * P

    def get_num_valid_subg(self):
        return self._info.num_subg

    def numpy(self, name):
        """per-subgraph numpy arrays (no Python-list round trip)"""
        return self._np[name]

    def _lists(self, name):
        # fixed capacity num_sampler_per_batch like the reference's *_vec (G.h:62-72); callers clip with get_num_valid_subg
        out = [a.tolist() for a in self._np[name]]
        out.extend([] for _ in range(self._num_per_batch - len(out)))
        return out

    def get_subgraph_indptr(self): return self._lists("indptr")
    def get_subgraph_indices(self): return self._lists("indices")
    def get_subgraph_data(self): return self._lists("data")
    def get_subgraph_node(self): return self._lists("node")
    def get_subgraph_edge_index(self): return self._lists("edge_index")
    def get_subgraph_target(self): return self._lists("target")
    def get_subgraph_hop(self): return self._lists("hop")
    def get_subgraph_ppr(self): return self._lists("ppr")
    def get_subgraph_drnl(self): return self._lists("drnl")


class ParallelSampler:
    """class ParallelSampler (PS.h:25-158).  Positional signature identical to the pybind ctor (PS.cpp:710-724):

        ParallelSampler(indptr, indices, data, num_sampler_per_batch, max_num_threads, fix_target,
                        sequential_traversal, edge_reweighted, num_subgraphs_ensemble,
                        path_indptr, path_indices, path_data, seed)

    `data`, `edge_reweighted`, `path_data` are accepted and ignored exactly as the reference ignores them (PS.h:48);
    `max_num_threads` has no meaning on the GPU.  Keyword-only extras select the device-side behaviour.
    """

    def __init__(self, indptr, indices, data, num_sampler_per_batch, max_num_threads, fix_target,
                 sequential_traversal, edge_reweighted=(), num_subgraphs_ensemble=1, path_indptr="", path_indices="",
                 path_data="", seed=-1, *, device=None, strict_reference_compat=True, rng="glibc", num_ring=2):
        import torch
        if not torch.cuda.is_available():
            raise _lib.ShadowError("shadow_gnn_b200.ParallelSampler needs a CUDA device (no CPU fallback)")
        if not fix_target or not sequential_traversal:
            # forced on by the only caller (shaDow/minibatch.py:373-375; asserted samplers_ensemble.py:93)
            raise ValueError("only fix_target=True, sequential_traversal=True are supported")
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.num_sampler_per_batch = int(num_sampler_per_batch)
        self.fixed_mode = not strict_reference_compat
        self.rng_mode = {"glibc": _lib.RNG_GLIBC, "philox": _lib.RNG_PHILOX}[rng]
        self._sequential = bool(sequential_traversal)
        h = C.c_void_p()
        ip = np.ascontiguousarray(indptr, dtype=np.uint32) if len(indptr) else None
        ix = np.ascontiguousarray(indices, dtype=np.uint32) if len(indices) else None
        check(lib.shadow_sampler_create(
            ip.ctypes.data_as(C.c_void_p) if ip is not None else None,
            ix.ctypes.data_as(C.c_void_p) if ix is not None else None,
            (ip.size - 1) if ip is not None else 0, ix.size if ix is not None else 0,
            str(path_indptr).encode(), str(path_indices).encode(), self.num_sampler_per_batch,
            int(num_subgraphs_ensemble), int(seed), self.device, int(num_ring), C.byref(h)))
        self._h = h
        self.num_ensemble = int(num_subgraphs_ensemble)
        self._user_stream = False

    @classmethod
    def from_device_csr(cls, indptr, indices, num_sampler_per_batch, num_subgraphs_ensemble=1, seed=-1, *,
                        strict_reference_compat=True, rng="glibc", num_ring=2):
        """Sampler over a CSR that already lives in HBM (int32/uint32 CUDA tensors, borrowed -- the sampler keeps them alive)."""
        self = cls.__new__(cls)
        assert indptr.is_cuda and indices.is_cuda and indptr.element_size() == 4 and indices.element_size() == 4
        assert indptr.is_contiguous() and indices.is_contiguous()
        self.device = indptr.device.index
        self.num_sampler_per_batch = int(num_sampler_per_batch)
        self.fixed_mode = not strict_reference_compat
        self.rng_mode = {"glibc": _lib.RNG_GLIBC, "philox": _lib.RNG_PHILOX}[rng]
        self._sequential = True
        self._graph_keepalive = (indptr, indices)
        h = C.c_void_p()
        check(lib.shadow_sampler_create_dev(C.c_void_p(indptr.data_ptr()), C.c_void_p(indices.data_ptr()), indptr.numel() - 1,
                                            indices.numel(), self.num_sampler_per_batch, int(num_subgraphs_ensemble), int(seed),
                                            self.device, int(num_ring), C.byref(h)))
        self._h = h
        self.num_ensemble = int(num_subgraphs_ensemble)
        self._user_stream = False
        return self

    def __del__(self, _destroy=lib.shadow_sampler_destroy):
        h = getattr(self, "_h", None)
        if h:
            _destroy(h)
            self._h = None

    # ---- PS.cpp:726-734 ----
    def num_nodes(self): return lib.shadow_sampler_num_nodes(self._h)
    def num_edges(self): return lib.shadow_sampler_num_edges(self._h)
    def num_nodes_target(self): return lib.shadow_sampler_num_nodes_target(self._h)
    def get_idx_root(self): return lib.shadow_sampler_get_idx_root(self._h)
    def is_seq_root_traversal(self): return self._sequential

    def shuffle_targets(self, targets_pre_shuffled):
        t = np.ascontiguousarray(np.asarray(targets_pre_shuffled).flatten(), dtype=np.uint32)
        if t.size == 0:
            raise ValueError("shuffle_targets([]) (in-place std::random_shuffle, PS.cpp:38) is not on the supported path")
        check(lib.shadow_sampler_shuffle_targets(self._h, t.ctypes.data_as(C.c_void_p), t.size))

    def preproc_ppr_approximate(self, preproc_target, k, alpha, epsilon, fname_neighs, fname_scores):
        t = np.ascontiguousarray(np.asarray(preproc_target).flatten(), dtype=np.uint32)
        check(lib.shadow_sampler_preproc_ppr_approximate(self._h, t.ctypes.data_as(C.c_void_p), t.size, int(k), float(alpha),
                                                         float(epsilon), str(fname_neighs).encode(), str(fname_scores).encode()))

    def drop_full_graph_info(self):
        check(lib.shadow_sampler_drop_full_graph_info(self._h))

    def parallel_sampler_ensemble(self, configs_samplers, configs_aug):
        self._launch(configs_samplers, configs_aug)
        return [SubgraphStructVec(self, b) for b in range(len(configs_samplers))]

    # ---- device-side extras ----
    def last_redo_count(self):
        """subgraphs of the last validated call that the one-warp PPR fast path handed to the generic kernel (diagnostics)"""
        return int(lib.shadow_sampler_last_redo_count(self._h))

    def last_sym(self):
        """True when the last call ran the symmetric-graph (upper-triangle scan + mirror) variant of the PPR fast path (diagnostics)"""
        return bool(lib.shadow_sampler_last_sym(self._h) == 1)

    def last_sequence_ms(self):
        """GPU time of all work of the last fast-path call: state reset, count, scan, main kernel, redo launch"""
        return float(lib.shadow_sampler_last_sequence_ms(self._h))

    def last_kernel_ms(self):
        """duration of the fast path's main kernel alone in the last call (CUDA events inside the library; waits for the kernel)"""
        return float(lib.shadow_sampler_last_kernel_ms(self._h))

    def set_stream(self, cuda_stream):
        """pin the sampler to one stream (default: whatever torch's current stream is at each call)"""
        self._user_stream = True
        check(lib.shadow_sampler_set_stream(self._h, C.c_void_p(int(cuda_stream))))

    def set_num_sampler_per_batch(self, n):
        check(lib.shadow_sampler_set_num_per_batch(self._h, int(n)))
        self.num_sampler_per_batch = int(n)

    def reseed(self, seed):
        check(lib.shadow_sampler_reseed(self._h, int(seed)))

    def set_ppr_tables(self, ptr, neighs, scores):
        """install top_ppr_neighs/top_ppr_scores (PS.h:145-146) given as CSR by node id"""
        ptr = np.ascontiguousarray(ptr, dtype=np.uint64)
        neighs = np.ascontiguousarray(neighs, dtype=np.uint32)
        scores = np.ascontiguousarray(scores, dtype=np.float32)
        assert ptr.size == self.num_nodes() + 1 and neighs.size == scores.size == int(ptr[-1])
        check(lib.shadow_sampler_set_ppr_tables(self._h, ptr.ctypes.data_as(C.c_void_p), neighs.ctypes.data_as(C.c_void_p),
                                                scores.ctypes.data_as(C.c_void_p)))

    def get_ppr_row(self, v, cap=4096):
        nb, sc, ln = np.empty(cap, np.uint32), np.empty(cap, np.float32), C.c_uint32()
        check(lib.shadow_sampler_get_ppr_row(self._h, int(v), cap, nb.ctypes.data_as(C.c_void_p), sc.ctypes.data_as(C.c_void_p),
                                             C.byref(ln)))
        return nb[:min(ln.value, cap)], sc[:min(ln.value, cap)]

    def shuffle_targets_device(self, targets_i32):
        """install a target list that already lives in HBM (int32/uint32 CUDA tensor); the caller keeps it alive"""
        self._targets_keepalive = targets_i32
        check(lib.shadow_sampler_shuffle_targets_dev(self._h, C.c_void_p(targets_i32.data_ptr()), targets_i32.numel()))

    def _launch(self, configs_samplers, configs_aug):
        # the sampler works on torch's CURRENT stream, so its kernels (and the lazy canonicalisation launched on first access of rowptr /
        # indices) are ordered with the consumers (gather_rows, ops.*) even under torch.cuda.stream(...); a ring slot is only rewritten
        # `num_ring` calls later on the same stream
        import torch
        if not self._user_stream:
            lib.shadow_sampler_set_stream(self._h, C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        n = len(configs_samplers)
        arr = (_lib.SamplerCfg * n)()
        for i, (cfg, aug) in enumerate(zip(configs_samplers, configs_aug)):
            arr[i] = parse_cfg(cfg, aug, self.fixed_mode, self.rng_mode)
        check(lib.shadow_sampler_sample(self._h, arr, n))

    def sample_to_device(self, configs_samplers, configs_aug):
        """parallel_sampler_ensemble without leaving HBM: one DeviceBatch per ensemble branch."""
        self._launch(configs_samplers, configs_aug)
        return [DeviceBatch(self, b) for b in range(len(configs_samplers))]


def gather_rows(feat, ids, out=None):
    """feat_full[node] on device (shaDow/minibatch.py:469) through shadow_gather_rows_f32."""
    import torch
    assert feat.is_cuda and feat.dtype == torch.float32 and feat.is_contiguous() and feat.dim() == 2
    assert ids.is_cuda and ids.dtype in (torch.int32, torch.uint32) and ids.is_contiguous()
    n = ids.numel()
    if out is None:
        out = torch.empty((n, feat.shape[1]), dtype=torch.float32, device=feat.device)
    check(lib.shadow_gather_rows_f32(C.c_void_p(feat.data_ptr()), feat.shape[0], feat.shape[1], C.c_void_p(ids.data_ptr()), n,
                                     C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream(feat.device).cuda_stream)))
    return out
