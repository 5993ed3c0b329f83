"""Device-resident minibatch extractor -- drop-in for the reference's `shaDow.minibatch.MinibatchShallowExtractor`
(shaDow/minibatch.py:143-495): same constructor arguments, same epoch protocol (`epoch_start_reset`, `shuffle_entity`,
`is_end_epoch`, `one_batch`, `epoch_end_reset`, `disable_cache`, `drop_full_graph_info`, `get_aug_dim`) and the same
`OneBatchSubgraph` fields.

What moved: the sampler call, the pool, the collation (`Subgraph.cat_to_block_diagonal`, frontend/graph.py:280-320), the
feature gather (`feat_full[node]`, minibatch.py:469) and the hop one-hot (graph.py:134-147) all happen in HBM.  One sampler
call produces a super-batch that already IS the block-diagonal CSR; a training batch is a zero-copy slice of it (row range +
column offset).  The reference's per-root Python cache of PPR subgraphs (minibatch.py:69-91) is not needed: re-sampling a
super-batch costs less than looking the subgraphs up.
"""
from collections import deque
from copy import deepcopy
from dataclasses import dataclass
from typing import Any, Dict, List, Optional

import os

import numpy as np
import torch

from . import ParallelSampler as PS
from .ops import DeviceCSR

TRAIN, VALID, TEST = 0, 1, 2
MODE2STR = {TRAIN: "train", VALID: "valid", TEST: "test"}
STR2MODE = {v: k for k, v in MODE2STR.items()}
REUSABLE_SAMPLER = {"ppr", "nodeIID"}            # CONFIG_TEMPLATE.yml:16-17


@dataclass
class OneBatchSubgraph:
    """fields of the reference's dataclass (minibatch.py:94-140); adj_ens holds DeviceCSR handles, the rest CUDA tensors"""
    adj_ens: List[Any]
    feat_ens: List[torch.Tensor]
    label: torch.Tensor
    size_subg_ens: Optional[torch.Tensor]
    target_ens: List[torch.Tensor]
    feat_aug_ens: Optional[List[Dict[str, Any]]]
    idx_raw: Optional[List[Any]] = None

    @property
    def num_ens(self):
        return len(self.adj_ens)

    @property
    def batch_size(self):
        return self.target_ens[0].numel()

    def pop_idx_raw(self):
        ret, self.idx_raw = self.idx_raw, None
        return ret

    def to_dict(self, keys=None):
        keys = self.__dataclass_fields__ if keys is None else keys
        return {k: getattr(self, k) for k in keys}


class _SuperBatch:
    """one sampler call of one ensemble branch, kept in HBM; `cursor` = next unconsumed subgraph"""

    def __init__(self, dev_batch, feat, aug, num_roots):
        self.b = dev_batch
        self.node_ptr_host = dev_batch.node_ptr.cpu().numpy().astype(np.int64)
        if dev_batch.has_csr:
            dev_batch.row_span, dev_batch.indices_raw, dev_batch.target, dev_batch.orig_node      # views fetched while the sampler still points at this call
        self.feat, self.aug, self.num_roots = feat, aug, num_roots
        self.val = torch.empty(max(dev_batch.total_edges, 1), dtype=torch.float32, device=feat.device)
        self.cursor = 0

    @property
    def remaining(self):
        return self.b.num_subg - self.cursor

    def take(self, bs):
        a, b = self.cursor, self.cursor + bs
        lo, hi = int(self.node_ptr_host[a]), int(self.node_ptr_host[b])
        self.cursor = b
        adj = DeviceCSR(self.b.row_span[lo:hi], self.b.indices_raw, lo, self.val)
        R = self.num_roots
        target = self.b.target[a * R:b * R].long() - lo
        sizes = torch.as_tensor(np.diff(self.node_ptr_host[a:b + 1]), device=self.feat.device)
        aug = {k: v[lo:hi] for k, v in self.aug.items()}
        return adj, self.feat[lo:hi], target, sizes, aug, self.b.orig_node[lo:hi]

    def prepare_canonical(self):
        if not hasattr(self, "edge_ptr_host"):
            self.edge_ptr_host = self.b.edge_ptr.cpu().numpy().astype(np.int64)        # runs the canonicalisation kernel once per super-batch
            self.b.rowptr, self.b.indices                                              # views fetched while the sampler still points at this call

    def take_canonical(self, bs):
        """the same batch as contiguous slices of the CANONICAL CSR of the super-batch (for the CUDA-graph trainer, which copies
        them into static buffers): returns (rowptr slice [n+1], indices slice [e], first row, first edge, feat slice, target slice)"""
        self.prepare_canonical()
        a, b = self.cursor, self.cursor + bs
        lo, hi = int(self.node_ptr_host[a]), int(self.node_ptr_host[b])
        e0, e1 = int(self.edge_ptr_host[a]), int(self.edge_ptr_host[b])
        self.cursor = b
        R = self.num_roots
        return self.b.rowptr[lo:hi + 1], self.b.indices[e0:e1], lo, e0, self.feat[lo:hi], self.b.target[a * R:b * R]


def negative_sampling(pos_edge, num_nodes, num_neg, device, max_rounds=8):
    """torch_geometric.utils.negative_sampling as the reference calls it (minibatch.py:289-293), on the device: `num_neg` distinct node pairs
    that are not in add_self_loops(to_undirected(pos_edge)).  Pairs are drawn uniformly from the N x N grid and rejected against the sorted
    key list of the excluded pairs; fewer than `num_neg` come back only if the graph is (nearly) complete.  (The random stream is torch's,
    not PyG's: the pairs differ from the reference's draw, their distribution and the exclusion rule do not.)"""
    pos = torch.as_tensor(np.asarray(pos_edge).astype(np.int64), device=device).reshape(-1, 2)
    N = int(num_nodes)
    loops = torch.arange(N, device=device, dtype=torch.int64)
    excluded = torch.unique(torch.cat([pos[:, 0] * N + pos[:, 1], pos[:, 1] * N + pos[:, 0], loops * N + loops]))
    got = torch.empty(0, dtype=torch.int64, device=device)
    for _ in range(max_rounds):
        need = num_neg - got.numel()
        if need <= 0:
            break
        cand = torch.randint(0, N * N, (int(need * 1.2) + 16,), device=device, dtype=torch.int64)
        idx = torch.searchsorted(excluded, cand).clamp_(max=excluded.numel() - 1)
        cand = cand[excluded[idx] != cand]
        got = torch.cat([got, cand])
        uniq, first = np.unique(got.cpu().numpy(), return_index=True)             # distinct pairs, draw order kept
        got = got[torch.as_tensor(np.sort(first), device=device)]
    got = got[:num_neg]
    return torch.stack([torch.div(got, N, rounding_mode="floor"), got % N], 1).cpu().numpy()


def hop2onehot(hop_i32, dim):
    """EntityEncoding.hop2onehot_vec (frontend/graph.py:134-147): column 0 = unreachable (0xFFFFFFFF, or >= 255), column h+1 = hop h
    for h <= dim-2; larger finite hops give an all-zero row"""
    hop = hop_i32.long() & 0xFFFFFFFF
    out = torch.zeros((hop.numel(), dim), dtype=torch.get_default_dtype(), device=hop.device)
    ok = hop <= dim - 2
    out[ok.nonzero(as_tuple=True)[0], hop[ok] + 1] = 1
    out[(hop >= 255).nonzero(as_tuple=True)[0], 0] = 1
    return out


def ppr2onehot(ppr, dim):
    """EntityEncoding.ppr2onehot_vec (frontend/graph.py:149-158): column i <=> 0.25**(i+1) <= ppr <= 0.25**i (last bin down to 0)"""
    edges = [0.25 ** i for i in range(dim)] + [0.0]
    out = torch.zeros((ppr.numel(), dim), dtype=torch.get_default_dtype(), device=ppr.device)
    for i in range(dim):
        out[:, i] = ((ppr <= edges[i]) & (ppr >= edges[i + 1])).to(out.dtype)
    return out


def drnl2onehot(drnl_i32, dim):
    """EntityEncoding.drnl2onehot_vec (frontend/graph.py:160-172): labels >= 255 or > dim-1 collapse to column 0"""
    d = drnl_i32.long() & 0xFFFFFFFF
    d = torch.where((d >= 255) | (d > dim - 1), torch.zeros_like(d), d)
    return torch.nn.functional.one_hot(d, dim).to(torch.get_default_dtype())


class MinibatchShallowExtractor:
    FULL, SUBG = 0, 1
    NUM_RING = 2            # result buffers of the sampler: a super-batch stays valid while the next one is sampled

    def __init__(self, name_data, dir_data, adjs, entity_set, sampler_config_ensemble, aug_feats, percent_per_epoch, feat_full, label_full,
                 dim_feat_raw: int, is_transductive: bool, parallelism: int, full_tensor_on_gpu: bool = True, bin_adj_files=None,
                 nocache_modes: set = frozenset(), optm_level="high", seed_cpp=-1, metrics_profile=None, *, device=None,
                 num_subg_per_batch=500, strict_reference_compat=True, rng="glibc"):
        if not torch.cuda.is_available():
            raise RuntimeError("shadow_gnn_b200.minibatch needs a CUDA device (no CPU fallback)")
        self.dev_torch = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.name_data, self.dir_data, self.aug_feats, self.optm_level = name_data, dir_data, aug_feats, optm_level
        self.batch_num = -1
        self.batch_size = {TRAIN: 0, VALID: 0, TEST: 0}
        self.raw_entity_set = entity_set
        if isinstance(entity_set[TRAIN], dict):            # link prediction: {"pos": [E,2] (, "neg": [E',2])} per mode (minibatch.py:183-189)
            for v in entity_set.values():
                assert set(v.keys()).issubset({"pos", "neg"}) and "pos" in v
            assert label_full is None, "link prediction takes its labels from the edge sets"
            self.prediction_task = "link"
        else:
            self.prediction_task = "node"
        self.num_roots = 2 if self.prediction_task == "link" else 1          # size_root (minibatch.py:373): both end points root one subgraph
        self.entity_epoch = {m: None for m in (TRAIN, VALID, TEST)}
        self.label_epoch = {m: None for m in (TRAIN, VALID, TEST)}
        self.is_transductive = is_transductive
        self.adj = adjs
        assert isinstance(feat_full, torch.Tensor) and (label_full is None or isinstance(label_full, torch.Tensor))
        self.feat_full = feat_full.to(self.dev_torch).float().contiguous()       # the path keeps the full tensors in HBM
        self.label_full = label_full.to(self.dev_torch) if label_full is not None else None
        self.dim_feat_raw = dim_feat_raw
        self.idx_entity_evaluated = {VALID: 0, TEST: 0, TRAIN: 0}
        self.end_epoch = {VALID: False, TEST: False, TRAIN: False}
        self.percent_per_epoch = {TRAIN: 1.0, VALID: 1.0, TEST: 1.0}
        for k, v in (percent_per_epoch or {}).items():
            self.percent_per_epoch[STR2MODE[k]] = float(v)
        self.nocache_modes = set(nocache_modes)
        self.graph_sampler = {TRAIN: None, VALID: None, TEST: None}
        self.sampler_cfgs = {}
        self._cfg_ensemble = deepcopy(sampler_config_ensemble)
        self.num_ensemble = 0
        for sc in self._cfg_ensemble["configs"]:
            lens = [len(v) for k, v in sc.items() if k != "method"]
            assert len(lens) == 0 or max(lens) == min(lens)
            self.num_ensemble += lens[0] if lens else 1
        assert "full" not in [c["method"] for c in self._cfg_ensemble["configs"]], "FULL (no sampling) mode is a preprocessing-only path of the reference"
        self.mode_sample = self.SUBG
        self.seed_cpp, self.bin_adj_files = seed_cpp, bin_adj_files
        self._per_call, self._compat, self._rng = int(num_subg_per_batch), strict_reference_compat, rng
        self.pool = {m: [deque() for _ in range(self.num_ensemble)] for m in (TRAIN, VALID, TEST)}
        self.record_subgraphs = {}
        self.is_stochastic_sampler = {}
        self.dtype = torch.get_default_dtype()
        self.dim_1hot_hop, self.dim_1hot_ppr, self.dim_1hot_drnl = 5 + 2, 1, 25 + 1        # minibatch.py:246-248
        self.profiler = None
        self.num_sampler_calls = 0
        self.prefetch_canonical = False      # GraphedTrainer: build the canonical CSR (and its host offsets) inside the sampler call

    # ------------------------------------------------------------------ epoch protocol (minibatch.py:252-343)
    def _get_cur_batch_size(self, mode):
        self.end_epoch[mode] = False
        return min(self.entity_epoch[mode].shape[0] - self.idx_entity_evaluated[mode], self.batch_size[mode])

    def _update_batch_stat(self, mode, batch_size):
        self.batch_num += 1
        self.idx_entity_evaluated[mode] += batch_size
        if self.idx_entity_evaluated[mode] >= self.entity_epoch[mode].shape[0]:
            assert self.idx_entity_evaluated[mode] == self.entity_epoch[mode].shape[0]
            self.idx_entity_evaluated[mode] = 0
            self.end_epoch[mode] = True
            assert all(sum(sb.remaining for sb in q) == 0 for q in self.pool[mode])

    def shuffle_entity(self, mode):
        if self.prediction_task == "link":
            return self._shuffle_edges(mode)
        perm = np.random.permutation(self.raw_entity_set[mode].size)
        if self.percent_per_epoch[mode] < 1.0:
            perm = perm[:int(np.ceil(self.percent_per_epoch[mode] * perm.size))]
        self.entity_epoch[mode] = np.asarray(self.raw_entity_set[mode])[perm]
        self.label_epoch[mode] = self.label_full[torch.as_tensor(self.entity_epoch[mode].astype(np.int64), device=self.dev_torch)]
        self.graph_sampler[mode].shuffle_targets(self.entity_epoch[mode])

    def _shuffle_edges(self, mode):
        """link prediction epoch (minibatch.py:281-304): positives + as many negatives (given, or drawn among the node pairs that are neither
        a positive edge in either direction nor a self loop), label 1 / 0, one permutation over both; the sampler gets the end points
        flattened, two consecutive targets root one subgraph"""
        es = self.raw_entity_set[mode]
        pos = np.asarray(es["pos"]).astype(np.int64).reshape(-1, 2)
        if "neg" in es:
            neg = np.asarray(es["neg"]).astype(np.int64).reshape(-1, 2)
        else:
            indptr = self.adj[mode][0] if isinstance(self.adj[mode], (tuple, list)) else self.adj[mode].indptr
            neg = negative_sampling(pos, int(len(indptr) - 1), pos.shape[0], self.dev_torch)
        edge_set = np.concatenate([pos, neg], axis=0)
        label = np.repeat([1, 0], [pos.shape[0], neg.shape[0]])[:, np.newaxis]
        perm = np.random.permutation(edge_set.shape[0])
        if self.percent_per_epoch[mode] < 1.0:
            perm = perm[:int(np.ceil(self.percent_per_epoch[mode] * perm.size))]
        self.entity_epoch[mode] = edge_set[perm]
        self.label_epoch[mode] = torch.from_numpy(label[perm]).to(self.dev_torch)
        self.graph_sampler[mode].shuffle_targets(self.entity_epoch[mode].reshape(-1))

    def epoch_start_reset(self, epoch, mode):
        self.batch_num = -1
        if self.graph_sampler[mode] is None:
            self.instantiate_sampler(mode)
            self.record_subgraphs[mode] = ["noncache"] * self.num_ensemble

    def is_end_epoch(self, mode):
        return self.end_epoch[mode]

    def epoch_end_reset(self, mode):
        self.end_epoch[mode] = False

    def disable_cache(self, mode):
        self.nocache_modes.add(mode)

    def drop_full_graph_info(self, mode):
        pass        # nothing is cached host-side; the graph stays resident for re-sampling

    def get_aug_dim(self, aug_type):
        return getattr(self, f"dim_1hot_{aug_type[:-1]}")

    # ------------------------------------------------------------------ sampler (minibatch.py:344-401)
    def instantiate_sampler(self, mode):
        cfgs = []
        for cfg in deepcopy(self._cfg_ensemble["configs"]):
            method = cfg.pop("method")
            cnt = [len(v) for v in cfg.values()]
            cnt = cnt[0] if cnt else 1
            for i in range(cnt):
                c = {k: v[i] for k, v in cfg.items()}
                c["method"] = "ppr" if (method == "ppr_st" and mode in (VALID, TEST)) else method          # minibatch.py:367-370
                cfgs.append(c)
        self.batch_size[mode] = self._cfg_ensemble["batch_size"]
        bs = self.batch_size[mode]
        per_call = max(bs, (self._per_call // bs) * bs)           # whole batches per sampler call: a batch never straddles two calls
        adj = self.adj[mode]
        indptr, indices = (adj.indptr, adj.indices) if hasattr(adj, "indptr") else adj
        if isinstance(indptr, torch.Tensor) and indptr.is_cuda:          # CSR already resident in HBM: borrow it
            s = PS.ParallelSampler.from_device_csr(indptr, indices, per_call, len(cfgs), self.seed_cpp, strict_reference_compat=self._compat,
                                                   rng=self._rng)
        elif self.bin_adj_files is not None and self.bin_adj_files.get(mode):
            f = self.bin_adj_files[mode]
            s = PS.ParallelSampler([], [], [], per_call, 1, True, True, [], len(cfgs), f["indptr"], f["indices"], f.get("data", ""), self.seed_cpp,
                                   device=self.dev_torch.index, strict_reference_compat=self._compat, rng=self._rng, num_ring=self.NUM_RING)
        else:
            s = PS.ParallelSampler(indptr, indices, [], per_call, 1, True, True, [], len(cfgs), "", "", "", self.seed_cpp,
                                   device=self.dev_torch.index, strict_reference_compat=self._compat, rng=self._rng, num_ring=self.NUM_RING)
        cpp_cfgs = []
        for c in cfgs:
            tf = lambda key: "true" if c.get(key, False) else "false"
            d = {"method": c["method"], "num_roots": str(self.num_roots), "add_self_edge": tf("add_self_edge"),
                 "include_target_conn": "false" if self.prediction_task == "link" else tf("include_target_conn")}      # minibatch.py:376-377
            if c["method"] == "khop":
                d.update(depth=str(c["depth"]), budget=str(c["budget"]))
            elif c["method"] in ("ppr", "ppr_st"):
                d.update(k=str(c["k"]), threshold=str(c.get("threshold", 0)))
            cpp_cfgs.append(d)
        ppr = [c for c in cfgs if c["method"] in ("ppr", "ppr_st")]
        if ppr:         # one table for the largest k (samplers_ensemble.py:212-248)
            top = max(ppr, key=lambda c: int(c["k"]) * (2 if c["method"] == "ppr_st" else 1))
            k = int(top["k"]) * (2 if top["method"] == "ppr_st" else 1)
            alpha, eps = float(top.get("alpha", 0.85)), float(top.get("epsilon", 1e-5))
            fn = fs = ""
            if self.dir_data is not None and not self.dir_data.get("is_adj_changed", False):
                import os
                d = f"{self.dir_data['local']}/{self.name_data}/ppr_float"
                os.makedirs(d, exist_ok=True)
                suffix = f"{'transductive' if self.is_transductive else 'inductive'}_{MODE2STR[mode]}_{alpha}_{eps}_{k}.bin"      # samplers_cpp.py:135-170
                fn, fs = f"{d}/neighs_{suffix}", f"{d}/scores_{suffix}"
            # node task: the mode's targets; link task: every node can be an end point (minibatch.py:380-389)
            prep = np.asarray(self.raw_entity_set[mode]) if self.prediction_task == "node" else np.arange(len(indptr) - 1)
            s.preproc_ppr_approximate(prep, k, alpha, eps, fn, fs)
        self.graph_sampler[mode] = s
        self.sampler_cfgs[mode] = cpp_cfgs
        self.is_stochastic_sampler[mode] = any(c["method"] in ("khop", "ppr_st") for c in cfgs)

    def par_graph_sample(self, mode):
        """one sampler call -> one super-batch per ensemble branch, features gathered, aug one-hots built (minibatch.py:403-426,469-477).
        Runs on the caller's stream.  (Moving it to a side stream so that queued training steps hide it was tried and dropped: the call is
        ~0.35 ms for 160 roots once the result buffers stop being re-allocated, and the overlap did not pay for its hazards.)"""
        s = self.graph_sampler[mode]
        self.num_sampler_calls += 1
        augs = [set(self.aug_feats) & {"hops", "pprs", "drnls"} for _ in self.sampler_cfgs[mode]]
        batches = s.sample_to_device(self.sampler_cfgs[mode], augs)
        for i, b in enumerate(batches):
            feat = PS.gather_rows(self.feat_full, b.orig_node)
            aug = {}
            if "hops" in self.aug_feats:
                aug["hops"] = hop2onehot(b.hop, self.dim_1hot_hop)
            if "pprs" in self.aug_feats:
                aug["pprs"] = ppr2onehot(b.ppr, self.dim_1hot_ppr)
            if "drnls" in self.aug_feats:
                aug["drnls"] = drnl2onehot(b.drnl, self.dim_1hot_drnl)
            sb = _SuperBatch(b, feat, aug, self.num_roots)
            if self.prefetch_canonical:
                sb.prepare_canonical()
            # the views must outlive the sampler's ring slot: one super-batch per branch is in flight, the ring holds NUM_RING
            self.pool[mode][i].append(sb)

    def _front(self, mode, i, bs):
        """super-batch of branch i that holds the next `bs` subgraphs (sampling a new one when the pool is dry)"""
        q = self.pool[mode][i]
        while q and q[0].remaining == 0:
            q.popleft()
        if not q:
            self.par_graph_sample(mode)
            q = self.pool[mode][i]
        assert q[0].remaining >= bs, "sampler call size must be a multiple of the batch size"
        return q[0]

    def one_batch(self, mode=TRAIN, ret_raw_idx=False):
        """minibatch.py:428-487"""
        bs = self._get_cur_batch_size(mode)
        adj_ens, feat_ens, target_ens, aug_ens, idx_raw, size_ens = [], [], [], [], [], []
        for i in range(self.num_ensemble):
            adj, feat, target, sizes, aug, node = self._front(mode, i, bs).take(bs)
            adj_ens.append(adj); feat_ens.append(feat); target_ens.append(target); aug_ens.append(aug); idx_raw.append(node); size_ens.append(sizes)
        a = self.idx_entity_evaluated[mode]
        label = self.label_epoch[mode][a:a + bs]
        self._update_batch_stat(mode, bs)
        ret = OneBatchSubgraph(adj_ens, feat_ens, label, torch.stack(size_ens, 0), target_ens, aug_ens)
        if ret_raw_idx:
            ret.idx_raw = idx_raw
        return ret
