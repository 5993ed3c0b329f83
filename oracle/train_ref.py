"""Port of the reference's training loop for the north-star config (5-layer GraphSAGE, center pooling, softmax loss), restated
with the same library calls the reference makes.  TEST / BASELINE INFRASTRUCTURE ONLY: used by bench.py's cpu_baseline and
`--impl reference` legs (the reference's Python tree cannot travel to the GPU box; its C++ sampler does, as oracle/_ref).

  RefModel       `DeepGNN.step` (shaDow/models.py:209-237): scipy CSR -> COO -> torch.sparse on the model's device
                 (graph_utils.py:48-56), adj_norm_rw with the CPU-side dropedge draw (graph_utils.py:81-95), torch.sparse.mm, nn.Linear,
                 norm_feat, F.normalize, classifier MLP, CE loss, clip_grad_norm_(5), torch.optim.Adam.  `device` = "cpu" (the host-core
                 baseline) or "cuda" (the reference as its users run it: `--gpu 0`, model and full feature tensor on the same B200).
  RefMinibatch   `MinibatchShallowExtractor` in train mode (shaDow/minibatch.py:403-487) with its three cache behaviours
                 (minibatch.py:69-91,311-333,410-421; shaDow/globals.py:51-52):
                   "record"  epoch 0: sampler call of 500 roots (minibatch.py:397), subgraphs stored in the per-root dict
                   "reuse"   epochs >= 1 of a deterministic sampler (PPR): the sampler only returns the roots (dummy_sampler,
                             PS.cpp:653-659), subgraphs come from the dict
                   "nocache" `--nocache all`: every epoch samples
                 then pool -> collate (Subgraph.cat_to_block_diagonal, graph.py:280-320) -> to_csr_sp -> feat_full[node] -> labels.
"""
from collections import deque

import numpy as np
import scipy.sparse as sp
import torch
import torch.nn.functional as F
from torch import nn


class SageLayer(nn.Module):                       # layers.py:447-494
    def __init__(self, din, dout, dropout):
        super().__init__()
        self.f_lin_self, self.f_lin_neigh = nn.Linear(din, dout), nn.Linear(din, dout)
        self.offset, self.scale = nn.Parameter(torch.zeros(2, dout)), nn.Parameter(torch.ones(2, dout))
        self.drop = nn.Dropout(dropout)

    def norm(self, h, i):                         # layers.py:329-338
        mean = h.mean(dim=1, keepdim=True)
        var = h.var(dim=1, unbiased=False).view(h.shape[0], 1) + 1e-9
        return (h - mean) * self.scale[i] * torch.rsqrt(var) + self.offset[i]

    def forward(self, x, adj):
        x = self.drop(x)
        return self.norm(F.relu(self.f_lin_self(x)), 0) + self.norm(F.relu(self.f_lin_neigh(torch.sparse.mm(adj, x))), 1)


class RefModel(nn.Module):
    def __init__(self, din, dim, ncls, nlayers, dropout, dropedge, lr, device="cpu"):
        super().__init__()
        self.layers = nn.ModuleList(SageLayer(din if i == 0 else dim, dim, dropout) for i in range(nlayers))
        self.cls = nn.Linear(dim, ncls)
        self.cls_offset, self.cls_scale = nn.Parameter(torch.zeros(ncls)), nn.Parameter(torch.ones(ncls))
        self.dropedge = dropedge
        self.device = torch.device(device)
        self.to(self.device)
        self.opt = torch.optim.Adam(self.parameters(), lr=lr)

    def adj_norm_rw(self, adj_sp):
        """GraphSAGE.forward, first layer (layers.py:466-469): coo_scipy2torch(adj.tocoo()).to(device), then adj_norm_rw (graph_utils.py:81-95)"""
        coo = adj_sp.tocoo()
        i = torch.LongTensor(np.vstack((coo.row, coo.col)))
        v = torch.FloatTensor(coo.data).type(torch.get_default_dtype())
        adj = torch.sparse_coo_tensor(i, v, torch.Size(coo.shape)).to(self.device)
        vals, rows = adj._values(), adj._indices()[0]
        n = coo.shape[0]
        deg0 = torch.zeros(n, device=self.device).index_add_(0, rows, vals)                       # torch_scatter.scatter(..., reduce="sum")
        if self.dropedge > 0 and self.training:
            e = vals.size(0)
            idx = torch.floor(torch.rand(int(e * self.dropedge)) * e).long()                     # drawn on the host like the reference
            vals[idx] = 0
            deg1 = torch.zeros(n, device=self.device).index_add_(0, rows, vals)
        else:
            deg1 = deg0
        vals /= torch.clamp(torch.repeat_interleave(deg1, deg0.long()), min=1)
        return adj

    def step(self, adj_sp, feat, target, labels):
        """feat / labels already on the model's device (one_batch put them there)"""
        self.train()
        self.opt.zero_grad()
        adj = self.adj_norm_rw(adj_sp)
        h = feat
        for l in self.layers:
            h = l(h, adj)
        emb = F.normalize(h[torch.as_tensor(target, device=self.device)], p=2, dim=1)
        z = self.cls(emb)
        mean, var = z.mean(1, keepdim=True), z.var(1, unbiased=False, keepdim=True) + 1e-9
        preds = (z - mean) * self.cls_scale * torch.rsqrt(var) + self.cls_offset
        loss = F.cross_entropy(preds, labels)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(self.parameters(), 5)
        self.opt.step()
        return loss


def _narrow(sub, cap_node_subg, cap_edge_subg):
    """dtype narrowing of Subgraph.__post_init__ (graph.py:228-255) + the all-ones `data` collapsed to a broadcast view"""
    f = lambda n: np.uint16 if n < 2 ** 16 else np.uint32
    return dict(indptr=sub["indptr"].astype(f(cap_edge_subg), copy=False), indices=sub["indices"].astype(f(cap_node_subg), copy=False),
                data=np.broadcast_to(np.array([1.]), sub["indices"].size), node=sub["node"].astype(np.uint32, copy=False),
                edge_index=sub["edge_index"].astype(np.uint32, copy=False), target=sub["target"].astype(f(cap_node_subg), copy=False))


def cat_to_block_diagonal(subgs):
    """Subgraph.cat_to_block_diagonal (graph.py:280-320), statement by statement"""
    offset_indices = np.cumsum([s["node"].size for s in subgs])
    offset_indptr = np.cumsum([s["edge_index"].size for s in subgs])
    offset_indices[1:] = offset_indices[:-1]; offset_indices[0] = 0
    offset_indptr[1:] = offset_indptr[:-1]; offset_indptr[0] = 0
    node_batch = np.concatenate([s["node"] for s in subgs])
    edge_index_batch = np.concatenate([s["edge_index"] for s in subgs])
    data_batch = np.concatenate([s["data"] for s in subgs])
    target_itr = [s["target"].astype(np.int64) for s in subgs]
    indptr_itr = [s["indptr"].astype(np.int64) for s in subgs]
    indices_itr = [s["indices"].astype(np.int64) for s in subgs]
    target_batch, indptr_batch, indices_batch = [], [], []
    for i in range(len(subgs)):
        target_batch.append(target_itr[i] + offset_indices[i])
        if i > 0:
            indptr_itr[i] = indptr_itr[i][1:]
        indptr_batch.append(indptr_itr[i] + offset_indptr[i])
        indices_batch.append(indices_itr[i] + offset_indices[i])
    out = dict(indptr=np.concatenate(indptr_batch), indices=np.concatenate(indices_batch), data=data_batch, node=node_batch,
               edge_index=edge_index_batch, target=np.concatenate(target_batch))
    if np.all(out["data"] == 1.):                                   # __post_init__ of the concatenated Subgraph
        out["data"] = np.broadcast_to(np.array([1.]), out["data"].size)
    assert out["node"].size == out["indptr"].size - 1 and out["indices"].size == out["indptr"][-1]          # check_valid
    return out


class RefMinibatch:
    def __init__(self, call_sampler, conv, roots, labels_all, feat_dev, batch, k, num_edges_full, device, cache_mode="record"):
        """call_sampler(return_target_only: bool) -> SubgraphStructVec ; conv(vec) -> list of per-subgraph dicts (list -> numpy,
        samplers_ensemble.py:254-265)"""
        assert cache_mode in ("record", "reuse", "nocache")
        self.call, self.conv, self.roots, self.B, self.dev = call_sampler, conv, np.asarray(roots), batch, torch.device(device)
        self.feat, self.labels_epoch = feat_dev, torch.as_tensor(labels_all[self.roots.astype(np.int64)]).to(self.dev)
        self.cap_node, self.cap_edge = k, min(num_edges_full, k * k)            # samplers_ensemble.py:266-274
        self.cache, self.pool, self.mode, self.cursor = {}, deque(), cache_mode, 0

    def set_mode(self, mode):
        self.mode = mode

    def par_graph_sample(self):                                     # minibatch.py:403-426
        if self.mode == "reuse":
            for s in self.conv(self.call(True)):                    # roots only
                self.pool.append(self.cache[int(s["node"][0])])
        else:
            for s in self.conv(self.call(False)):
                s = _narrow(s, self.cap_node, self.cap_edge)
                if self.mode == "record":
                    self.cache[int(s["node"][s["target"]][0])] = s
                self.pool.append(s)

    def one_batch(self):                                            # minibatch.py:428-487
        while len(self.pool) < self.B:
            self.par_graph_sample()
        col = cat_to_block_diagonal([self.pool.popleft() for _ in range(self.B)])
        n = col["indptr"].size - 1
        adj = sp.csr_matrix((col["data"], col["indices"], col["indptr"]), shape=(n, n))           # to_csr_sp
        feat = self.feat[torch.as_tensor(col["node"].astype(np.int64))].to(self.dev)              # feat_full[subgs.node].to(dev)
        label = self.labels_epoch[self.cursor:self.cursor + self.B]
        self.cursor = (self.cursor + self.B) % self.roots.size
        return adj, feat, col["target"], label
