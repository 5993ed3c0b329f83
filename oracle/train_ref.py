"""CPU port of the reference's training step for the north-star config (5-layer GraphSAGE, center pooling, softmax loss):
`MinibatchShallowExtractor.one_batch` + `DeepGNN.step` (shaDow/minibatch.py:428-487, shaDow/models.py:209-237) restated with
the same library calls the reference makes -- scipy CSR -> COO -> torch.sparse (graph_utils.py:48-56), adj_norm_rw with CPU
dropedge (graph_utils.py:81-95), torch.sparse.mm, nn.Linear, norm_feat, clip_grad_norm_(5), Adam.
TEST / BASELINE INFRASTRUCTURE ONLY: used by bench.py's cpu_baseline and `--impl reference` legs."""
import numpy as np
import torch
import torch.nn.functional as F
from torch import nn


class SageLayer(nn.Module):                       # layers.py:447-494
    def __init__(self, din, dout, dropout):
        super().__init__()
        self.f_lin_self, self.f_lin_neigh = nn.Linear(din, dout), nn.Linear(din, dout)
        self.offset, self.scale = nn.Parameter(torch.zeros(2, dout)), nn.Parameter(torch.ones(2, dout))
        self.drop = nn.Dropout(dropout)

    def norm(self, h, i):
        mean = h.mean(1, keepdim=True)
        var = h.var(1, unbiased=False, keepdim=True) + 1e-9
        return (h - mean) * self.scale[i] * torch.rsqrt(var) + self.offset[i]

    def forward(self, x, adj):
        x = self.drop(x)
        return self.norm(F.relu(self.f_lin_self(x)), 0) + self.norm(F.relu(self.f_lin_neigh(torch.sparse.mm(adj, x))), 1)


class RefModel(nn.Module):
    def __init__(self, din, dim, ncls, nlayers, dropout, dropedge, lr):
        super().__init__()
        self.layers = nn.ModuleList(SageLayer(din if i == 0 else dim, dim, dropout) for i in range(nlayers))
        self.cls = nn.Linear(dim, ncls)
        self.cls_offset, self.cls_scale = nn.Parameter(torch.zeros(ncls)), nn.Parameter(torch.ones(ncls))
        self.dropedge = dropedge
        self.opt = torch.optim.Adam(self.parameters(), lr=lr)

    def adj_norm_rw(self, indptr, indices, n):    # coo_scipy2torch + adj_norm_rw (graph_utils.py:48-56, 81-95)
        rows = torch.repeat_interleave(torch.arange(n), torch.as_tensor(np.diff(indptr)))
        vals = torch.ones(len(indices))
        deg0 = torch.zeros(n).index_add_(0, rows, vals)
        if self.dropedge > 0 and self.training:
            idx = torch.floor(torch.rand(int(len(indices) * self.dropedge)) * len(indices)).long()
            vals[idx] = 0
        deg1 = torch.zeros(n).index_add_(0, rows, vals)
        vals = vals / torch.clamp(torch.repeat_interleave(deg1, deg0.long()), min=1)
        return torch.sparse_coo_tensor(torch.stack([rows, torch.as_tensor(indices).long()]), vals, (n, n))

    def step(self, batch, feat, labels):
        self.train()
        self.opt.zero_grad()
        n = batch["indptr"].size - 1
        adj = self.adj_norm_rw(batch["indptr"], batch["indices"], n)
        h = feat
        for l in self.layers:
            h = l(h, adj)
        emb = F.normalize(h[torch.as_tensor(batch["target"])], p=2, dim=1)
        z = self.cls(emb)
        mean, var = z.mean(1, keepdim=True), z.var(1, unbiased=False, keepdim=True) + 1e-9
        preds = (z - mean) * self.cls_scale * torch.rsqrt(var) + self.cls_offset
        loss = F.cross_entropy(preds, labels)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(self.parameters(), 5)
        self.opt.step()
        return float(loss)
