"""shaDow GNN layers on the CUDA kernels of libshadow_b200 -- drop-in for the reference's `shaDow.layers`
(shaDow/layers.py): same class names, constructor signatures, `forward((feat, adj, is_normed, dropedge), sizes_subg)`
contract and parameter names / shapes, so `state_dict()`s interchange with the reference's.

What is different is underneath: the adjacency is a `DeviceCSR` handle on the batch the sampler left in HBM (a scipy CSR is
still accepted on the first layer and uploaded once), the sparse products are hand-written CSR kernels (ops.spmm,
ops.gat_aggregate), activation + norm_feat is one fused kernel (ops.act_norm); only the dense `nn.Linear`s use cuBLAS.
"""
import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from .ops import DeviceCSR

# activation table of the reference (layers.py:26-39)
_ACT = {
    "relu": lambda dim_out: nn.ReLU(),
    "I": lambda dim_out: nn.LeakyReLU(negative_slope=1),
    "elu": lambda dim_out: nn.ELU(),
    "tanh": lambda dim_out: nn.Tanh(),
    "leakyrelu": lambda dim_out: nn.LeakyReLU(negative_slope=0.2),
    "prelu": lambda dim_out: nn.PReLU(),
    "prelu+": lambda dim_out: nn.PReLU(num_parameters=dim_out),
}
F_ACT = _ACT


def get_torch_act(act, args):
    return _ACT[act](args.get("dim_out", 1) if isinstance(args, dict) else 1)


class _Dropedge:
    """counter-based dropedge stream: the position lives in a device tensor that is bumped on the stream (so a captured CUDA
    graph draws fresh edges on every replay); seeded by torch.initial_seed()"""
    _step = {}

    @classmethod
    def next(cls, device):
        key = str(device)
        if key not in cls._step:
            cls._step[key] = torch.zeros(1, dtype=torch.int32, device=device)
        cls._step[key].add_(1)
        return int(torch.initial_seed()) & 0xFFFFFFFF, cls._step[key]


def _as_device_csr(adj, device):
    if isinstance(adj, DeviceCSR):
        return adj
    import scipy.sparse as sp
    if sp.issparse(adj):
        return DeviceCSR.from_scipy(adj.tocsr(), device)
    raise TypeError(f"adjacency must be a DeviceCSR or a scipy CSR matrix, got {type(adj)}")


class EnsembleDummy(nn.Module):
    """single branch: identity (layers.py:42-53)"""

    def __init__(self, dim_in=0, dim_out=0, **kwargs):
        super().__init__()

    def forward(self, Xi):
        assert len(Xi) == 1, "the dummy ensembler takes exactly one branch"
        return Xi[0]


class shaDowLayer(nn.Module):
    """common part of the conv layers: dropout, activation, norm_feat parameters (layers.py:299-373)"""

    def __init__(self, dim_in, dim_out, dropout=0.0, act="relu", norm="norm_feat", **kwargs):
        super().__init__()
        self.dim_in, self.dim_out, self.dropout = dim_in, dim_out, dropout
        self.act_name = act
        self.act = _ACT[act](dim_out)
        self.f_dropout = nn.Dropout(p=dropout)
        self.norm = norm
        self.norm_dim = kwargs.get("norm_dim", (1, dim_out))
        if norm == "norm_feat":
            self.offset = nn.Parameter(torch.zeros(self.norm_dim))
            self.scale = nn.Parameter(torch.ones(self.norm_dim))
        elif norm not in ("none",):
            raise NotImplementedError(f"norm '{norm}' (the reference's pairnorm path stops at a breakpoint(), layers.py:358)")

    def _act_norm(self, z, idx):
        """norm_feat_idx(act(z)) for one (branch[, head]) slot of scale/offset"""
        if self.act_name in ops.ACT_ID:
            if self.norm == "norm_feat":
                return ops.act_norm(z, self.scale[idx], self.offset[idx], self.act_name, True)
            return ops.act_norm(z, None, None, self.act_name, False)
        a = self.act(z)                       # PReLU carries parameters: torch activation, fused kernel for the norm only
        if self.norm == "norm_feat":
            return ops.act_norm(a, self.scale[idx], self.offset[idx], "I", True)
        return a

    def _norm_only(self, a, idx):
        if self.norm == "norm_feat":
            return ops.act_norm(a, self.scale[idx], self.offset[idx], "I", True)
        return a


class MLP(shaDowLayer):
    """norm(act(W dropout(x) + b))  (layers.py:376-400)"""

    def __init__(self, dim_in, dim_out, dropout=0.0, act="relu", norm="norm_feat", **kwargs):
        assert norm in ("norm_feat", "none")
        kwargs["norm_dim"] = (1, dim_out)
        super().__init__(dim_in, dim_out, dropout=dropout, act=act, norm=norm, **kwargs)
        self.f_lin = nn.Linear(dim_in, dim_out)

    def forward(self, feat_in):
        if self.act_name in ops.ACT_ID:
            return ops.linear_act_norm(self.f_dropout(feat_in), self.f_lin, getattr(self, "scale", None), getattr(self, "offset", None), 0, self.act_name,
                                       self.norm == "norm_feat")
        return self._act_norm(self.f_lin(self.f_dropout(feat_in)), 0)


class MLPSGC(MLP):
    """MLP taking the conv-layer input tuple (layers.py:403-414)"""

    def forward(self, inputs, **kwargs):
        assert isinstance(inputs, (list, tuple)) and len(inputs) == 4
        return super().forward(inputs[0]), None, None, None


class GCN(shaDowLayer):
    """norm(act(W (A_sym dropout(x)) + b))  (layers.py:417-444)"""

    def __init__(self, dim_in, dim_out, dropout=0.0, act="relu", norm="norm_feat", **kwargs):
        kwargs["norm_dim"] = (1, dim_out)
        super().__init__(dim_in, dim_out, dropout=dropout, act=act, norm=norm, **kwargs)
        self.f_lin = nn.Linear(dim_in, dim_out, bias=True)

    def forward(self, inputs, sizes_subg):
        feat_in, adj, is_normed, dropedge = inputs
        feat_in = self.f_dropout(feat_in)
        adj = _as_device_csr(adj, feat_in.device)
        if not is_normed:
            adj.normalize_sym(dropedge, *_Dropedge.next(feat_in.device))
        agg = ops.spmm(adj, feat_in)
        if self.act_name in ops.ACT_ID:
            feat_out = ops.linear_act_norm(agg, self.f_lin, getattr(self, "scale", None), getattr(self, "offset", None), 0, self.act_name,
                                           self.norm == "norm_feat")
        else:
            feat_out = self._act_norm(self.f_lin(agg), 0)
        return feat_out, adj, True, 0.0


class GraphSAGE(shaDowLayer):
    """norm_0(act(W_s x + b_s)) + norm_1(act(W_n (A_rw x) + b_n)), x = dropout(in)  (layers.py:447-494)"""

    def __init__(self, dim_in, dim_out, dropout=0.0, act="relu", norm="norm_feat", **kwargs):
        kwargs["norm_dim"] = (2, dim_out)
        super().__init__(dim_in, dim_out, dropout=dropout, act=act, norm=norm, **kwargs)
        self.f_lin_self = nn.Linear(dim_in, dim_out)
        self.f_lin_neigh = nn.Linear(dim_in, dim_out)

    def forward(self, inputs, sizes_subg):
        feat_in, adj, is_normed, dropedge = inputs
        adj = _as_device_csr(adj, feat_in.device)
        if not is_normed:
            adj.normalize_rw(dropedge, *_Dropedge.next(feat_in.device))
        feat_in = self.f_dropout(feat_in)
        if self.act_name in ops.ACT_ID:                 # whole layer body as one autograd node (fused forward / backward launches)
            out = ops.sage_layer(feat_in, adj, self.f_lin_self, self.f_lin_neigh, getattr(self, "scale", None), getattr(self, "offset", None),
                                 self.act_name, self.norm == "norm_feat")
            return out, adj, True, 0.0
        h_self = self._act_norm(self.f_lin_self(feat_in), 0)
        h_neigh = self._act_norm(self.f_lin_neigh(ops.spmm(adj, feat_in)), 1)
        return h_self + h_neigh, adj, True, 0.0


class GIN(shaDowLayer):
    """norm(act(MLP(A x + (1 + eps) x)))  (layers.py:497-536); returns is_normed=False like the reference"""

    def __init__(self, dim_in, dim_out, dropout=0.0, act="relu", norm="norm_feat", eps=0, **kwargs):
        kwargs["norm_dim"] = (1, dim_out)
        super().__init__(dim_in, dim_out, dropout=dropout, act=act, norm=norm, **kwargs)
        self.mlp = nn.Sequential(nn.Linear(dim_in, dim_out, bias=True), nn.ReLU(), nn.Linear(dim_out, dim_out, bias=True))
        self.eps = nn.Parameter(torch.Tensor([eps]))

    def forward(self, inputs, sizes_subg):
        feat_in, adj, is_normed, dropedge = inputs
        assert not is_normed
        feat_in = self.f_dropout(feat_in)
        first = not isinstance(adj, DeviceCSR) or adj.normed is None
        adj = _as_device_csr(adj, feat_in.device)
        if first:                                  # the reference drops / rescales only while adj is still scipy (first layer)
            adj.normalize_gin(dropedge, *_Dropedge.next(feat_in.device))
        feat_aggr = ops.spmm(adj, feat_in) + (1 + self.eps) * feat_in
        return self._act_norm(self.mlp(feat_aggr), 0), adj, False, 0.0


class GAT(shaDowLayer):
    """multi-head attention aggregation of the reference (layers.py:539-645).  norm slots: [0,k] neighbour branch, [1,k] self."""

    def __init__(self, dim_in, dim_out, dropout=0.0, act="relu", norm="norm_feat", mulhead=1, **kwargs):
        self.mulhead = mulhead
        assert dim_out % mulhead == 0, "invalid output dimension: need to be divisible by mulhead"
        self.dim_slice = dim_out // mulhead
        kwargs["norm_dim"] = (2, mulhead, self.dim_slice)
        super().__init__(dim_in, dim_out, dropout=dropout, act=act, norm=norm, **kwargs)
        self.att_act = nn.LeakyReLU(negative_slope=0.2)
        self.f_lin = nn.ModuleList(nn.Linear(dim_in, dim_out, bias=True) for _ in range(2))     # [0] self, [1] neighbour
        self.attention = nn.Parameter(torch.ones(2, mulhead, self.dim_slice))
        nn.init.xavier_uniform_(self.attention)

    def forward(self, inputs, sizes_subg):
        feat_in, adj, is_normed, dropedge = inputs
        adj = _as_device_csr(adj, feat_in.device)
        if not is_normed:
            adj.mask_only(dropedge, *_Dropedge.next(feat_in.device))
        feat_in = self.f_dropout(feat_in)
        N, H, d = feat_in.shape[0], self.mulhead, self.dim_slice
        if feat_in.is_cuda and self.act_name in ops.ACT_ID and ops.gat_layer_supported(feat_in, self.f_lin[0].weight, H):
            has_norm = self.norm == "norm_feat"       # (without norm_feat the scale / offset arguments are placeholders the kernels never read)
            out = ops.gat_layer(feat_in, adj, self.f_lin[0], self.f_lin[1], self.attention, self.scale if has_norm else self.attention,
                                self.offset if has_norm else self.attention, self.act_name, H, has_norm)
            return out, adj, True, 0.0
        h_self = self.act(self.f_lin[0](feat_in))
        h_neigh = self.act(self.f_lin[1](feat_in))
        a_self = self.att_act((h_self.view(N, H, d) * self.attention[0]).sum(-1))
        a_neigh = self.att_act((h_neigh.view(N, H, d) * self.attention[1]).sum(-1))
        agg = ops.gat_aggregate(adj, a_self, a_neigh, h_neigh, H)
        outs = []
        for k in range(H):                                  # norm_feat per (branch, head) over the head's slice
            sl = slice(k * d, (k + 1) * d)
            outs.append(self._norm_only(h_self[:, sl].contiguous(), (1, k)) + self._norm_only(agg[:, sl].contiguous(), (0, k)))
        return torch.cat(outs, dim=1) / 2, adj, True, 0.0


class GATScatter(shaDowLayer):
    """DGL-style variant (layers.py:648-743): one attention vector on the source features"""

    def __init__(self, dim_in, dim_out, dropout=0.0, act="relu", norm="norm_feat", mulhead=1, **kwargs):
        kwargs["norm_dim"] = (1, dim_out)
        super().__init__(dim_in, dim_out, dropout=dropout, act=act, norm=norm, **kwargs)
        self.mulhead = mulhead
        self.att_act = nn.LeakyReLU(negative_slope=0.2)
        assert dim_out % mulhead == 0, "invalid output dimension: need to be divisible by mulhead"
        self.dim_slice = dim_out // mulhead
        self.f_lin = nn.ModuleList(nn.Linear(dim_in, dim_out, bias=True) for _ in range(2))
        self.attention = nn.Parameter(torch.empty(1, mulhead, self.dim_slice))
        gain = nn.init.calculate_gain(act)
        nn.init.xavier_normal_(self.f_lin[0].weight, gain=gain)
        nn.init.xavier_normal_(self.f_lin[1].weight, gain=gain)
        nn.init.xavier_normal_(self.attention, gain=gain)

    def forward(self, inputs, sizes_subg):
        feat_in, adj, is_dropped, dropedge = inputs
        N, H, d = feat_in.shape[0], self.mulhead, self.dim_slice
        h = self.f_dropout(feat_in)
        feat_src = self.f_lin[0](h)
        adj = _as_device_csr(adj, feat_in.device)
        if not is_dropped:
            adj.mask_only(dropedge, *_Dropedge.next(feat_in.device))
        el = self.att_act((feat_src.view(N, H, d) * self.attention).sum(-1))
        agg = ops.gat_aggregate(adj, torch.zeros_like(el), el, feat_src, H)
        return self._act_norm(agg + self.f_lin[1](h), 0), adj, True, 0.0


class ResPool(nn.Module):
    """residue over layers + subgraph pooling + root readout (layers.py:57-233)"""

    def __init__(self, dim_in, dim_out, num_layers, type_res, type_pool, dropout, act, args_pool=None, prediction_task="node"):
        super().__init__()
        self.dim_out, self.type_pool, self.type_res, self.prediction_task = dim_out, type_pool, type_res, prediction_task
        cat = type_res in ("cat", "concat")
        if type_pool == "center":
            if type_res == "none":
                if prediction_task == "node":
                    self.dim_in = self.dim_out = 0
                else:
                    self.dim_in = dim_in
            else:
                self.dim_in = num_layers * dim_in if cat else dim_in
        else:
            self.dim_in = 2 * dim_in * (num_layers if cat else 1)
            if type_pool == "sort":
                assert args_pool is not None and "k" in args_pool, "Sort pooling needs the budget k as input!"
                self.k = args_pool["k"]
                self.nn_pool = nn.Sequential(nn.Dropout(p=dropout), nn.Linear(self.k * (self.dim_in // 2), self.dim_in // 2), _ACT[act](dim_out))
        if self.dim_in > 0 and self.dim_out > 0:
            self.nn = nn.Sequential(nn.Dropout(p=dropout), nn.Linear(self.dim_in, self.dim_out, bias=True), _ACT[act](dim_out))
            self.offset = nn.Parameter(torch.zeros(self.dim_out))
            self.scale = nn.Parameter(torch.ones(self.dim_out))

    def f_residue(self, feat_l):
        if self.type_res in ("cat", "concat"):
            return torch.cat(feat_l, dim=1)
        if self.type_res == "sum":
            return torch.stack(feat_l, dim=0).sum(dim=0)
        if self.type_res == "max":
            return torch.stack(feat_l, dim=0).max(dim=0).values
        raise NotImplementedError(self.type_res)

    def aggr_target_emb(self, feat):
        if self.prediction_task == "node":
            return feat
        b, f = feat.shape
        pair = feat.reshape(b // 2, 2, f)
        return pair[:, 0] * pair[:, 1]

    def forward(self, feats_in_l, idx_targets, sizes_subg):
        idx_targets = torch.as_tensor(idx_targets, device=feats_in_l[-1].device).long()
        # (index_select, not x[idx]: its backward is one index_add_ -- the roots of a batch are distinct rows -- instead of a sort + index_put)
        if self.type_pool == "center":
            if self.type_res == "none":
                feat_in = feats_in_l[-1].index_select(0, idx_targets)
                if self.prediction_task == "node":
                    return feat_in
            else:
                feat_in = self.f_residue([f.index_select(0, idx_targets) for f in feats_in_l])
            feat_in = self.aggr_target_emb(feat_in)
        elif self.type_pool in ("max", "mean", "sum"):
            if self.type_res == "none":
                feat_pool = ops.segment_pool(feats_in_l[-1], sizes_subg, self.type_pool)
                feat_root = feats_in_l[-1].index_select(0, idx_targets)
            else:
                feat_pool = self.f_residue([ops.segment_pool(f, sizes_subg, self.type_pool) for f in feats_in_l])
                feat_root = self.f_residue([f.index_select(0, idx_targets) for f in feats_in_l])
            feat_in = torch.cat([self.aggr_target_emb(feat_root), feat_pool], dim=1)
        elif self.type_pool == "sort":
            if self.type_res == "none":
                feat_pool_in, feat_root = feats_in_l[-1], feats_in_l[-1].index_select(0, idx_targets)
            else:
                feat_pool_in, feat_root = self.f_residue(feats_in_l), self.f_residue([f.index_select(0, idx_targets) for f in feats_in_l])
            feat_pool = self.nn_pool(_sort_pool(feat_pool_in, sizes_subg, self.k))
            feat_in = torch.cat([self.aggr_target_emb(feat_root), feat_pool], dim=1)
        else:
            raise NotImplementedError(self.type_pool)
        return ops.act_norm(self.nn(feat_in), self.scale, self.offset, "I", True)


def _sort_pool(x, sizes_subg, k):
    """global_sort_pool (PyG): per subgraph keep the k rows with the largest last channel, zero-pad, flatten"""
    B, Fd = sizes_subg.numel(), x.shape[1]
    batch = torch.repeat_interleave(torch.arange(B, device=x.device), sizes_subg.long())
    key = x[:, -1].detach()
    order = torch.argsort(key, descending=True, stable=True)
    order = order[torch.argsort(batch[order], stable=True)]
    start = torch.cumsum(sizes_subg.long(), 0) - sizes_subg.long()
    rank = torch.arange(x.shape[0], device=x.device) - start[batch[order]]
    keep = rank < k
    out = x.new_zeros(B, k, Fd)
    out[batch[order][keep], rank[keep]] = x[order][keep]
    return out.reshape(B, k * Fd)


class EnsembleAggregator(nn.Module):
    """attention over the ensemble branches (layers.py:236-296)"""

    def __init__(self, dim_in, dim_out, num_ensemble, dropout=0.0, act="leakyrelu", type_dropout="none"):
        super().__init__()
        self.dim_in, self.dim_out, self.dropout = dim_in, dim_out, dropout
        self.act = nn.ModuleList(_ACT[act](dim_out) for _ in range(num_ensemble))
        self.f_lin = nn.Linear(dim_in, dim_out, bias=True)
        self.f_dropout = nn.Dropout(p=dropout)
        self.q = nn.Parameter(torch.ones(dim_out))
        assert type_dropout in ("none", "feat", "coef")
        self.type_dropout = type_dropout

    def forward(self, Xi):
        omega = []
        for i, X in enumerate(Xi):
            if self.type_dropout == "coef":
                X_ = self.f_dropout(X)
            elif self.type_dropout == "feat":
                Xi[i] = self.f_dropout(X)
                X_ = Xi[i]
            else:
                X_ = X
            omega.append(self.act[i](self.f_lin(X_)).mm(self.q.view(-1, 1)))
        w = F.softmax(torch.cat(omega, 1), dim=1)
        return sum(w[:, i].view(-1, 1) * X for i, X in enumerate(Xi))
