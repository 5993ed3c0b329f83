"""DeepGNN on the CUDA layers -- drop-in for the reference's `shaDow.models.DeepGNN` (shaDow/models.py:16-237):
same constructor signature, same sub-module names (so checkpoints interchange), same `forward` / `step` contract.

Differences under the hood: layers come from shadow_gnn_b200.layers, the optimizer step (gradient-norm clip at 5 + Adam)
is one fused pass over a flat parameter buffer (ops.FlatAdamClip), and in the data-parallel run the same flat gradient
buffer is all-reduced with a single NCCL call before the clip, so the clip sees the reduced gradient exactly like the
single-GPU run.
"""
from typing import Any, Dict

import torch
import torch.nn.functional as F
from torch import nn

from . import layers
from .ops import FlatAdamClip
from .parallel import allreduce_flat_gradients

TRAIN, VALID, TEST = 0, 1, 2


class DeepGNN(nn.Module):
    NAME2CLS = {"mlp": layers.MLP, "gcn": layers.GCN, "gin": layers.GIN, "sage": layers.GraphSAGE, "gat": layers.GAT,
                "gatscat": layers.GATScatter, "sgc": layers.MLPSGC, "sign": layers.MLPSGC}

    def __init__(self, dim_feat_raw: int, dim_feat_smooth: int, dim_label_raw: int, dim_label_smooth: int, arch_gnn: Dict[str, Any],
                 aug_feat, num_ensemble: int, train_params: Dict[str, Any], prediction_task: str):
        super().__init__()
        assert prediction_task in {"link", "node"}
        assert dim_feat_raw <= dim_feat_smooth
        self.prediction_task = prediction_task
        self.num_gnn_layers, self.num_cls_layers = arch_gnn["num_layers"], arch_gnn["num_cls_layers"]
        self.dropout, self.dropedge = train_params["dropout"], train_params["dropedge"]
        self.mulhead = int(arch_gnn["heads"])
        self.branch_sharing = arch_gnn["branch_sharing"]
        self.type_feature_augment = aug_feat
        self.num_classes, self.dim_label_in, self.dim_feat_in, self.dim_hid = dim_label_raw, dim_label_smooth, dim_feat_smooth, arch_gnn["dim"]
        act, layer_norm = arch_gnn["act"], arch_gnn["layer_norm"]
        self.feat_aug_ops = arch_gnn["feature_augment_ops"]
        aug_layers, conv_layers, res_pool_layers = [], [], []
        for i in range(num_ensemble):
            dim_aug_add = 0
            if len(self.type_feature_augment) > 0:
                dim_aug_out = self.dim_feat_in if self.feat_aug_ops == "sum" else self.dim_hid
                dim_aug_add = 0 if self.feat_aug_ops == "sum" else dim_aug_out
                aug_layers.append(nn.ModuleList(nn.Linear(d, dim_aug_out) for _, d in self.type_feature_augment))
            if i == 0 or not self.branch_sharing:
                convs = [DeepGNN.NAME2CLS[arch_gnn["aggr"]](
                    (self.dim_feat_in + self.dim_label_in + dim_aug_add) if j == 0 else self.dim_hid, self.dim_hid,
                    dropout=self.dropout, act=act, norm=layer_norm, mulhead=self.mulhead) for j in range(self.num_gnn_layers)]
                conv_layers.append(nn.Sequential(*convs))
            else:
                conv_layers.append(conv_layers[-1])
            pool = arch_gnn["pooling"].split("-")
            args_pool = {"k": int(pool[1])} if pool[0].lower() == "sort" else {}
            res_pool_layers.append(layers.ResPool(self.dim_hid, self.dim_hid, self.num_gnn_layers, arch_gnn["residue"].lower(), pool[0].lower(),
                                                  dropout=self.dropout, act=act, args_pool=args_pool, prediction_task=prediction_task))
        self.aug_layers = nn.ModuleList(aug_layers) if aug_layers else []
        self.conv_layers = nn.ModuleList(conv_layers)
        self.res_pool_layers = nn.ModuleList(res_pool_layers)
        if num_ensemble == 1:
            self.ensembler = layers.EnsembleDummy()
        else:
            self.ensembler = layers.EnsembleAggregator(self.dim_hid, self.dim_hid, num_ensemble, dropout=self.dropout,
                                                       type_dropout=train_params["ensemble_dropout"], act=arch_gnn["ensemble_act"])
        norm_type = "norm_feat" if prediction_task == "node" else "none"
        cls = []
        for i in range(self.num_cls_layers):
            last = i == self.num_cls_layers - 1
            cls.append(layers.MLP(dim_in=self.dim_hid, dim_out=self.num_classes if last else self.dim_hid, act="I" if last else act,
                                  dropout=0.0 if last else self.dropout, norm=norm_type))
        self.classifier = nn.Sequential(*cls)
        self.lr = train_params["lr"]
        self.sigmoid_loss = arch_gnn["loss"] == "sigmoid"
        self.num_ensemble = num_ensemble
        self.optimizer = None            # created on first training step, once the parameters live on their device
        self._world = None

    # ---- models.py:156-166 ----
    def _loss(self, preds, labels):
        if self.sigmoid_loss:
            assert preds.shape == labels.shape
            return nn.BCEWithLogitsLoss()(preds, labels.type(preds.dtype)) * preds.shape[1]
        if labels.dim() == 2:
            labels = labels.max(dim=1)[1]
        return nn.CrossEntropyLoss()(preds, labels)

    # ---- models.py:169-204 ----
    def forward(self, mode, feat_ens, adj_ens, target_ens, size_subg_ens, feat_aug_ens, dropedge):
        emb_subg_ens = []
        for i in range(len(feat_ens)):
            x = feat_ens[i]
            tgt = torch.as_tensor(target_ens[i], device=x.device).long()
            if self.dim_label_in > 0 and mode == TRAIN:
                x[tgt, -self.dim_label_in:] = 0
            for ia, (ta, _dim) in enumerate(self.type_feature_augment):
                emb = self.aug_layers[i][ia](feat_aug_ens[i][ta])
                if self.feat_aug_ops == "sum":
                    x = torch.cat([x[:, :self.dim_feat_in] + emb, x[:, self.dim_feat_in:]], dim=1) if x.shape[1] > self.dim_feat_in else x + emb
                else:
                    x = torch.cat([x, emb], dim=1)
            xjk, xmd = [], (x, adj_ens[i], False, dropedge)
            for md in self.conv_layers[i]:
                xmd = md(xmd, sizes_subg=size_subg_ens[i])
                xjk.append(xmd[0])
            emb = self.res_pool_layers[i](xjk, tgt, size_subg_ens[i])
            emb_subg_ens.append(F.normalize(emb, p=2, dim=1))
        return self.classifier(self.ensembler(emb_subg_ens)), emb_subg_ens

    def predict(self, preds):
        return torch.sigmoid(preds) if self.sigmoid_loss else F.softmax(preds, dim=1)

    def _ensure_optimizer(self):
        if self.optimizer is None:
            self.optimizer = FlatAdamClip(list(self.parameters()), lr=self.lr, max_norm=5.0)
            import torch.distributed as dist
            self._world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
            if self._world > 1:        # replicas must start from the same weights whatever each rank's RNG state was
                dist.broadcast(self.optimizer.flat, src=0)
            if self.optimizer.planes is not None:      # weights restored from a checkpoint later on: the TF32 planes follow
                self.register_load_state_dict_post_hook(lambda module, incompatible: module.optimizer.planes.mark_stale())
        return self.optimizer

    # ---- models.py:209-237 ----
    def step(self, mode, status, batch_data):
        assert status in ("running", "final")
        args = batch_data.to_dict({"feat_ens", "adj_ens", "target_ens", "size_subg_ens", "feat_aug_ens"})
        label_targets = batch_data.label
        if label_targets.dim() == 1 and self.num_classes > 1:
            label_targets = F.one_hot(label_targets.to(torch.int64), num_classes=self.num_classes)
        if mode == TRAIN and status == "running":
            opt = self._ensure_optimizer()
            self.train()
            opt.zero_grad()
            preds, emb_ens = self(mode, dropedge=self.dropedge, **args)
            loss = self._loss(preds, label_targets)
            loss.backward()
            # one NCCL all-reduce of the flat gradient bucket; the clip then sees the averaged gradient
            # the exchange: fused into the optimizer over NVLink peer memory (ops._P2PGrad), else one NCCL all-reduce of the flat bucket;
            # either way the clip sees the averaged gradient
            opt.step(grad_scale=1.0 if opt.p2p is not None else allreduce_flat_gradients(opt.grad))
        else:
            self.eval()
            with torch.no_grad():
                preds, emb_ens = self(mode, dropedge=0.0, **args)
                loss = self._loss(preds, label_targets)
        assert preds.shape[0] == label_targets.shape[0]
        return {"batch_size": preds.shape[0], "loss": loss, "labels": label_targets, "preds": self.predict(preds), "emb_ens": emb_ens}
