"""Whole-step CUDA graph for the training loop of the reference (`one_batch` + `DeepGNN.step`, shaDow/main.py:151-164).

At batch 32 the step is ~150 small kernels over an L2-resident working set: launch-bound.  The step is therefore captured ONCE
at a fixed capacity (rows, edges) and replayed: every batch is copied (3 small device copies) from the sampler's super-batch into
static buffers whose padding rows have empty adjacency rows, so they never reach a real row, the loss, or any gradient.
Dropout uses torch's graph-safe Philox state; dropedge reads its stream position and its draw count on the device
(csrc/layers.cu: dropedge_kernel), so every replay drops different edges.  Batches that do not fit the captured capacity (the short
batch at the end of an epoch, an unusually large scope) run through the eager `DeepGNN.step`.
"""
import os

import numpy as np
import torch
import torch.nn.functional as F

from .minibatch import TRAIN
from ._lib import check, lib
from .ops import DeviceCSR, _p, _stream
from .parallel import allreduce_flat_gradients, allreduce_bucket, join_buckets


class GraphedTrainer:
    def __init__(self, model, minibatch, row_cap, edge_cap, mode=TRAIN):
        assert minibatch.prediction_task == "node", "link prediction (two targets per subgraph) runs through the eager DeepGNN.step"
        pools = {rp.type_pool for rp in model.res_pool_layers}
        assert "sort" not in pools, "sort pooling has data-dependent shapes (repeat_interleave): use the eager DeepGNN.step"
        self._needs_sizes = pools != {"center"}
        self.model, self.mb, self.mode = model, minibatch, mode
        minibatch.prefetch_canonical = True
        self.B = minibatch._cfg_ensemble["batch_size"]
        dev, Fd = minibatch.dev_torch, minibatch.feat_full.shape[1]
        self.row_cap, self.edge_cap = int(row_cap), int(edge_cap)
        # one set of static buffers per ensemble branch (subgraph ensemble, shaDow/layers.py:236-296); feature-augmentation one-hots
        # (hops / pprs / drnls, shaDow/minibatch.py:469-477) get static buffers of their own
        self.E = E = minibatch.num_ensemble
        self.rowptr = [torch.zeros(self.row_cap + 1, dtype=torch.int32, device=dev) for _ in range(E)]
        self.span = [torch.zeros((self.row_cap, 2), dtype=torch.int32, device=dev) for _ in range(E)]
        self.col = [torch.zeros(self.edge_cap, dtype=torch.int32, device=dev) for _ in range(E)]
        self.val = [torch.zeros(self.edge_cap, dtype=torch.float32, device=dev) for _ in range(E)]
        self.feat = [torch.zeros((self.row_cap, Fd), dtype=torch.float32, device=dev) for _ in range(E)]
        self.target = [torch.zeros(self.B, dtype=torch.int64, device=dev) for _ in range(E)]
        self.aug = [{k: torch.zeros((self.row_cap, minibatch.get_aug_dim(k)), dtype=torch.float32, device=dev)
                     for k in sorted(set(minibatch.aug_feats) & {"hops", "pprs", "drnls"})} for _ in range(E)]
        lf = minibatch.label_full
        self.label = torch.zeros((self.B,) + tuple(lf.shape[1:]), dtype=lf.dtype, device=dev)      # class ids [B] or multi-hot rows [B, C]
        # rows per subgraph of the loaded batch (max / mean / sum pooling segments; padding rows lie behind the last segment)
        self.sizes = torch.ones((E, self.B), dtype=torch.int64, device=dev)
        self.loss = torch.zeros((), dtype=torch.float32, device=dev)
        self.keep_preds = False                              # step_logged(): the captured step also leaves predict(preds) in a static buffer
        self.preds = None
        self.graph = None
        self.world = 1
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            self.world = torch.distributed.get_world_size()
        self.graph_bwd = None
        self.eager_steps = self.graph_steps = 0
        # data parallel: the gradient exchange and the optimizer live INSIDE the captured step.  The flat gradient buffer is cut in two
        # at the first parameter of conv layer `_split_layer`: the tail bucket (that layer and everything behind it: later layers, pooling,
        # classifier) is all-reduced on a side stream as soon as the layer's input gradient exists, i.e. while the earlier layers are
        # still in their backward pass; only the head bucket's exchange is exposed.  SHADOW_DP_GRAPH=0 keeps the exchange outside the graph.
        self.dp_in_graph = self.world > 1 and os.environ.get("SHADOW_DP_GRAPH", "1") != "0"
        self.p2p = False                                     # set at capture time: the optimizer exchanges over peer memory, no NCCL buckets
        self._side = None
        self._split = 0                                      # first float of the lowest side-stream bucket (0: one exchange after the backward pass)
        self._buckets = []                                   # [(layer whose output gradient fires the exchange, lo, hi)], hi = None: to the end

    # ------------------------------------------------------------------
    def _fwd_bwd(self, exchange=False):
        m = self.model
        adjs = [DeviceCSR(self.span[i], self.col[i], 0, self.val[i], row_ord=self.rowptr[i]) for i in range(self.E)]
        handles = []
        if exchange and self._split > 0:
            opt = m.optimizer
            for prev, lo, hi in self._buckets:
                def fire(module, inputs, output, lo=lo, hi=hi):      # this layer's OUTPUT gradient exists <=> every later layer's backward is done
                    output[0].register_hook(lambda g: allreduce_bucket(opt.grad, lo, hi, self._side))
                handles.append(prev.register_forward_hook(fire))
        if self.mode != TRAIN:                               # evaluation epochs (shaDow/main.py:186-187): forward + loss + predictions only
            with torch.no_grad():
                preds, _ = m(self.mode, list(self.feat), adjs, list(self.target), self.sizes, self.aug, 0.0)
                self.loss.copy_(m._loss(preds, self.label))
                p = m.predict(preds)
                if self.preds is None:
                    self.preds = torch.zeros_like(p)
                self.preds.copy_(p)
            return
        preds, _ = m(self.mode, list(self.feat), adjs, list(self.target), self.sizes, self.aug, m.dropedge)
        for h in handles:
            h.remove()
        loss = m._loss(preds, self.label)
        if self.keep_preds:
            with torch.no_grad():
                p = m.predict(preds.detach())
                if self.preds is None:
                    self.preds = torch.zeros_like(p)
                self.preds.copy_(p)
        loss.backward()
        if exchange and not self.p2p:
            opt = m.optimizer
            if self._split > 0:
                join_buckets(opt.grad, self._split, self._side)
            else:
                allreduce_flat_gradients(opt.grad)
        self.loss.copy_(loss.detach())

    def _plan_buckets(self, opt):
        """Three buckets over the layer-ordered flat gradient buffer.  A = conv layers [L - L//2 ..) + pooling + classifier: exchanged on a
        side stream as soon as the first of those layers has finished its backward pass (about half of the pass is still to come); B = the
        layers in between: exchanged when layer 1 is done, hidden behind layer 0's backward pass; C = layer 0 (and whatever precedes it):
        the only exchange that is exposed, on the current stream after the pass (~0.2 MB: latency only)."""
        convs = list(self.model.conv_layers[0])
        if len(convs) < 2 or self.E > 1:                     # several branches run their backward passes one after the other: one exchange at the end
            return
        off = lambda layer: opt.offsets.get(id(next(iter(layer.parameters()), None)))
        k = len(convs) - max(1, len(convs) // 2)             # first layer of bucket A
        if off(convs[k]) is None:
            return
        self._buckets = [(convs[k - 1], off(convs[k]), None)]
        self._split = off(convs[k])
        if k >= 2 and off(convs[1]) is not None and 0 < off(convs[1]) < off(convs[k]):
            self._buckets.append((convs[0], off(convs[1]), off(convs[k])))
            self._split = off(convs[1])
        self._side = torch.cuda.Stream()

    def _capture_eval(self):
        m = self.model
        m.eval()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                self._fwd_bwd()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self._fwd_bwd()

    def _capture(self):
        m = self.model
        if self.mode != TRAIN:
            return self._capture_eval()
        opt = m._ensure_optimizer()
        m.train()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                       # warm-up off the capture stream (cuBLAS workspaces, autograd buffers)
            for _ in range(3):
                opt.zero_grad()
                self._fwd_bwd()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        opt.zero_grad()
        self.p2p = opt.p2p is not None
        if self.p2p:
            self.dp_in_graph = True                          # the exchange is part of opt.step(): always inside the captured step
        elif self.dp_in_graph:
            self._plan_buckets(opt)
            torch.distributed.all_reduce(torch.zeros(8, device=opt.grad.device))      # the communicator exists before the capture starts
            torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        # thread_local: NCCL's watchdog thread may touch the CUDA API while this thread captures
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            opt.zero_grad()
            self._fwd_bwd(exchange=self.dp_in_graph)
            if self.world == 1 or self.dp_in_graph:
                opt.step(1.0 / self.world)
        # SHADOW_DP_GRAPH=0: the gradient all-reduce sits between the captured fwd/bwd and the (eager, 4-launch) optimizer step

    def _load_static(self, sb, bs, i=0):
        a = sb.cursor
        rowptr, indices, lo, e0, feat, target = sb.take_canonical(bs)
        n, e = rowptr.numel() - 1, indices.numel()
        if n > self.row_cap or e > self.edge_cap:
            return False
        if self._needs_sizes:
            self.sizes[i].copy_(torch.from_numpy(np.diff(sb.node_ptr_host[a:a + bs + 1])))
        # one launch: CSR slice rebased to row 0 / edge 0, padding rows empty, feature rows copied, targets rebased (csrc/gather.cu)
        check(lib.shadow_load_batch(_p(rowptr), e0, n, e, self.row_cap, _p(self.rowptr[i]), _p(self.span[i]), _p(indices) if e else None, lo, _p(self.col[i]),
                                    _p(feat), feat.shape[1], _p(self.feat[i]), _p(target), target.numel(), _p(self.target[i]), _stream(feat)))
        for k, buf in self.aug[i].items():                   # padding rows keep stale one-hots: nothing reads their outputs
            buf[:n].copy_(sb.aug[k][lo:lo + n])
        return True

    def step_logged(self):
        """`step` for the trainer shell (main.one_epoch): the dict `DeepGNN.step` returns (models.py:209-237 of the reference) -- loss, labels
        and predictions are device tensors (copies of the captured step's static buffers), nothing synchronises"""
        if not self.keep_preds and self.mode == TRAIN:
            self.keep_preds, self.graph = True, None         # (re)capture with the prediction buffer
        out = self.step(_full=True)
        if isinstance(out, dict):                            # eager fallback batch
            return out
        lab = self.label
        if lab.dim() == 1 and self.model.num_classes > 1:
            lab = F.one_hot(lab.to(torch.int64), num_classes=self.model.num_classes)
        return {"batch_size": self.B, "loss": self.loss.clone(), "labels": lab.clone(), "preds": self.preds.clone(), "emb_ens": None}

    def step(self, _full=False):
        """one training step on the next batch of the epoch; returns the loss (a device scalar, no sync)"""
        mb, mode = self.mb, self.mode
        bs = mb._get_cur_batch_size(mode)
        a = mb.idx_entity_evaluated[mode]
        label = mb.label_epoch[mode][a:a + bs]
        sbs = [mb._front(mode, i, bs) for i in range(self.E)]
        if bs == self.B:
            cursors = [sb.cursor for sb in sbs]
            if all(self._load_static(sb, bs, i) for i, sb in enumerate(sbs)):
                self.label.copy_(label)
                mb._update_batch_stat(mode, bs)
                if self.graph is None:
                    self._capture()
                self.graph.replay()
                if self.mode == TRAIN and self.world > 1 and not self.dp_in_graph:
                    opt = self.model.optimizer
                    opt.step(allreduce_flat_gradients(opt.grad))
                self.graph_steps += 1
                return self.loss
            for sb, c in zip(sbs, cursors):                  # did not fit: hand the batch to the eager path
                sb.cursor = c
        self.eager_steps += 1
        out = self.model.step(mode, "running", mb.one_batch(mode))
        if mode != TRAIN and self.graph is not None:
            self.model.eval()
        return out if _full else out["loss"].detach()
