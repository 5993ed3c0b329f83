"""Secondary measurement (bench.py is the contract): BASELINE configs[0] / configs[1] on the S-arxiv stand-in.
  C1  k-hop(depth 2, budget 10) + hops sampler, super-batches through the generic kernel (sample_induce_kernel), Philox stream and glibc replay
  C2  3-layer GCN-256 + hops augmentation on those batches, batch 32, whole-step CUDA graph: train samples/s
Prints one JSON line.  python scripts/bench_c1c2.py [steps]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch


def main():
    import shadow_gnn_b200.ParallelSampler as PS
    from shadow_gnn_b200 import minibatch as MB
    from shadow_gnn_b200.models import DeepGNN
    from shadow_gnn_b200.synth import PRESETS, powerlaw_graph_torch
    from shadow_gnn_b200.train import GraphedTrainer
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    N, nnz, dmax, F, C, ntrain, seed = PRESETS["S-arxiv"]
    dev = torch.device("cuda:0")
    ip64, ix = powerlaw_graph_torch(N, nnz, seed, dmax, dev)
    ip = torch.where(ip64 >= 2 ** 31, ip64 - 2 ** 32, ip64).to(torch.int32)
    targets = np.random.default_rng(seed).permutation(N)[:ntrain].astype(np.uint32)
    cfg = dict(method="khop", depth="2", budget="10", num_roots="1", add_self_edge="true", include_target_conn="false")
    out = {"workload": "S-arxiv (169,343 nodes, 2.3 M edges): k-hop(2, 10) + hops sampler; 3-layer GCN-256 + hops, batch 32", "data": "synthetic"}
    for rng, P in (("philox", 16384), ("glibc", 4096)):
        s = PS.ParallelSampler.from_device_csr(ip, ix, P, seed=1, num_ring=2, rng=rng)
        s.shuffle_targets(targets)
        for _ in range(3):
            b = s.sample_to_device([cfg], [{"hops"}])[0]
        torch.cuda.synchronize()
        t0, n = time.perf_counter(), 0
        for _ in range(10):
            b = s.sample_to_device([cfg], [{"hops"}])[0]
            n += b.num_subg
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out[f"c1_sampler_{rng}"] = {"value": n / dt, "unit": "subgraphs/s", "roots_per_call": P, "avg_nodes": b.total_nodes / b.num_subg, "avg_edges": b.total_edges / b.num_subg}
    feat = torch.randn(N, F, device=dev)
    labels = torch.randint(0, C, (N,), device=dev)
    B = 32
    mcfg = {"batch_size": B, "configs": [{"method": "khop", "depth": [2], "budget": [10], "add_self_edge": [True]}]}
    mb = MB.MinibatchShallowExtractor("S-arxiv", None, {m: (ip, ix) for m in range(3)}, {0: targets.astype(np.int64), 1: targets[:B].astype(np.int64), 2: targets[:B].astype(np.int64)},
                                      mcfg, {"hops"}, None, feat, labels, F, True, 1, seed_cpp=1, num_subg_per_batch=B * 64, rng="philox")
    arch = dict(num_layers=3, num_cls_layers=1, heads=1, branch_sharing=False, dim=256, act="relu", layer_norm="norm_feat", feature_augment_ops="sum",
                aggr="gcn", residue="none", pooling="center", loss="softmax", ensemble_act="leakyrelu")
    model = DeepGNN(F, F, C, 0, arch, [("hops", mb.get_aug_dim("hops"))], 1, dict(dropout=0.1, dropedge=0.1, lr=0.001, ensemble_dropout="none"), "node").to(dev)
    mb.epoch_start_reset(0, MB.TRAIN); mb.shuffle_entity(MB.TRAIN)
    tr = GraphedTrainer(model, mb, row_cap=B * 111, edge_cap=B * 111 * 40)
    for _ in range(20):
        tr.step()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        tr.step()
    ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    out["c2_train"] = {"value": steps * B / (ms * 1e-3), "unit": "samples/s", "ms_per_step": ms / steps, "graph_steps": tr.graph_steps, "eager_steps": tr.eager_steps,
                       "params": sum(p.numel() for p in model.parameters())}
    print(json.dumps(out), flush=True)
    os._exit(0)


if __name__ == "__main__":
    main()
