"""ncu target: a few launches of the fused sampler kernel on the S-products stand-in (P subgraphs, PPR k=150)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import shadow_gnn_b200.ParallelSampler as PS
from shadow_gnn_b200.synth import powerlaw_graph_torch, PRESETS
P = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
N, nnz, dmax, F, Cc, ntrain, seed = PRESETS["S-products"]
dev = torch.device("cuda:0")
indptr64, indices = powerlaw_graph_torch(N, nnz, seed, dmax, dev)
indptr = indptr64.to(torch.int32)
s = PS.ParallelSampler.from_device_csr(indptr, indices, P, seed=1)
targets = np.random.default_rng(seed).permutation(N)[:P].astype(np.uint32)
s.preproc_ppr_approximate(targets, 150, 0.85, 1e-5, "", "")
cfg = dict(method="ppr", k="150", threshold="0", num_roots="1", add_self_edge="false", include_target_conn="false")
s.shuffle_targets(targets)
for _ in range(4):
    b = s.sample_to_device([cfg], [set()])[0]
torch.cuda.synchronize()
print("done", b.total_nodes, b.total_edges)
os._exit(0)
