import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import shadow_gnn_b200.ParallelSampler as PS
from shadow_gnn_b200.synth import powerlaw_graph_torch, PRESETS
N, nnz, dmax, F, Cc, ntrain, seed = PRESETS["S-products"]
dev = torch.device("cuda:0")
indptr64, indices = powerlaw_graph_torch(N, nnz, seed, dmax, dev)
s = PS.ParallelSampler.from_device_csr(indptr64.to(torch.int32), indices, 4096, seed=1)
t = np.random.default_rng(seed).permutation(N)[:4096].astype(np.uint32)
s.preproc_ppr_approximate(t, 150, 0.85, 1e-5, "", "")
s.shuffle_targets(t)
b = s.sample_to_device([dict(method="ppr", k="150", threshold="0", num_roots="1", add_self_edge="false", include_target_conn="false")], [set()])[0]
deg = torch.diff(indptr64)[b.orig_node.long() & 0xFFFFFFFF].float()
tot = deg.sum()
print("rows", deg.numel(), "sum deg", int(tot), "mean", float(deg.mean()), "median", float(deg.median()))
for th in (32, 64, 128, 256, 512, 1024, 2048, 4096, 8192):
    m = deg > th
    print(f"deg>{th}: rows {float(m.float().mean())*100:5.2f}%  slots {float(deg[m].sum()/tot)*100:5.2f}%")
items32 = torch.ceil(deg / 32).sum(); items128 = torch.ceil(deg / 128).sum()
print("items32 per subg", float(items32) / 4096, "items128 per subg", float(items128) / 4096, "lane util 32:", float(tot / (items32 * 32)), "128:", float(tot / (items128 * 128)))
