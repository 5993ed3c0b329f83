import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import shadow_gnn_b200.ParallelSampler as PS
from shadow_gnn_b200.synth import powerlaw_graph_torch, PRESETS
N, nnz, dmax, F, Cc, ntrain, seed = PRESETS["S-products"]
dev = torch.device("cuda:0")
indptr64, indices = powerlaw_graph_torch(N, nnz, seed, dmax, dev)
indptr = indptr64.to(torch.int32)
s = PS.ParallelSampler.from_device_csr(indptr, indices, 4096, seed=1)
perm = np.random.default_rng(seed).permutation(N)
for lo, hi in ((0, 16384), (16384, 32768), (32768, 65536), (65536, 131072), (131072, 196615)):
    t0 = time.time()
    s.preproc_ppr_approximate(perm[lo:hi].astype(np.uint32), 150, 0.85, 1e-5, "", "")
    print(lo, hi, "ok", round(time.time() - t0, 2), flush=True)
