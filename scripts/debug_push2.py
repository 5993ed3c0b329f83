import sys, os, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
if len(sys.argv) > 1:
    import torch
    import shadow_gnn_b200.ParallelSampler as PS
    from shadow_gnn_b200.synth import powerlaw_graph_torch, PRESETS
    lo, hi = int(sys.argv[1]), int(sys.argv[2])
    N, nnz, dmax, F, Cc, ntrain, seed = PRESETS["S-products"]
    dev = torch.device("cuda:0")
    indptr64, indices = powerlaw_graph_torch(N, nnz, seed, dmax, dev)
    indptr = indptr64.to(torch.int32)
    perm = np.random.default_rng(seed).permutation(N)
    t = perm[lo:hi]
    deg = torch.diff(indptr64)
    if hi - lo <= 8:
        for v in t:
            nb = indices[indptr64[v]:indptr64[v + 1]].long()
            print("target", v, "deg", int(deg[v]), "nbr degs", deg[nb].tolist()[:20], "nbrs", nb.tolist()[:20], flush=True)
            for u in nb.tolist()[:3]:
                nb2 = indices[indptr64[u]:indptr64[u + 1]].long()
                print("   nbr", u, "deg", int(deg[u]), "->", nb2.tolist()[:10], deg[nb2].tolist()[:10], flush=True)
    s = PS.ParallelSampler.from_device_csr(indptr, indices, 4096, seed=1)
    t0 = time.time()
    s.preproc_ppr_approximate(t.astype(np.uint32), 150, 0.85, 1e-5, "", "")
    print(lo, hi, "ok", round(time.time() - t0, 2), flush=True)
    os._exit(0)
lo, hi = 32768, 65536
while hi - lo > 1:
    mid = (lo + hi) // 2
    try:
        subprocess.run([sys.executable, __file__, str(lo), str(mid)], timeout=25, check=True, capture_output=True)
        lo = mid            # first half fine -> culprit in second half
    except subprocess.TimeoutExpired:
        hi = mid
    print("range", lo, hi, flush=True)
print(subprocess.run([sys.executable, __file__, str(lo), str(hi)], timeout=40, capture_output=True, text=True).stdout)
