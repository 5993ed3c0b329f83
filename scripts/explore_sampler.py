"""Exploratory timing of the sampler on the S-products stand-in (scratch; bench.py is the contract)."""
import sys, time, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import shadow_gnn_b200.ParallelSampler as PS
from shadow_gnn_b200.synth import powerlaw_graph_torch, PRESETS

name = sys.argv[1] if len(sys.argv) > 1 else "S-products"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
N, nnz, dmax, F, Cc, ntrain, seed = PRESETS[name]
dev = torch.device("cuda:0")
t0 = time.time()
indptr64, indices = powerlaw_graph_torch(N, nnz, seed, dmax, dev)
indptr = indptr64.to(torch.int32)
torch.cuda.synchronize()
print(f"graph {name}: N={N} nnz={indices.numel()} dmax={int(torch.diff(indptr64).max())} built in {time.time()-t0:.1f}s", flush=True)
feat = torch.randn(N, F, device=dev)
s = PS.ParallelSampler.from_device_csr(indptr, indices, 4096, seed=1)
rng = np.random.default_rng(seed)
targets = rng.permutation(N)[:T].astype(np.uint32)
t0 = time.time()
s.preproc_ppr_approximate(targets, 150, 0.85, 1e-5, "", "")
print(f"GPU ppr push: {T} targets in {time.time()-t0:.2f}s -> {T/(time.time()-t0):.0f} targets/s", flush=True)
cfg = dict(method="ppr", k="150", threshold="0", num_roots="1", add_self_edge="false", include_target_conn="false")
deg = torch.diff(indptr64)
for P in (512, 4096, 16384):
    if P > T: continue
    s.set_num_sampler_per_batch(P)
    s.shuffle_targets(targets[:P])
    b = s.sample_to_device([cfg], [set()])[0]   # warm-up + capacity growth
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    times = []
    for it in range(5):
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev).fill_(1)  # flush L2
        torch.cuda.synchronize()
        ev[0].record(); s._launch([cfg], [set()]); ev[1].record(); torch.cuda.synchronize()
        times.append(ev[0].elapsed_time(ev[1]))
    b = PS.DeviceBatch(s, 0)
    on = b.orig_node.long() & 0xFFFFFFFF
    nV, nE = b.total_nodes, b.total_edges
    alg = int((8 + 4 * (deg[on] + 1)).sum()) + 4 * nV + 4 * (nV + P) + 4 * nV + 8 * nE + 4 * P + 4 * nV + 8 * nV
    tg = []
    for it in range(5):
        torch.cuda.synchronize(); ev[0].record(); x = PS.gather_rows(feat, b.orig_node); ev[1].record(); torch.cuda.synchronize()
        tg.append(ev[0].elapsed_time(ev[1]))
    t = min(times); g = min(tg)
    print(json.dumps(dict(P=P, ms=round(t, 4), all_ms=[round(x, 3) for x in times], subg_per_s=round(P / t * 1e3), nodes=nV, edges=nE,
                          avg_deg_scoped=float(deg[on].float().mean()), alg_MB=round(alg / 1e6, 2), GBps=round(alg / t / 1e6, 1),
                          gather_ms=round(g, 4), gather_GBps=round(2 * 4 * F * nV / g / 1e6, 1))), flush=True)
