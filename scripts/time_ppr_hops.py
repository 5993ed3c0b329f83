"""Timing (scratch; bench.py is the contract): single-root PPR + hop labels on the S-products stand-in, fast path vs generic kernel."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import shadow_gnn_b200.ParallelSampler as PS
from shadow_gnn_b200.synth import powerlaw_graph_torch, PRESETS
P = 65536
N, nnz, dmax, F, Cc, ntrain, seed = PRESETS["S-products"]
dev = torch.device("cuda:0")
ip64, ix = powerlaw_graph_torch(N, nnz, seed, dmax, dev)
ip = ip64.to(torch.int32)
targets = np.random.default_rng(seed).permutation(N)[:P].astype(np.uint32)
cfg = dict(method="ppr", k="150", threshold="0", num_roots="1", add_self_edge="false", include_target_conn="false")
out = {}
for name, env in (("fast_path", {}), ("generic_kernel", {"SHADOW_NO_WARP_PPR": "1"})):
    os.environ.update(env)
    s = PS.ParallelSampler.from_device_csr(ip, ix, P, seed=1)
    s.preproc_ppr_approximate(targets, 150, 0.85, 1e-5, "", "")
    s.shuffle_targets(targets)
    for aug in (set(), {"hops"}):
        for _ in range(3):
            b = s.sample_to_device([cfg], [aug])[0]
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(5):
            s._launch([cfg], [aug])
        ev[1].record(); torch.cuda.synchronize()
        out[f"{name}{'+hops' if aug else ''}"] = {"ms_per_65536": ev[0].elapsed_time(ev[1]) / 5, "sym": s.last_sym()}
    for k in env: os.environ.pop(k)
print(json.dumps(out))
os._exit(0)
