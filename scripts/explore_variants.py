"""Exploration (scratch; bench.py is the contract): time the single-root PPR sampler launch on the S-products stand-in for
every kernel-variant build under shadow_gnn_b200/variants/ (scripts/build_variants.sh) x a few run-time knobs.
One subprocess per library build (SHADOW_B200_LIB), graph cached in /dev/shm between them."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CACHE = "/dev/shm/sprod_graph.pt"
P = int(os.environ.get("EXPLORE_P", "16384"))


def worker():
    import numpy as np
    import torch
    import shadow_gnn_b200.ParallelSampler as PS
    from shadow_gnn_b200.synth import powerlaw_graph_torch, PRESETS
    N, nnz, dmax, F, Cc, ntrain, seed = PRESETS["S-products"]
    dev = torch.device("cuda:0")
    if os.path.exists(CACHE):
        indptr, indices = torch.load(CACHE)
        indptr, indices = indptr.to(dev), indices.to(dev)
    else:
        indptr64, indices = powerlaw_graph_torch(N, nnz, seed, dmax, dev)
        indptr = indptr64.to(torch.int32)
        torch.save((indptr.cpu(), indices.cpu()), CACHE)
    s = PS.ParallelSampler.from_device_csr(indptr, indices, P, seed=1)
    targets = np.random.default_rng(seed).permutation(N)[:P].astype(np.uint32)
    s.preproc_ppr_approximate(targets, 150, 0.85, 1e-5, "", "")
    s.shuffle_targets(targets)
    cfg = dict(method="ppr", k="150", threshold="0", num_roots="1", add_self_edge="false", include_target_conn="false")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    knobs = json.loads(os.environ.get("EXPLORE_KNOBS", "[{}]"))
    for kn in knobs:
        for k, v in kn.items():
            os.environ[k] = str(v)
        try:
            b = s.sample_to_device([cfg], [set()])[0]      # warm-up + capacity growth
            redo = s.last_redo_count()
            times = []
            for _ in range(6):
                flush.fill_(1)
                torch.cuda.synchronize()
                ev[0].record(); s._launch([cfg], [set()]); ev[1].record(); torch.cuda.synchronize()
                times.append(ev[0].elapsed_time(ev[1]))
            b = PS.DeviceBatch(s, 0)
            print(json.dumps(dict(lib=os.path.basename(os.environ.get("SHADOW_B200_LIB", "default")), P=P, knobs=kn, ms=round(min(times), 4), us_per_kilo_subg=round(min(times) * 1e6 / P, 2),
                                  all=[round(x, 3) for x in times], redo=redo, nodes=b.total_nodes, edges=b.total_edges)), flush=True)
        except Exception as e:      # noqa
            print(json.dumps(dict(lib=os.path.basename(os.environ.get("SHADOW_B200_LIB", "default")), knobs=kn, error=str(e)[:300])), flush=True)
        for k in kn:
            os.environ.pop(k, None)
    os._exit(0)


def main():
    vdir = os.path.join(ROOT, "shadow_gnn_b200", "variants")
    libs = [None] + sorted(os.path.join(vdir, f) for f in os.listdir(vdir) if f.endswith(".so")) if os.path.isdir(vdir) else [None]
    knobs_default = json.loads(os.environ.get("EXPLORE_KNOBS_ALL", '[{}, {"SHADOW_WARP_NF": 1}]'))
    jobs = [(None, 65536, knobs_default)] + [(lib, 65536, knobs_default) for lib in libs if lib]
    for lib, P_, kn in jobs:
        env = dict(os.environ)
        env["EXPLORE_WORKER"] = "1"
        env["EXPLORE_P"] = str(P_)
        if lib:
            env["SHADOW_B200_LIB"] = lib
        env["EXPLORE_KNOBS"] = json.dumps(kn)
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, capture_output=True, text=True, timeout=240)
            out = r.stdout.strip()
            print(out if out else f"# {lib}: no output; stderr tail: {r.stderr[-400:]}", flush=True)
        except subprocess.TimeoutExpired:
            print(f"# {lib}: TIMEOUT", flush=True)
        print(f"# {os.path.basename(lib) if lib else 'default'} P={P_} took {time.time() - t0:.1f}s", flush=True)


if __name__ == "__main__":
    worker() if os.environ.get("EXPLORE_WORKER") else main()
