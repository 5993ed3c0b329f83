#!/usr/bin/env python
"""bench.py -- the driver's measurement contract (see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA path through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's own CPU implementation of the path

Workload (BASELINE.json configs[2], the one the metric is quoted on): ogbn-products stand-in S-products
(N=2,449,029, nnz~123.7 M, F=100; SURVEY.md 8d -- no network, so synthetic), PPR sampler k=150 eps=1e-5
threshold 0, targets = a permutation of 196,615 train nodes, sampler seed 1.
  --task sampler : a step = one super-batch of `--superbatch` roots: select + induce + feature gather, result
                   resident in HBM as one block-diagonal batch.  metric = sampler_subgraphs_per_sec.
  --task train   : a step = one training batch of 32 targets through sample -> 5-layer SAGE fwd/bwd -> Adam.
                   metric = train_samples_per_sec.   (enabled once shadow_gnn_b200.models exists)
One JSON line on stdout (rank 0).  Timed with CUDA events on the launching stream, barrier + synchronize on both
sides, max over ranks.  Inputs (495 MB of CSR indices, 980 MB of features) are larger than the 126 MB L2 and every
step samples different roots, so no step re-reads what the previous one left in L2.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PPR_K, PPR_ALPHA, PPR_EPS = 150, 0.85, 1e-5
SAMPLER_CFG = dict(method="ppr", k=str(PPR_K), threshold="0", num_roots="1", add_self_edge="false", include_target_conn="false")


T0 = time.time()


def log(*a):
    print(f"[bench {time.time() - T0:7.2f}s]", *a, file=sys.stderr, flush=True)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--task", default=None, choices=["sampler", "train"])
    ap.add_argument("--superbatch", type=int, default=16384)
    ap.add_argument("--superbatch-train", type=int, default=4096)
    ap.add_argument("--graph", default="S-products")
    ap.add_argument("--cpu-sample", type=int, default=4000, help="roots of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip().split(", "))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_graph(name, device):
    """S-products etc. on the GPU (setup, untimed); returns int32-viewed CUDA tensors + int64 degrees."""
    import torch
    from shadow_gnn_b200.synth import PRESETS, powerlaw_graph_torch
    N, nnz, dmax, F, C, ntrain, seed = PRESETS[name]
    indptr64, indices = powerlaw_graph_torch(N, nnz, seed, dmax, device)
    train = torch.from_numpy(np.random.default_rng(seed).permutation(N)[:ntrain].astype(np.int64))
    return dict(N=N, F=F, C=C, seed=seed, indptr64=indptr64, indptr=indptr64.to(torch.int32), indices=indices, train=train)


def algorithmic_bytes(batch, deg, F, P):
    """bytes each unit of work must move once (4-byte ids / floats, no cache credit)."""
    on = batch.orig_node.long() & 0xFFFFFFFF
    nV, nE = batch.total_nodes, batch.total_edges
    b_induce = int((8 + 4 * (deg[on] + 1)).sum()) + 4 * nV + 4 * (nV + P) + 4 * nV + 8 * nE + 4 * P + 4 * nV   # incl. the ppr column
    b_ppr = 8 * nV
    b_gather = 2 * 4 * F * nV
    return b_induce + b_ppr, b_gather


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own sampler (oracle/_ref = unmodified ParallelSampler.cpp) or, failing that, the oracle port
# ------------------------------------------------------------------------------------------------
def cpu_sampler_arm(g_host, roots, steps, warmup, threads):
    log('cpu arm: start', roots.size, 'roots', threads, 'threads')
    """Times parallel_sampler_ensemble (500 roots per call, minibatch.py:397) + what the reference pays to turn the result
    into the same product we emit: list->numpy conversion (samplers_ensemble.py:254-265), cat_to_block_diagonal
    (graph.py:280-320) and the feature gather (minibatch.py:469).  Returns (subgraphs/s full product, subgraphs/s C++ call only, kind)."""
    from oracle import oracle as O
    indptr, indices, feat = g_host["indptr"], g_host["indices"], g_host["feat"]
    ref = O.load_ref()
    per_call = 500
    if ref is not None:
        d = tempfile.mkdtemp()
        fi, fx = os.path.join(d, "indptr.bin"), os.path.join(d, "indices.bin")
        indptr.tofile(fi); indices.tofile(fx)
        devnull = os.open(os.devnull, os.O_WRONLY); saved = os.dup(1); os.dup2(devnull, 1)     # the reference prints from C++
        try:
            s = ref.ParallelSampler([], [], [], per_call, threads, True, True, [], 1, fi, fx, "", 1)
            s.preproc_ppr_approximate(roots.tolist(), PPR_K, PPR_ALPHA, PPR_EPS, "", "")
        finally:
            os.dup2(saved, 1); os.close(devnull)
        log('cpu arm: reference sampler + PPR tables ready')
        s.shuffle_targets(roots.tolist())
        kind = "reference"

        def one_call():
            t0 = time.perf_counter()
            vec = s.parallel_sampler_ensemble([SAMPLER_CFG], [set()])[0]
            t1 = time.perf_counter()
            sub = O.ref_subgraphs(vec)
            return sub, t1 - t0
    else:
        s = O.OracleSampler(indptr, indices, per_call, threads, 1)
        s.preproc_ppr_approximate(roots, PPR_K, PPR_ALPHA, PPR_EPS)
        s.shuffle_targets(roots)
        cfg = O.cfg_from_cpp_config(SAMPLER_CFG)
        kind = "port"

        def one_call():
            t0 = time.perf_counter()
            b = s.sample(cfg)
            t1 = time.perf_counter()
            return b.subgraphs(), t1 - t0
    n_sub, t_full, t_cpp = 0, 0.0, 0.0
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        sub, tc = one_call()
        col = O.cat_to_block_diagonal(sub)
        x = feat[col["node"]]
        t1 = time.perf_counter()
        if it >= warmup:
            n_sub += len(sub); t_full += t1 - t0; t_cpp += tc
        del x
    return n_sub / t_full, n_sub / t_cpp, kind, n_sub, t_full


def host_graph(args):
    """the same synthetic graph on the host (for the CPU arm)"""
    import torch
    from shadow_gnn_b200.synth import PRESETS, powerlaw_graph
    N, nnz, dmax, F, C, ntrain, seed = PRESETS[args.graph]
    if torch.cuda.is_available():
        g = build_graph(args.graph, torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))))
        indptr = g["indptr64"].cpu().numpy().astype(np.uint32)
        indices = g["indices"].cpu().numpy().view(np.uint32)
        train = g["train"].numpy()
        del g
        torch.cuda.empty_cache()
    else:
        indptr, indices = powerlaw_graph(N, nnz, seed, dmax)
        train = np.random.default_rng(seed).permutation(N)[:ntrain]
    feat = np.random.default_rng(seed + 100).standard_normal((N, F), dtype=np.float32)
    return dict(indptr=indptr, indices=indices, feat=feat, train=train.astype(np.uint32), N=N, F=F)


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    gh = host_graph(args)
    per_step = 500
    need = per_step * (args.steps + args.warmup)
    roots = gh["train"][:min(need, gh["train"].size)]
    t0 = time.time()
    v_full, v_cpp, kind, n_sub, t_full = cpu_sampler_arm(gh, roots, args.steps, args.warmup, threads)
    line = {
        "impl": "reference", "metric": "sampler_subgraphs_per_sec", "value": v_full, "unit": "subgraphs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_full / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": f"{args.graph} PPR(k={PPR_K},eps={PPR_EPS}) sampler: select + node-induced CSR + block-diagonal collation + feature gather",
                   "step": f"one parallel_sampler_ensemble call of {per_step} roots (shaDow/minibatch.py:397) + list->numpy + cat_to_block_diagonal + feat[node]"},
        "cpu_baseline": {"value": v_full, "unit": "subgraphs/s", "cores": threads, "kind": kind,
                         "sample": f"{n_sub} roots of the same target order; C++ call only = {v_cpp:.0f} subgraphs/s", "cpp_call_only": v_cpp},
        "e2e": {"value": v_full, "unit": "subgraphs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "setup_s": time.time() - t0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import shadow_gnn_b200.ParallelSampler as PS
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    assert torch.cuda.is_available(), "bench.py (our arm) needs a GPU: there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    P = args.superbatch
    log('building graph')
    g = build_graph(args.graph, dev)
    log('graph built')
    N, F = g["N"], g["F"]
    feat = torch.randn(N, F, device=dev, generator=torch.Generator(device=dev).manual_seed(g["seed"] + 100))
    deg = torch.diff(g["indptr64"])
    # every rank takes its contiguous share of the epoch's target order (independent units: no data-path collective)
    train = g["train"]
    share = train[rank::world].contiguous()
    roots_host = share.numpy().astype(np.uint32)
    s = PS.ParallelSampler.from_device_csr(g["indptr"], g["indices"], P, seed=1, num_ring=2)
    s.set_stream(torch.cuda.current_stream().cuda_stream)
    t0 = time.time()
    log('ppr push for', roots_host.size, 'targets')
    s.preproc_ppr_approximate(roots_host, PPR_K, PPR_ALPHA, PPR_EPS, "", "")        # GPU forward push (untimed setup)
    t_ppr = time.time() - t0
    log('ppr push done', t_ppr)
    roots_dev = torch.from_numpy(roots_host.view(np.int32)).to(dev)
    s.shuffle_targets_device(roots_dev)
    out_feat = [torch.empty((P * (PPR_K + 1), F), device=dev) for _ in range(2)]

    def step(i):
        s._launch([SAMPLER_CFG], [set()])
        b = PS.DeviceBatch(s, 0)          # syncs the stream: sizes are needed to launch the gather
        PS.gather_rows(feat, b.orig_node, out=out_feat[i % 2][:b.total_nodes])
        return b

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local) if rank == 0 else None
    for i in range(args.warmup):
        step(i)
    log('warmup done')
    # ---- timed region: value (inputs resident in HBM) ----
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    alg_sampler = alg_gather = 0
    nsub = 0
    barrier()
    ev0.record()
    batches = []
    for i in range(args.steps):
        kev[i][0].record()
        s._launch([SAMPLER_CFG], [set()])
        kev[i][1].record()
        b = PS.DeviceBatch(s, 0)
        PS.gather_rows(feat, b.orig_node, out=out_feat[i % 2][:b.total_nodes])
        nsub += b.num_subg
        batches.append((b.num_subg, b.total_nodes, b.total_edges))
        if i == args.steps - 1:
            last = b
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    log('timed region done', ms)
    k_ms = float(np.mean([a.elapsed_time(b_) for a, b_ in kev]))
    clk = clocks.stop() if clocks else None
    a1, a2 = algorithmic_bytes(last, deg, F, last.num_subg)      # last step's batch stands for the average step
    # ---- e2e: reference-facing call with HOST buffers: roots come from pinned host memory every step, every result
    #      array goes back to pinned host memory (what a drop-in `parallel_sampler_ensemble` caller receives) ----
    e2e_steps = max(3, min(args.steps // 4, 25))
    pinned_roots = torch.from_numpy(roots_host.view(np.int32)).pin_memory()
    h2d = d2h = 0
    host_out = {}
    barrier()
    t0 = time.perf_counter()
    nsub_e2e = 0
    for i in range(e2e_steps):
        lo = (i * P) % max(roots_host.size - P, 1)
        chunk = pinned_roots[lo:lo + P]
        rd = chunk.to(dev, non_blocking=True)
        h2d = chunk.numel() * 4
        s2 = s
        s2.shuffle_targets_device(rd)
        b = s2.sample_to_device([SAMPLER_CFG], [set()])[0]
        x = PS.gather_rows(feat, b.orig_node, out=out_feat[i % 2][:b.total_nodes])
        d2h = 0
        for name in ("node_ptr", "rowptr", "indices", "orig_node", "orig_edge", "target", "ppr"):
            t = getattr(b, name)
            if name not in host_out or host_out[name].numel() < t.numel():
                host_out[name] = torch.empty(int(t.numel() * 1.2) + 16, dtype=t.dtype).pin_memory()
            host_out[name][:t.numel()].copy_(t, non_blocking=True)
            d2h += t.numel() * 4
        torch.cuda.synchronize()
        nsub_e2e += b.num_subg
    barrier()
    e2e_s = time.perf_counter() - t0
    log('e2e done', e2e_s)
    s.shuffle_targets_device(roots_dev)
    # ---- reduce over ranks ----
    stats = torch.tensor([ms, e2e_s], device=dev, dtype=torch.float64)
    counts = torch.tensor([nsub, nsub_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    if rank == 0:
        peak, peak_src = load_peaks()
        ms_all, e2e_all = float(stats[0]), float(stats[1])
        value = float(counts[0]) / (ms_all * 1e-3)
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("sample_induce_kernel_dram_bytes_per_launch")
        line = {
            "metric": "sampler_subgraphs_per_sec", "value": value, "unit": "subgraphs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_all / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": f"{args.graph} PPR(k={PPR_K},eps={PPR_EPS}) sampler: select + node-induced CSR (bug-compatible) + block-diagonal batch + feature gather, resident in HBM",
                       "superbatch": P, "per_gpu_targets": int(roots_host.size), "parallelism": f"dp{world} (targets partitioned, graph/tables/features replicated)",
                       "l2": "inputs (495 MB CSR indices + 980 MB features) exceed the 126 MB L2; every step samples different roots",
                       "avg_nodes_per_subgraph": batches[-1][1] / batches[-1][0], "avg_edges_per_subgraph": batches[-1][2] / batches[-1][0],
                       "ppr_push_setup_s": t_ppr},
            "roofline": {"kernel": "sample_induce_kernel", "bound": "hbm", "achieved": (a1 / 1e9) / (k_ms * 1e-3), "peak": peak, "unit": "GB/s",
                         "frac": (a1 / 1e9) / (k_ms * 1e-3) / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": a1, "kernel_ms": k_ms,
                         "note": "algorithmic bytes = SURVEY.md 8(d) B_ppr + B_induce per subgraph x subgraphs per launch"},
            "roofline_gather": {"kernel": "gather_rows_vec4_kernel", "bound": "hbm", "algorithmic_bytes_per_launch": a2},
            "e2e": {"value": float(counts[1]) / e2e_all, "unit": "subgraphs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "note": "reference-facing boundary: roots come from pinned host memory each step; every array parallel_sampler_ensemble returns (CSR, node ids, edge ids, targets, ppr) is copied back to pinned host memory; gathered features stay in HBM as in shaDow/minibatch.py:469"},
            "gpu_launches": 2 * args.steps, "clocks": clk,
        }
        if not args.no_cpu_baseline and world == 1:
            try:
                gh = dict(indptr=g["indptr64"].cpu().numpy().astype(np.uint32), indices=g["indices"].cpu().numpy().view(np.uint32),
                          feat=feat.cpu().numpy())
                n_roots = min(args.cpu_sample, roots_host.size)
                threads = os.cpu_count() or 1
                v_full, v_cpp, kind, n_sub, t_full = cpu_sampler_arm(gh, roots_host[:n_roots], max(1, n_roots // 500 - 1), 1, threads)
                line["cpu_baseline"] = {"value": v_full, "unit": "subgraphs/s", "cores": threads, "kind": kind, "cpp_call_only": v_cpp,
                                        "sample": f"{n_sub} roots (500 per call as shaDow/minibatch.py:397), same graph/targets/PPR parameters; incl. list->numpy, block-diagonal collation, feature gather"}
            except Exception as e:      # the baseline is reported, never required
                line["cpu_baseline"] = {"value": None, "unit": "subgraphs/s", "cores": 0, "kind": "unavailable", "sample": repr(e)[:200]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# train task: sample -> 5-layer GraphSAGE forward/backward -> clip + Adam, batch of 32 targets per step
# ------------------------------------------------------------------------------------------------
TRAIN_CFG = dict(batch=32, layers=5, dim=256, dropout=0.4, dropedge=0.05, lr=0.002)          # config_train/products/vanilla/sage_5_ppr.yml
ARCH = dict(num_layers=5, num_cls_layers=1, heads=1, branch_sharing=False, dim=256, act="relu", layer_norm="norm_feat",
            feature_augment_ops="sum", aggr="sage", residue="none", pooling="center", loss="softmax", ensemble_act="leakyrelu")


def cpu_train_arm(g_host, roots, labels_all, steps, warmup, threads, C):
    """reference sampler (oracle/_ref, 500 roots per call) + collation + feature gather + the reference model's math on the host cores"""
    import torch
    from oracle import oracle as O
    from oracle.train_ref import RefModel
    torch.set_num_threads(threads)
    indptr, indices, feat = g_host["indptr"], g_host["indices"], torch.from_numpy(g_host["feat"])
    ref = O.load_ref()
    B = TRAIN_CFG["batch"]
    if ref is not None:
        d = tempfile.mkdtemp()
        fi, fx = os.path.join(d, "indptr.bin"), os.path.join(d, "indices.bin")
        indptr.tofile(fi); indices.tofile(fx)
        devnull = os.open(os.devnull, os.O_WRONLY); saved = os.dup(1); os.dup2(devnull, 1)
        try:
            s = ref.ParallelSampler([], [], [], 500, threads, True, True, [], 1, fi, fx, "", 1)
            s.preproc_ppr_approximate(roots.tolist(), PPR_K, PPR_ALPHA, PPR_EPS, "", "")
        finally:
            os.dup2(saved, 1); os.close(devnull)
        s.shuffle_targets(roots.tolist())
        kind = "reference"
        sample = lambda: O.ref_subgraphs(s.parallel_sampler_ensemble([SAMPLER_CFG], [set()])[0])
    else:
        s = O.OracleSampler(indptr, indices, 500, threads, 1)
        s.preproc_ppr_approximate(roots, PPR_K, PPR_ALPHA, PPR_EPS)
        s.shuffle_targets(roots)
        kind = "port"
        sample = lambda: s.sample(O.cfg_from_cpp_config(SAMPLER_CFG)).subgraphs()
    model = RefModel(feat.shape[1], TRAIN_CFG["dim"], C, TRAIN_CFG["layers"], TRAIN_CFG["dropout"], TRAIN_CFG["dropedge"], TRAIN_CFG["lr"])
    pool, done, t_total, n_samples = [], 0, 0.0, 0
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        while len(pool) < B:
            pool.extend(sample())                               # minibatch.py:450-451
        sub, pool = pool[:B], pool[B:]
        col = O.cat_to_block_diagonal(sub)                      # graph.py:280-320
        x = feat[torch.as_tensor(col["node"].astype(np.int64))]   # minibatch.py:469
        y = torch.as_tensor(labels_all[col["node"][col["target"]].astype(np.int64)])
        model.step(col, x, y)
        t1 = time.perf_counter()
        if it >= warmup:
            t_total += t1 - t0; n_samples += B
    return n_samples / t_total, kind, n_samples, t_total


def run_reference_train(args):
    if int(os.environ.get("RANK", 0)) != 0:
        return
    threads = os.cpu_count() or 1
    gh = host_graph(args)
    from shadow_gnn_b200.synth import PRESETS
    C = PRESETS[args.graph][4]
    labels = np.random.default_rng(7).integers(0, C, gh["N"])
    B = TRAIN_CFG["batch"]
    roots = gh["train"][:min(B * (args.steps + args.warmup) + 500, gh["train"].size)]
    v, kind, n, t = cpu_train_arm(gh, roots, labels, args.steps, args.warmup, threads, C)
    line = {"impl": "reference", "metric": "train_samples_per_sec", "value": v, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t / max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.graph} 5-layer GraphSAGE-256, PPR(k={PPR_K}) sampler, batch {B}, full train step (sample, collate, gather, fwd, bwd, clip, Adam)",
                       "step": "one training batch; reference sampler called for 500 roots at a time (shaDow/minibatch.py:397); model = the reference's library calls on the host cores"},
            "cpu_baseline": {"value": v, "unit": "samples/s", "cores": threads, "kind": kind, "sample": f"{n} targets"},
            "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_ours_train(args):
    import torch
    import torch.distributed as dist
    from shadow_gnn_b200 import minibatch as MB
    from shadow_gnn_b200.models import DeepGNN
    from shadow_gnn_b200.synth import PRESETS
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    assert torch.cuda.is_available(), "bench.py (our arm) needs a GPU: there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(1234)
    np.random.seed(1234)
    g = build_graph(args.graph, dev)
    N, F, C = g["N"], g["F"], g["C"]
    feat = torch.randn(N, F, device=dev, generator=torch.Generator(device=dev).manual_seed(g["seed"] + 100))
    labels = torch.from_numpy(np.random.default_rng(7).integers(0, C, N)).to(dev)
    B = TRAIN_CFG["batch"]
    share = g["train"][rank::world].numpy()                       # this rank's targets (independent units; gradients meet in one all-reduce)
    cfg = {"batch_size": B, "configs": [{"method": "ppr", "k": [PPR_K], "threshold": [0.0], "epsilon": [PPR_EPS]}]}
    adjs = {m: (g["indptr"], g["indices"]) for m in range(3)}
    mb = MB.MinibatchShallowExtractor(args.graph, None, adjs, {0: share, 1: share[:B], 2: share[:B]}, cfg, set(), None, feat, labels, F, True, 1,
                                      seed_cpp=1, num_subg_per_batch=args.superbatch_train)
    model = DeepGNN(F, F, C, 0, ARCH, [], 1, dict(dropout=TRAIN_CFG["dropout"], dropedge=TRAIN_CFG["dropedge"], lr=TRAIN_CFG["lr"], ensemble_dropout="none"),
                    "node").to(dev)
    log("model + minibatch ready; PPR push + first super-batch next")
    mb.epoch_start_reset(0, MB.TRAIN)
    mb.shuffle_entity(MB.TRAIN)
    nparams = sum(p.numel() for p in model.parameters())

    def next_batch():
        if mb.is_end_epoch(MB.TRAIN):
            mb.epoch_end_reset(MB.TRAIN); mb.epoch_start_reset(1, MB.TRAIN); mb.shuffle_entity(MB.TRAIN)
        return mb.one_batch(MB.TRAIN)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local) if rank == 0 else None
    for _ in range(args.warmup):
        model.step(MB.TRAIN, "running", next_batch())
    log("warmup done")
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    nsamp = 0
    for _ in range(args.steps):
        b = next_batch()
        model.step(MB.TRAIN, "running", b)
        nsamp += b.batch_size
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    log("timed region done", ms)
    # ---- e2e: the user-facing call with HOST inputs: the batch's target ids + labels come from pinned host memory every step,
    #      the loss goes back to the host every step ----
    e2e_steps = max(5, min(args.steps, 200))
    tgt_host = torch.from_numpy(share.astype(np.int64)).pin_memory()
    lab_host = labels.cpu()[tgt_host].pin_memory()
    loss_host = torch.zeros(1).pin_memory()
    barrier()
    t0 = time.perf_counter()
    ne2e = 0
    for i in range(e2e_steps):
        lo = (i * B) % (share.size - B)
        t_dev = tgt_host[lo:lo + B].to(dev, non_blocking=True)           # H2D: this step's targets
        y_dev = lab_host[lo:lo + B].to(dev, non_blocking=True)           # H2D: this step's labels
        b = next_batch()
        b.label = y_dev if y_dev.numel() == b.label.numel() else b.label
        out = model.step(MB.TRAIN, "running", b)
        loss_host.copy_(out["loss"].detach().reshape(1), non_blocking=True)   # D2H: the step's loss
        torch.cuda.synchronize()
        ne2e += b.batch_size
        del t_dev
    barrier()
    e2e_s = time.perf_counter() - t0
    clk = clocks.stop() if clocks else None
    stats = torch.tensor([ms, e2e_s], device=dev, dtype=torch.float64)
    counts = torch.tensor([nsamp, ne2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    if rank == 0:
        ms_all, e2e_all = float(stats[0]), float(stats[1])
        line = {"metric": "train_samples_per_sec", "value": float(counts[0]) / (ms_all * 1e-3), "unit": "samples/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_all / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": f"{args.graph} 5-layer GraphSAGE-256 ({nparams} params), PPR(k={PPR_K},eps={PPR_EPS}) sampler, batch {B} per GPU, dropout 0.4, dropedge 0.05, Adam lr 0.002, clip 5",
                           "global_batch": B * world, "sampler_superbatch": args.superbatch_train,
                           "parallelism": f"dp{world}: targets partitioned, one NCCL all-reduce of the flat {nparams * 4} B gradient bucket per step",
                           "l2": "CSR (495 MB) and features (980 MB) exceed L2; every step trains on different roots"},
                "e2e": {"value": float(counts[1]) / e2e_all, "unit": "samples/s", "h2d_bytes_per_step": B * 16, "d2h_bytes_per_step": 4,
                        "note": "targets + labels from pinned host memory each step, loss read back each step"},
                "clocks": clk}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.task is None:
        a.task = "sampler"
    if a.impl == "reference":
        run_reference_train(a) if a.task == "train" else run_reference(a)
    else:
        run_ours_train(a) if a.task == "train" else run_ours(a)
