"""Generate tests/golden/sampler_golden.npz from the UNMODIFIED reference sampler (oracle/_ref).

Run here (container with /root/reference):   make -C oracle ref && python tests/golden/make_golden.py
The reference ships no golden vectors; these are outputs of the reference itself
(backend/ParallelSampler.cpp compiled as-is, single-threaded, fixed seed) on a small synthetic graph
that carries the disconnected 2-node tail component, so that no reference read is out of bounds.
"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O                      # noqa: E402
from shadow_gnn_b200.synth import small_parity_graph   # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sampler_golden.npz")
GRAPH = dict(n=600, avg_deg=10, seed=7, self_loops=15)
PPR_PARAMS = dict(k=48, alpha=0.85, epsilon=1e-4)
P = 16          # num_sampler_per_batch (even, samplers_ensemble.py:95)
SEED = 11


def tf(b):
    return "true" if b else "false"


def cases():
    out = []
    for depth, budget, se, nr, tc, aug in [
        (1, 5, False, 1, False, ()), (2, 10, False, 1, False, ("hops",)), (2, 10, True, 1, False, ("hops",)),
        (2, -1, False, 1, False, ()), (3, 5, True, 1, False, ()), (2, 5, False, 2, False, ("drnls",)),
        (2, 5, True, 2, True, ("drnls",)), (2, 20, False, 2, True, ()), (1, -1, True, 2, False, ("hops",)),
    ]:
        out.append((dict(method="khop", depth=str(depth), budget=str(budget), num_roots=str(nr),
                         add_self_edge=tf(se), include_target_conn=tf(tc)), aug, 37 * nr))
    for k, thr, se, nr, aug in [
        (1, 0, False, 1, ()), (30, 0, False, 1, ("pprs",)), (30, 0, True, 1, ("hops",)), (48, 0.01, False, 1, ()),
        (48, 0.002, True, 1, ("pprs",)), (30, 0.3, False, 1, ()), (20, 0, False, 2, ("drnls",)), (48, 0.01, True, 2, ()),
    ]:
        out.append((dict(method="ppr", k=str(k), threshold=str(thr), num_roots=str(nr),
                         add_self_edge=tf(se), include_target_conn="false"), aug, 41 * nr))
    out.append((dict(method="nodeIID", num_roots="1", add_self_edge="false", include_target_conn="false"), (), 20))
    out.append((dict(method="nodeIID", num_roots="2", add_self_edge="false", include_target_conn="false"), (), 40))
    out.append((dict(method="ppr", k="20", threshold="0", num_roots="1", return_target_only="true"), (), 20))
    out.append((dict(method="ppr_st", k="12", threshold="0.01", num_roots="1", add_self_edge="true",
                     include_target_conn="false"), (), 30))
    return out


def read_ppr_bin(fn, dtype):
    """Binary layout of write_PPR_to_binary_file (backend/ParallelSampler.cpp:94-139)."""
    raw = open(fn, "rb").read()
    alpha, eps = np.frombuffer(raw, np.float32, 2, 0)
    k = int(np.frombuffer(raw, np.int32, 1, 8)[0])
    cnt = int(np.frombuffer(raw, np.uint32, 1, 12)[0])
    body = np.frombuffer(raw, np.uint32, -1, 16)
    rows, pos = [], 0
    for _ in range(cnt):
        ln = int(body[pos]); pos += 1
        rows.append(body[pos:pos + ln].view(dtype).copy()); pos += ln
    return float(alpha), float(eps), k, rows


def main():
    ref = O.load_ref()
    assert ref is not None, "build oracle/_ref first (make -C oracle ref)"
    indptr, indices = small_parity_graph(**GRAPH)
    N = indptr.size - 1
    store = dict(indptr=indptr, indices=indices)
    meta = dict(graph=GRAPH, ppr=PPR_PARAMS, P=P, seed=SEED, cases=[])
    # ---- PPR push golden (all nodes as targets) ----
    tmp = tempfile.mkdtemp()
    fnn, fns = os.path.join(tmp, "n.bin"), os.path.join(tmp, "s.bin")
    r = ref.ParallelSampler(indptr.tolist(), indices.tolist(), [], P, 1, True, True, [], 1, "", "", "", SEED)
    r.preproc_ppr_approximate(list(range(N)), PPR_PARAMS["k"], PPR_PARAMS["alpha"], PPR_PARAMS["epsilon"], fnn, fns)
    a, e, k, nrows = read_ppr_bin(fnn, np.uint32)
    _, _, _, srows = read_ppr_bin(fns, np.float32)
    meta["ppr_file_header"] = dict(alpha=a, epsilon=e, k=k)
    lens = np.array([len(x) for x in nrows], np.uint64)
    store["ppr_ptr"] = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    store["ppr_neighs"] = np.concatenate(nrows).astype(np.uint32)
    store["ppr_scores"] = np.concatenate(srows).astype(np.float32)
    # ---- sampler cases ----
    rng = np.random.default_rng(5)
    for ci, (cfg, aug, T) in enumerate(cases()):
        targets = rng.permutation(N - 2)[:T].astype(np.uint32)      # tail nodes are never roots
        r = ref.ParallelSampler(indptr.tolist(), indices.tolist(), [], P, 1, True, True, [], 1, "", "", "", SEED)
        r.shuffle_targets(targets.tolist())
        if cfg["method"] in ("ppr", "ppr_st"):
            r.preproc_ppr_approximate(list(range(N)), PPR_PARAMS["k"], PPR_PARAMS["alpha"], PPR_PARAMS["epsilon"], fnn, fns)
        subgs, calls = [], []
        for _epoch_call in range(64):
            sub = O.ref_subgraphs(r.parallel_sampler_ensemble([cfg], [set(aug)])[0])
            subgs.extend(sub); calls.append(len(sub))
            if r.get_idx_root() == 0:
                break
        for f in O.FlatBatch.FIELDS:
            arrs = [s[f] for s in subgs]
            store[f"c{ci}_{f}"] = np.concatenate(arrs) if arrs else np.zeros(0)
            store[f"c{ci}_{f}_len"] = np.array([x.size for x in arrs], np.int64)
        store[f"c{ci}_targets"] = targets
        meta["cases"].append(dict(cfg=cfg, aug=list(aug), calls=calls))
    store["meta"] = np.frombuffer(json.dumps(meta).encode(), np.uint8)
    np.savez_compressed(OUT, **store)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", len(meta["cases"]), "cases")


if __name__ == "__main__":
    main()
