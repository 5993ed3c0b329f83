"""CPU model of the single-root PPR fast path (shadow_gnn_b200/csrc/ppr_warp_kernel.cuh) checked against the oracle.

The CUDA kernel cannot run without a GPU, but its ALGORITHM can: this file restates, lane by lane and window by window, what one warp
does for one subgraph -- node set from the id-sorted table row (root slot reserved while scanning), the PS.cpp:401 slot folded into
the row when no self edge is inserted, the flat space of aligned chunks with the 32-chunk row resolution (`REDUX.OR` mask + popc),
validity masks for the alignment padding, lane-major staging order, per-row counters, and the self-edge / bug-slot post pass -- and
compares the resulting CSR with the oracle's (which is pinned bit-exact to the compiled reference).  It guards the logic the GPU parity
tests (`tests/test_sampler_gpu.py -m gpu`) exercise for real; it is test infrastructure like `oracle/`.
"""
import itertools

import numpy as np
import pytest

from oracle import oracle as O
from shadow_gnn_b200.synth import small_parity_graph

NONE = 0xFFFFFFFF
def popc(x):
    return bin(x & 0xFFFFFFFF).count("1")


def build_rev(indptr, indices):
    """sym_build_kernel (sampler.cu): rev[slot of (u,v)] = slot of (v,u); None when the graph is not symmetric with strictly ascending rows"""
    N = indptr.size - 1
    rev = np.zeros(indices.size, np.int64)
    for u in range(N):
        s, e = int(indptr[u]), int(indptr[u + 1])
        for slot in range(s, e):
            v = int(indices[slot])
            if v >= N or (slot + 1 < e and int(indices[slot + 1]) <= v): return None
            lo, end = int(indptr[v]), int(indptr[v + 1])
            lo += int(np.searchsorted(indices[lo:end], u, side="left"))
            if lo >= end or int(indices[lo]) != u: return None
            rev[slot] = lo
    return rev


def upper(indptr, indices, i):
    """ppr_upper_kernel: {first slot of row i whose neighbour is >= i, slots to the row end}"""
    s, e = int(indptr[i]), int(indptr[i + 1])
    lb = s + int(np.searchsorted(indices[s:e], i, side="left"))
    return (lb, e - lb)


def emulate(indptr, indices, ptr, neighs, scores, t, k, thr, add_self, fixed, CS, U, ecap=10**9, rev=None):
    sym = rev is not None
    E = indices.size
    # ---- id-sorted table row (install-time)
    off = int(ptr[t]); len_all = int(ptr[t+1]) - off
    nb_row = neighs[off:off+len_all]; sc_row = scores[off:off+len_all]
    order = np.argsort(nb_row, kind="stable")
    sid = nb_row[order]; sscore = sc_row[order]; srank = order.astype(np.int64)
    srow = [upper(indptr, indices, int(i)) if sym else (int(indptr[i]), int(indptr[i+1]-indptr[i])) for i in sid]
    # ---- A
    size_neigh = min(len_all, k)
    max_ppr = np.float32(sc_row[1]) if size_neigh > 1 else np.float32(0)
    cut = size_neigh
    for i in range(len_all):
        r = srank[i]
        if r < size_neigh and (max_ppr == 0 or np.float32(sscore[i]) / max_ppr < np.float32(thr)): cut = min(cut, r)
    nodes = {}; run = 0; n_below = 0; root_in = False
    for base in range(0, len_all, 32):
        lanes = range(base, min(base+32, len_all))
        sel = {i: srank[i] < cut for i in lanes}
        selp = {i: sel[i] and sid[i] != t for i in lanes}
        m = [i for i in lanes if selp[i]]
        for i in lanes:
            if sel[i]:
                at = run + sum(1 for j in m if j < i) + (1 if (selp[i] and sid[i] > t) else 0)
                assert at not in nodes
                nodes[at] = (int(sid[i]), np.float32(sscore[i]), srow[i])
        run += len(m); n_below += sum(1 for i in m if sid[i] < t); root_in |= any(sel[i] and sid[i] == t for i in lanes)
    if not root_in:
        assert n_below not in nodes
        nodes[n_below] = (int(t), np.float32(sc_row[0]) if (size_neigh <= 1 and len_all > 0) else np.float32(-1), upper(indptr, indices, int(t)) if sym else (int(indptr[t]), int(indptr[t+1]-indptr[t])))
    n = run + 1
    assert sorted(nodes) == list(range(n))
    ids = [nodes[i][0] for i in range(n)]; pprv = [nodes[i][1] for i in range(n)]
    assert ids == sorted(set(ids))
    s = [nodes[i][2][0] for i in range(n)]; d = [nodes[i][2][1] for i in range(n)]
    ext = (not add_self) and (not fixed)
    members = set(ids)
    # ---- B
    extf = [ext and s[i] + d[i] < E for i in range(n)]               # bit 31 of rs[].y in the SYM kernel
    if ext: d = [d[i] + (1 if s[i] + d[i] < E else 0) for i in range(n)]
    CSH = {4: 2, 8: 3}[CS]
    ch = [(((s[i] + d[i] + CS - 1) >> CSH) - (s[i] >> CSH)) if d[i] else 1 for i in range(n)]
    cp = [0]
    for c in ch: cp.append(cp[-1] + c)
    total = cp[n]
    E_al = E & ~(CS - 1)
    st = []; cnt = 0; rbase = 0
    c0 = 0
    while c0 < total and cnt <= ecap:
        for u in range(U):
            cw = c0 + 32*u
            if cw >= total: continue
            mask = 0
            for lane in range(32):
                rel = cp[min(rbase + 1 + lane, n)] - cw
                assert rel >= 0
                if rel < 32: mask |= 1 << rel
            rows = [min(rbase + popc(mask & ((2 << lane) - 1)), n - 1) for lane in range(32)]
            rbase += popc(mask)
            hits = []
            for lane in range(32):
                c = cw + lane
                if c < total:
                    row = rows[lane]
                    assert cp[row] <= c < cp[row+1], (cp[row], c, cp[row+1])
                    b = ((s[row] >> CSH) + (c - cp[row])) << CSH
                    for e in range(CS):
                        slot = b + e
                        o = (slot - s[row]) & 0xFFFFFFFF
                        if b < E_al or slot < E: v = int(indices[slot]) if slot < E else 0
                        else: v = 0
                        if v in members and o < d[row]: hits.append((v, slot, row))
            if cnt + len(hits) <= ecap: st.extend(hits)
            cnt += len(hits)
        c0 += 32 * U
    if cnt > ecap: return None
    if sym: return emit_sym(indices, rev, ids, pprv, s, d, extf, st, t, add_self, fixed, members)
    # per-row facts as the kernel accumulates them: one counter word per row (kept | kept below v | v itself kept)
    kept = [0] * n
    less = [0] * n
    present = [False] * n
    for v, slot, row in st:
        kept[row] += 1
        less[row] += 1 if v < ids[row] else 0
        present[row] |= v == ids[row]
    rlo = [0]
    for c in kept:
        rlo.append(rlo[-1] + c)                                   # first staged entry of every row
    if not add_self:                                              # nothing is inserted: the staged stream is the CSR
        indptr_loc = rlo
        out_idx = [ids.index(x[0]) for x in st]
        out_eid = [x[1] for x in st]
    else:
        rins, rbug, cnts = [], [], []
        for i in range(n):
            bsub = NONE
            if present[i] and not fixed:                          # no insertion => slot e is tested too (PS.cpp:401)
                e = s[i] + d[i]
                if e < E and int(indices[e]) in members:
                    bsub = ids.index(int(indices[e]))
            rins.append(NONE if present[i] else less[i])
            rbug.append(bsub)
            cnts.append(kept[i] + (0 if present[i] else 1) + (1 if bsub != NONE else 0))
        cpn = [0]
        for c in cnts:
            cpn.append(cpn[-1] + c)
        m = cpn[n]
        out_idx = [None] * m
        out_eid = [None] * m
        for g, (v, slot, row) in enumerate(st):
            pos = cpn[row] + (g - rlo[row]) + (1 if (rins[row] != NONE and v > ids[row]) else 0)
            assert out_idx[pos] is None
            out_idx[pos] = ids.index(v)
            out_eid[pos] = slot
        for r in range(n):
            if rins[r] != NONE:
                pos = cpn[r] + rins[r]
                assert out_idx[pos] is None
                out_idx[pos] = r
                out_eid[pos] = NONE
            if rbug[r] != NONE:
                pos = cpn[r + 1] - 1
                assert out_idx[pos] is None
                out_idx[pos] = rbug[r]
                out_eid[pos] = s[r] + d[r]
        indptr_loc = cpn
    return dict(indptr=np.array(indptr_loc), indices=np.array(out_idx, dtype=np.int64), node=np.array(ids), edge_index=np.array(out_eid, dtype=np.int64),
                target=np.array([ids.index(t)]), ppr=np.array(pprv, dtype=np.float32))



def emit_sym(indices, rev, ids, pprv, s, d, extf, st, t, add_self, fixed, members):
    """SYM variant behind the scan (resolve pass, counters, cursors, emit) as the kernel does it; st = staged upper-part hits in CSR order"""
    E = indices.size
    n = len(ids)
    rc = [0] * n                                                  # kept | mirrored << 14 | self loop kept << 28
    for v, slot, row in st:
        rc[row] += 1 + ((1 << 28) if (add_self and v == ids[row]) else 0)
    ent = []
    for v, slot, row in st:                                       # resolve pass
        sv = int(np.searchsorted(ids, v, side="left"))
        is_ext = extf[row] and slot == s[row] + d[row] - 1
        mir = (not is_ext) and v > ids[row]
        if mir: rc[sv] += 1 << 14
        ent.append((row, sv, mir, slot))
    cp = [0]; rlo = [0] * n; rins = [NONE] * n; rbug = [NONE] * n; cur = [0] * n
    carry_k = 0
    for i in range(n):
        w = rc[i]; k_i = w & 0x3fff; mc = (w >> 14) & 0x3fff
        if not add_self: c_i = k_i + (w >> 14)
        else:
            present = (w >> 28) != 0
            bsub = NONE
            if present and not fixed:
                e = s[i] + d[i]
                if e < E and int(indices[e]) in members: bsub = ids.index(int(indices[e]))
            rins[i] = NONE if present else mc
            rbug[i] = bsub
            c_i = k_i + (0 if present else 1) + (1 if bsub != NONE else 0) + mc
        rlo[i] = (cp[i] + mc + (1 if (add_self and rins[i] != NONE) else 0) - carry_k) & 0xFFFFFFFF
        cur[i] = cp[i]
        cp.append(cp[i] + c_i); carry_k += k_i
    m = cp[n]
    out_idx = [None] * m; out_eid = [None] * m
    def put(pos, a, b):
        assert 0 <= pos < m and out_idx[pos] is None, pos
        out_idx[pos] = a; out_eid[pos] = b
    for base in range(0, len(ent), 32):                           # emit, 32 staged entries at a time
        batch = ent[base:base + 32]
        for lane, (row, sv, mir, slot) in enumerate(batch):
            put((rlo[row] + base + lane) & 0xFFFFFFFF, sv, slot)
        snap = list(cur)                                          # every lane reads its cursor before any leader updates it
        for lane, (row, sv, mir, slot) in enumerate(batch):
            if not mir: continue
            rank = sum(1 for l2 in range(lane) if batch[l2][2] and batch[l2][1] == sv)
            put(snap[sv] + rank, row, int(rev[slot]))
            cur[sv] += 1
    if add_self:
        for r in range(n):
            if rins[r] != NONE: put(cp[r] + rins[r], r, NONE)
            if rbug[r] != NONE: put(cp[r + 1] - 1, rbug[r], s[r] + d[r])
    assert all(x is not None for x in out_idx)
    return dict(indptr=np.array(cp), indices=np.array(out_idx, dtype=np.int64), node=np.array(ids), edge_index=np.array(out_eid, dtype=np.int64),
                target=np.array([ids.index(t)]), ppr=np.array(pprv, dtype=np.float32))


@pytest.mark.parametrize("sym", [False, True])
@pytest.mark.parametrize("graph", [(800, 10, 20), (300, 40, 30)])
def test_fast_path_model_matches_oracle(graph, sym):
    N, avg, sl = graph
    indptr, indices = small_parity_graph(N, avg, 13 + N % 7, self_loops=sl)
    NN = indptr.size - 1
    alln = np.arange(NN, dtype=np.uint32)
    nb, sc, ln = O.ppr_push(indptr, indices, alln, 60, 0.85, 1e-4, 8)
    ptr, fn, fs = O.ppr_rows_to_csr(NN, alln, nb, sc, ln)
    rng = np.random.default_rng(N)
    rev = None
    if sym:
        rev = build_rev(indptr, indices)
        assert rev is not None, "small_parity_graph is symmetric with strictly ascending rows"
    checked = 0
    for k, thr, se, fixed, (CS, U) in itertools.product([1, 20, 60], [0, 0.01], [False, True], [False, True], [(4, 4), (8, 2)]):
        t = rng.permutation(NN - 2)[:6].astype(np.uint32)
        t[0] = NN - 3                      # the last real node: its row ends where the tail component's begins
        o = O.OracleSampler(indptr, indices, 64, 1, 1)
        o.set_ppr(ptr, fn, fs)
        o.shuffle_targets(t)
        cfg = O.make_cfg("ppr", num_roots=1, k=k, threshold=thr, add_self_edge=se, include_target_conn=False, fixed_mode=fixed)
        for p, w in enumerate(o.sample(cfg).subgraphs()):
            got = emulate(indptr, indices, ptr, fn, fs, int(t[p]), k, thr, se, fixed, CS, U, rev=rev)
            for f in ("indptr", "indices", "node", "edge_index", "target"):
                assert np.array_equal(np.asarray(w[f]).astype(np.int64), got[f].astype(np.int64)), (k, thr, se, fixed, CS, U, p, f)
            assert w["ppr"].astype(np.float32).tobytes() == got["ppr"].tobytes(), (k, thr, se, fixed, p, "ppr")
            checked += 1
    assert checked == 2 * 2 * 2 * 2 * 3 * 6


def test_sym_check_rejects_directed_and_unsorted_graphs():
    indptr, indices = small_parity_graph(200, 6, 3)
    assert build_rev(indptr, indices) is not None
    ind2 = indices.copy()
    s, e = int(indptr[5]), int(indptr[6])
    victim = next(u for u in range(200) if indptr[u + 1] - indptr[u] >= 2)
    s, e = int(indptr[victim]), int(indptr[victim + 1])
    ind2[s], ind2[s + 1] = ind2[s + 1], ind2[s]                   # unsorted row
    assert build_rev(indptr, ind2) is None
    ind3 = indices.copy()
    ind3[s] = (int(ind3[s]) + 1) if (s + 1 >= e or int(ind3[s]) + 1 < int(ind3[s + 1])) else int(ind3[s])
    if not np.array_equal(ind3, indices):                         # one endpoint changed: its reverse edge is missing
        assert build_rev(indptr, ind3) is None


def _warp_bfs_model(indptr, indices, n, src):
    """lane-level restatement of warp_bfs_hops (ppr_warp_kernel.cuh): 32 frontier nodes at a time, prefix of their row lengths, lane k takes
    edge k of the concatenation, owner = largest lane whose exclusive prefix is <= k (5-step bisection)"""
    dist = [NONE] * n
    dist[src] = 0
    cur, lvl = [src], 0
    while cur:
        nxt = []
        for base in range(0, len(cur), 32):
            s = [0] * 32; ln = [0] * 32
            for lane in range(32):
                if base + lane < len(cur):
                    u = cur[base + lane]; s[lane] = int(indptr[u]); ln[lane] = int(indptr[u + 1]) - s[lane]
            incl = list(np.cumsum(ln)); excl = [incl[i] - ln[i] for i in range(32)]
            tot = incl[31]
            for kb in range(0, tot, 32):
                for lane in range(32):
                    k = kb + lane
                    j = 0
                    for d in (16, 8, 4, 2, 1):
                        if j + d < 32 and excl[(j + d) & 31] <= k: j += d
                    if k < tot:
                        assert excl[j] <= k < incl[j], (k, j, excl, incl)
                        c = int(indices[s[j] + (k - excl[j])])
                        if dist[c] == NONE:
                            dist[c] = lvl + 1; nxt.append(c)
        cur, lvl = nxt, lvl + 1
    return dist


def test_warp_bfs_model_matches_plain_bfs():
    from collections import deque
    rng = np.random.default_rng(5)
    for n, avg in ((1, 0), (40, 1.5), (151, 2.0), (151, 20.0), (400, 3.0)):
        deg = rng.poisson(avg, n) if avg else np.zeros(n, int)
        deg[0] = min(n, 140)                                       # the root's long row
        indptr = np.concatenate([[0], np.cumsum(deg)])
        indices = rng.integers(0, n, int(indptr[-1]))
        want = [NONE] * n; want[0] = 0; q = deque([0])
        while q:
            u = q.popleft()
            for c in indices[indptr[u]:indptr[u + 1]]:
                if want[c] == NONE: want[c] = want[u] + 1; q.append(int(c))
        assert _warp_bfs_model(indptr, indices, n, 0) == want, (n, avg)
