"""GPU parity tests (-m gpu): the CUDA sampler, called through the C ABI (ctypes), against
 (1) the golden vectors produced by the unmodified reference, (2) the CPU oracle on seeded random inputs,
 (3) size-independent properties at larger sizes.   Integer / index outputs: bit-exact.  ppr column: bit pattern."""
import itertools

import numpy as np
import pytest

from tests.common import Golden, assert_subgraph_equal, FIELDS

pytestmark = pytest.mark.gpu


def _product():
    import shadow_gnn_b200.ParallelSampler as PS
    return PS


def _subgraphs(vec):
    return [{f: vec.numpy(f)[p] for f in FIELDS} for p in range(vec.get_num_valid_subg())]


G = Golden()


@pytest.mark.parametrize("ci", range(G.num_cases))
def test_cuda_matches_reference_golden(ci):
    PS = _product()
    cfg, aug, targets, calls, want = G.case(ci)
    s = PS.ParallelSampler(G.indptr, G.indices, [], G.meta["P"], 1, True, True, [], 1, "", "", "", G.meta["seed"])
    s.shuffle_targets(targets)
    if cfg["method"] in ("ppr", "ppr_st"):
        s.set_ppr_tables(G.ppr_ptr, G.ppr_neighs, G.ppr_scores)
    got = []
    for ncall in calls:
        vec = s.parallel_sampler_ensemble([cfg], [set(aug)])[0]
        assert vec.get_num_valid_subg() == ncall
        got.extend(_subgraphs(vec))
    assert s.get_idx_root() == 0
    for i, (a, b) in enumerate(zip(want, got)):
        assert_subgraph_equal(a, b, f"case {ci} subgraph {i}")


def test_pybind_list_api_shapes():
    """getters return list[list] of capacity num_sampler_per_batch like the pybind module (G.h:62-72)"""
    PS = _product()
    cfg, aug, targets, calls, want = G.case(1)
    s = PS.ParallelSampler(G.indptr.tolist(), G.indices.tolist(), [], G.meta["P"], 1, True, True, [], 1, "", "", "", G.meta["seed"])
    s.shuffle_targets(targets.tolist())
    vec = s.parallel_sampler_ensemble([cfg], [set(aug)])[0]
    clip = vec.get_num_valid_subg()
    assert len(vec.get_subgraph_indptr()) == G.meta["P"]
    assert vec.get_subgraph_node()[0] == want[0]["node"].tolist()
    assert vec.get_subgraph_data()[0] == [1.0] * want[0]["indices"].size
    assert vec.get_subgraph_hop()[clip - 1] == want[clip - 1]["hop"].tolist()
    assert s.num_nodes() == G.indptr.size - 1 and s.num_edges() == G.indices.size and s.is_seq_root_traversal()


def test_config_errors_match_pybind_exceptions():
    PS = _product()
    s = PS.ParallelSampler(G.indptr, G.indices, [], 4, 1, True, True, [], 1, "", "", "", 1)
    s.shuffle_targets(np.arange(8, dtype=np.uint32))
    with pytest.raises(IndexError):        # config.at("depth") -> std::out_of_range
        s.parallel_sampler_ensemble([dict(method="khop", budget="3", num_roots="1")], [set()])
    with pytest.raises(ValueError):        # std::stoi("x") -> std::invalid_argument
        s.parallel_sampler_ensemble([dict(method="khop", depth="x", budget="3", num_roots="1")], [set()])
    with pytest.raises(ValueError):
        s.parallel_sampler_ensemble([dict(depth="1", budget="3", num_roots="1")], [set()])
    with pytest.raises(Exception):         # ppr before the tables exist
        s.parallel_sampler_ensemble([dict(method="ppr", k="3", threshold="0", num_roots="1")], [set()])


def _oracle_vs_cuda(indptr, indices, targets, P, seed, cfg, aug, ppr_tables=None, fixed=False):
    from oracle import oracle as O
    PS = _product()
    o = O.OracleSampler(indptr, indices, P, 1, seed)
    o.shuffle_targets(targets)
    s = PS.ParallelSampler(indptr, indices, [], P, 1, True, True, [], 1, "", "", "", seed, strict_reference_compat=not fixed)
    s.shuffle_targets(targets)
    if ppr_tables is not None:
        o.set_ppr(*ppr_tables)
        s.set_ppr_tables(*ppr_tables)
    n = 0
    for _ in range(1000):
        want = o.sample(O.cfg_from_cpp_config(cfg, aug=aug, fixed_mode=fixed)).subgraphs()
        got = _subgraphs(s.parallel_sampler_ensemble([cfg], [set(aug)])[0])
        assert len(want) == len(got)
        for i, (a, b) in enumerate(zip(want, got)):
            assert_subgraph_equal(a, b, f"{cfg} {aug} subgraph {n + i}")
        n += len(want)
        assert o.get_idx_root() == s.get_idx_root()
        if o.get_idx_root() == 0:
            break
    return n


def test_khop_grid_vs_oracle():
    from shadow_gnn_b200.synth import small_parity_graph
    indptr, indices = small_parity_graph(4000, 14, 21, self_loops=60)
    N = indptr.size - 1
    rng = np.random.default_rng(3)
    for depth, budget, se, nr, tc, aug in itertools.product([1, 2, 3], [-1, 5, 20], [False, True], [1, 2], [False, True],
                                                            [(), ("hops",), ("drnls",)]):
        if (aug == ("drnls",) and nr == 1) or (depth == 3 and budget != 5) or (tc and nr == 1):
            continue
        cfg = dict(method="khop", depth=str(depth), budget=str(budget), num_roots=str(nr),
                   add_self_edge="true" if se else "false", include_target_conn="true" if tc else "false")
        T = (130 if budget != -1 else 40) * nr
        _oracle_vs_cuda(indptr, indices, rng.permutation(N - 2)[:T].astype(np.uint32), 50, 9, cfg, aug)


def test_khop_fixed_mode_and_stream_persistence_vs_oracle():
    from shadow_gnn_b200.synth import small_parity_graph
    indptr, indices = small_parity_graph(3000, 12, 5)
    N = indptr.size - 1
    t = np.random.default_rng(0).permutation(N - 2)[:700].astype(np.uint32)
    cfg = dict(method="khop", depth="2", budget="10", num_roots="1", add_self_edge="false", include_target_conn="false")
    assert _oracle_vs_cuda(indptr, indices, t, 100, 1, cfg, ("hops",), fixed=True) == 700
    assert _oracle_vs_cuda(indptr, indices, t, 500, 1, cfg, ("hops",)) == 700      # two calls: the rand() stream continues


def test_ppr_grid_vs_oracle():
    from oracle import oracle as O
    from shadow_gnn_b200.synth import small_parity_graph
    indptr, indices = small_parity_graph(5000, 16, 8, self_loops=50)
    N = indptr.size - 1
    rng = np.random.default_rng(4)
    alln = np.arange(N, dtype=np.uint32)
    nb, sc, ln = O.ppr_push(indptr, indices, alln, 200, 0.85, 1e-5, 8)
    tables = O.ppr_rows_to_csr(N, alln, nb, sc, ln)
    for k, thr, se, nr, aug in itertools.product([1, 30, 150, 200], [0, 0.002, 0.01], [False, True], [1, 2], [(), ("hops",), ("drnls",)]):
        if aug == ("drnls",) and nr == 1:
            continue
        cfg = dict(method="ppr", k=str(k), threshold=str(thr), num_roots=str(nr), add_self_edge="true" if se else "false",
                   include_target_conn="false")
        _oracle_vs_cuda(indptr, indices, rng.permutation(N - 2)[:90 * nr].astype(np.uint32), 64, 2, cfg, aug, ppr_tables=tables)


def test_edge_cases_vs_oracle():
    """degree-0 roots, epoch tail shorter than a batch, odd root count for 2-root groups, P larger than the epoch"""
    indptr = np.array([0, 0, 2, 3, 5, 5, 6, 7], dtype=np.uint32)       # nodes 0 and 4 are isolated; (5,6) is the tail component
    indices = np.array([2, 3, 1, 1, 3, 6, 5], dtype=np.uint32)         # node 3 carries a self loop
    t = np.array([0, 1, 2, 3, 4, 1, 0], dtype=np.uint32)
    for nr, se in itertools.product([1, 2], [False, True]):
        for m in ("khop", "nodeIID"):
            cfg = dict(method=m, depth="2", budget="1", num_roots=str(nr), add_self_edge="true" if se else "false",
                       include_target_conn="false")
            _oracle_vs_cuda(indptr, indices, t, 4, 3, cfg, ("hops",) if nr == 1 else ("drnls",))
            _oracle_vs_cuda(indptr, indices, t, 64, 3, cfg, ())


def test_return_target_only_and_ensemble():
    PS = _product()
    cfg_a = dict(method="khop", depth="1", budget="-1", num_roots="1", add_self_edge="true", include_target_conn="false")
    cfg_b = dict(method="nodeIID", num_roots="1", add_self_edge="false", include_target_conn="false", return_target_only="true")
    s = PS.ParallelSampler(G.indptr, G.indices, [], 8, 1, True, True, [], 2, "", "", "", 1)
    t = np.arange(20, dtype=np.uint32)
    s.shuffle_targets(t)
    a, b = s.parallel_sampler_ensemble([cfg_a, cfg_b], [set(), set()])
    assert a.get_num_valid_subg() == b.get_num_valid_subg() == 8
    for p in range(8):                              # both branches saw the same roots (PS.cpp:672)
        assert b.numpy("node")[p].tolist() == [t[p]]
        assert a.numpy("node")[p][a.numpy("target")[p][0]] == t[p]
        assert b.numpy("indptr")[p].size == 0
    s.drop_full_graph_info()
    with pytest.raises(Exception):
        s.parallel_sampler_ensemble([cfg_a, cfg_b], [set(), set()])


def test_device_batch_is_block_diagonal_collation():
    """sample_to_device == Subgraph.cat_to_block_diagonal of the per-subgraph results (frontend/graph.py:280-320)"""
    import torch
    from oracle import oracle as O
    PS = _product()
    cfg, aug, targets, calls, want = G.case(2)
    s = PS.ParallelSampler(G.indptr, G.indices, [], 64, 1, True, True, [], 1, "", "", "", G.meta["seed"])
    s.shuffle_targets(targets)
    b = s.sample_to_device([cfg], [set(aug)])[0]
    col = O.cat_to_block_diagonal(want[:b.num_subg])
    assert np.array_equal(b.rowptr.cpu().numpy(), col["indptr"])
    assert np.array_equal(b.indices.cpu().numpy(), col["indices"])
    assert np.array_equal(b.orig_node.cpu().numpy().view(np.uint32), col["node"])
    assert np.array_equal(b.orig_edge.cpu().numpy().view(np.uint32), col["edge_index"])
    assert np.array_equal(b.target.cpu().numpy().ravel(), col["target"])
    assert np.array_equal(torch.diff(b.node_ptr).cpu().numpy(), col["size_subg"])


def test_ppr_push_gpu_bit_exact_vs_reference_golden():
    PS = _product()
    N = G.indptr.size - 1
    p = G.meta["ppr"]
    s = PS.ParallelSampler(G.indptr, G.indices, [], 4, 1, True, True, [], 1, "", "", "", 1)
    s.preproc_ppr_approximate(np.arange(N, dtype=np.uint32), p["k"], p["alpha"], p["epsilon"], "", "")
    for v in range(N):
        nb, sc = s.get_ppr_row(v)
        a, b = int(G.ppr_ptr[v]), int(G.ppr_ptr[v + 1])
        assert np.array_equal(nb, G.ppr_neighs[a:b]), v
        assert sc.tobytes() == G.ppr_scores[a:b].tobytes(), v


def test_ppr_bin_cache_roundtrip(tmp_path):
    """cache written by the CUDA path has the reference's layout (PS.cpp:94-139) and is re-read with its validity rule"""
    PS = _product()
    N = G.indptr.size - 1
    fn, fs = str(tmp_path / "neighs.bin"), str(tmp_path / "scores.bin")
    s = PS.ParallelSampler(G.indptr, G.indices, [], 4, 1, True, True, [], 1, "", "", "", 1)
    s.preproc_ppr_approximate(np.arange(N, dtype=np.uint32), 48, 0.85, 1e-4, fn, fs)
    raw = np.fromfile(fn, np.uint32)
    assert np.frombuffer(raw[:2].tobytes(), np.float32)[0] == np.float32(1 - np.float32(0.85))
    assert raw[2] == 48 and raw[3] == N
    s2 = PS.ParallelSampler(G.indptr, G.indices, [], 4, 1, True, True, [], 1, "", "", "", 1)
    s2.preproc_ppr_approximate(np.zeros(0, np.uint32), 20, 0.85, 1.05e-4, fn, fs)      # smaller k, eps within 10 %: loads
    nb, sc = s2.get_ppr_row(7)
    a = int(G.ppr_ptr[7])
    assert np.array_equal(nb, G.ppr_neighs[a:a + nb.size]) and nb.size == min(20, int(G.ppr_ptr[8]) - a)


def test_gather_rows():
    import torch
    PS = _product()
    for F in (100, 128, 7):
        feat = torch.randn(5000, F, device="cuda")
        ids = torch.randint(0, 5000, (12345,), device="cuda", dtype=torch.int32)
        out = PS.gather_rows(feat, ids)
        assert torch.equal(out, feat[ids.long()])


def test_superbatch_properties_philox():
    """larger scale, production RNG: structural invariants that do not need the oracle"""
    import torch
    from shadow_gnn_b200.synth import powerlaw_graph
    PS = _product()
    indptr, indices = powerlaw_graph(200_000, 3_000_000, 5, dmax=2000)
    N = indptr.size - 1
    P = 4096
    s = PS.ParallelSampler(indptr, indices, [], P, 1, True, True, [], 1, "", "", "", 7, rng="philox", strict_reference_compat=False)
    t = np.random.default_rng(1).permutation(N)[:P].astype(np.uint32)
    s.shuffle_targets(t)
    cfg = dict(method="khop", depth="2", budget="10", num_roots="1", add_self_edge="true", include_target_conn="false")
    b = s.sample_to_device([cfg], [{"hops"}])[0]
    assert b.num_subg == P
    node_ptr, rowptr, idx = b.node_ptr.long(), b.rowptr.long(), b.indices.long()
    n = torch.diff(node_ptr)
    assert int(n.max()) <= 111 and int(n.min()) >= 1
    assert int(rowptr[-1]) == b.total_edges and bool((torch.diff(rowptr) >= 1).all())        # self edge => no empty row
    sub_of_row = torch.repeat_interleave(torch.arange(P, device="cuda"), n)
    row_of_edge = torch.repeat_interleave(torch.arange(b.total_nodes, device="cuda"), torch.diff(rowptr))
    assert bool((sub_of_row[idx] == sub_of_row[row_of_edge]).all())                          # block diagonal
    on = b.orig_node.long() & 0xFFFFFFFF
    assert bool((on[b.target.long().ravel()] == torch.as_tensor(t.astype(np.int64), device="cuda")).all())
    # every edge is a real edge of the full graph or the inserted self loop; rows are sorted (fixed mode)
    oe = b.orig_edge.long() & 0xFFFFFFFF
    real = oe != 0xFFFFFFFF
    full_idx = torch.as_tensor(indices.astype(np.int64), device="cuda")
    assert bool((full_idx[oe[real]] == on[idx[real]]).all())
    assert bool((on[idx[~real]] == on[row_of_edge[~real]]).all())
    same_row = row_of_edge[1:] == row_of_edge[:-1]
    assert bool((idx[1:][same_row] > idx[:-1][same_row]).all())
    # hop label of the root is 0 and every 1-hop neighbour has label 1
    hop = b.hop.long() & 0xFFFFFFFF
    assert bool((hop[b.target.long().ravel()] == 0).all()) and int(hop.max()) <= 2
    # same seed => same sample (idempotence); different epoch => different sample
    s2 = PS.ParallelSampler(indptr, indices, [], P, 1, True, True, [], 1, "", "", "", 7, rng="philox", strict_reference_compat=False)
    s2.shuffle_targets(t)
    b2 = s2.sample_to_device([cfg], [{"hops"}])[0]
    assert torch.equal(b.orig_node, b2.orig_node) and torch.equal(b.indices, b2.indices)
    b3 = s2.sample_to_device([cfg], [{"hops"}])[0]
    assert not torch.equal(b3.orig_node[:1000], b.orig_node[:1000])


def test_ppr_st_grid_vs_oracle():
    """ppr_stochastic incl. its integer-division quirk (PS.cpp:635) and rand() stream consumption, 1 and 2 roots, two calls per epoch"""
    from oracle import oracle as O
    from shadow_gnn_b200.synth import small_parity_graph
    indptr, indices = small_parity_graph(3000, 12, 11, self_loops=30)
    N = indptr.size - 1
    alln = np.arange(N, dtype=np.uint32)
    nb, sc, ln = O.ppr_push(indptr, indices, alln, 60, 0.85, 1e-4, 8)
    tables = O.ppr_rows_to_csr(N, alln, nb, sc, ln)
    rng = np.random.default_rng(6)
    for k, thr, se, nr in itertools.product([1, 20, 30], [0, 0.01, 0.3], [False, True], [1, 2]):
        cfg = dict(method="ppr_st", k=str(k), threshold=str(thr), num_roots=str(nr), add_self_edge="true" if se else "false", include_target_conn="false")
        _oracle_vs_cuda(indptr, indices, rng.permutation(N - 2)[:70 * nr].astype(np.uint32), 48, 5, cfg, (), ppr_tables=tables)


# ------------------------------------------------------------------------------------------------
# single-root PPR fast path (one warp per subgraph, csrc/ppr_warp_kernel.cuh) and its hand-over to the generic kernel
# ------------------------------------------------------------------------------------------------
class _Env:
    def __init__(self, **kw):
        self.kw = kw

    def __enter__(self):
        import os
        self.old = {k: os.environ.get(k) for k in self.kw}
        for k, v in self.kw.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = str(v)

    def __exit__(self, *a):
        import os
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _ppr_tables(indptr, indices, k, eps=1e-5):
    from oracle import oracle as O
    N = indptr.size - 1
    alln = np.arange(N, dtype=np.uint32)
    nb, sc, ln = O.ppr_push(indptr, indices, alln, k, 0.85, eps, 8)
    return O.ppr_rows_to_csr(N, alln, nb, sc, ln)


@pytest.mark.parametrize("env", [dict(), dict(SHADOW_WARP_ECAP_MULT=0), dict(SHADOW_WARP_NF=2, SHADOW_WARP_BUCKET_MULT=1),
                                 dict(SHADOW_WARP_ECAP_MULT=1, SHADOW_WARP_NF=4), dict(SHADOW_NO_WARP_PPR=1), dict(SHADOW_NO_SYM=1),
                                 dict(SHADOW_NO_SYM=1, SHADOW_WARP_ECAP_MULT=0),
                                 dict(SHADOW_WARP_BISECT=1), dict(SHADOW_WARP_BISECT=1, SHADOW_NO_SYM=1), dict(SHADOW_WARP_BISECT=0)])
def test_ppr_warp_path_and_redo_vs_oracle(env):
    """the fast path, its staging-overflow and bucket-overflow hand-over (redo launch of the generic kernel) and the generic
    kernel alone all reproduce the oracle bit for bit: bug-compatible and fixed mode, self edge on/off, thresholds, k=1"""
    from shadow_gnn_b200.synth import small_parity_graph
    indptr, indices = small_parity_graph(6000, 24, 13, self_loops=80)
    N = indptr.size - 1
    tables = _ppr_tables(indptr, indices, 160)
    rng = np.random.default_rng(9)
    with _Env(**env):
        for k, thr, se, fixed in itertools.product([1, 40, 150], [0, 0.004], [False, True], [False, True]):
            cfg = dict(method="ppr", k=str(k), threshold=str(thr), num_roots="1", add_self_edge="true" if se else "false",
                       include_target_conn="false")
            t = rng.permutation(N - 2)[:300].astype(np.uint32)
            assert _oracle_vs_cuda(indptr, indices, t, 128, 2, cfg, (), ppr_tables=tables, fixed=fixed) == 300


def test_ppr_sym_variant_is_selected_and_directed_graphs_fall_back():
    """the symmetric-graph variant of the fast path (upper-triangle scan + mirrored edges through the reverse-slot index) runs on graphs that
    pass the one-time check and only there: dropping reverse edges (a directed graph) or SHADOW_NO_SYM=1 selects the full scan; all three
    reproduce the oracle bit for bit"""
    from shadow_gnn_b200.synth import small_parity_graph
    PS = _product()
    indptr, indices = small_parity_graph(3000, 20, 21, self_loops=40)
    N = indptr.size - 1
    rows = np.repeat(np.arange(N), np.diff(indptr.astype(np.int64)))
    keep = ~((rows > indices) & ((rows + indices) % 7 == 0))          # drop every 7th-ish lower-triangle edge: no longer symmetric
    ip_d = np.zeros(N + 1, np.int64); np.cumsum(np.bincount(rows[keep], minlength=N), out=ip_d[1:])
    directed = (ip_d.astype(np.uint32), indices[keep])
    rng = np.random.default_rng(4)
    for (ip, ix), env, want_sym in (((indptr, indices), dict(), True), ((indptr, indices), dict(SHADOW_NO_SYM=1), False), (directed, dict(), False)):
        tables = _ppr_tables(ip, ix, 80)
        with _Env(**env):
            for se, fixed in itertools.product([False, True], [False, True]):
                cfg = dict(method="ppr", k="80", threshold="0.001", num_roots="1", add_self_edge="true" if se else "false", include_target_conn="false")
                t = rng.permutation(N - 2)[:256].astype(np.uint32)
                assert _oracle_vs_cuda(ip, ix, t, 128, 2, cfg, (), ppr_tables=tables, fixed=fixed) == 256
            s = PS.ParallelSampler(ip, ix, [], 64, 1, True, True, [], 1, "", "", "", 1)
            s.set_ppr_tables(*tables); s.shuffle_targets(t)
            s.sample_to_device([cfg], [set()])
            assert s.last_sym() == want_sym, (env, want_sym)


@pytest.mark.parametrize("env", [dict(), dict(SHADOW_NO_SYM=1), dict(SHADOW_WARP_ECAP_MULT=0), dict(SHADOW_NO_WARP_PPR=1), dict(SHADOW_WARP_BISECT=1)])
def test_ppr_with_hop_labels_on_the_fast_path_vs_oracle(env):
    """`feature_augment: hops` with a PPR sampler (46 of the reference's 61 configs): the one-warp fast path labels hops with a warp-level BFS
    (both scan variants, the redo hand-over, and the generic kernel for comparison) -- every array incl. `hop` equals the oracle's"""
    from shadow_gnn_b200.synth import small_parity_graph
    PS = _product()
    indptr, indices = small_parity_graph(5000, 20, 17, self_loops=60)
    N = indptr.size - 1
    tables = _ppr_tables(indptr, indices, 150)
    rng = np.random.default_rng(6)
    with _Env(**env):
        for k, thr, se, fixed in itertools.product([1, 30, 150], [0, 0.01], [False, True], [False, True]):
            cfg = dict(method="ppr", k=str(k), threshold=str(thr), num_roots="1", add_self_edge="true" if se else "false", include_target_conn="false")
            t = rng.permutation(N - 2)[:192].astype(np.uint32)
            t[0] = N - 1                              # a node of the 2-node tail component: most of its PPR scope is unreachable or tiny
            assert _oracle_vs_cuda(indptr, indices, t, 96, 2, cfg, ("hops",), ppr_tables=tables, fixed=fixed) == 192
        s = PS.ParallelSampler(indptr, indices, [], 64, 1, True, True, [], 1, "", "", "", 1)
        s.set_ppr_tables(*tables); s.shuffle_targets(t)
        b = s.sample_to_device([cfg], [{"hops"}])[0]
        assert b.hop is not None and int(b.hop.numel()) == b.total_nodes
        if not env:
            assert s.last_sym(), "hop labels must not push a PPR config off the fast path"


def test_ppr_k400_papers_config_vs_oracle():
    """BASELINE configs[4] sampler parameters (papers100M: k = 400, threshold 0.002; node capacity 401 > the 320 up to which the symmetric
    variant's row offsets share the chunk queue's shared memory, 2 Bloom words per lane) against the oracle, self edge on and off"""
    from oracle import oracle as O
    from shadow_gnn_b200.synth import small_parity_graph
    indptr, indices = small_parity_graph(8000, 40, 5, self_loops=50)
    N = indptr.size - 1
    t = np.random.default_rng(3).permutation(N - 2)[:192].astype(np.uint32)
    nb, sc, ln = O.ppr_push(indptr, indices, t, 400, 0.85, 1e-5, 8)
    tables = O.ppr_rows_to_csr(N, t, nb, sc, ln)
    for env in (dict(), dict(SHADOW_WARP_BISECT=0), dict(SHADOW_NO_SYM=1)):      # default at this size: exact membership by bisection (no hash table)
        with _Env(**env):
            for se, thr in (("false", "0.002"), ("true", "0")):
                cfg = dict(method="ppr", k="400", threshold=thr, num_roots="1", add_self_edge=se, include_target_conn="false")
                assert _oracle_vs_cuda(indptr, indices, t, 64, 2, cfg, (), ppr_tables=tables) == 192


def test_ppr_warp_redo_is_exercised():
    """with a staging area of just one scan stage every subgraph that needs a second stage goes through the redo launch (and still matches)"""
    from shadow_gnn_b200.synth import small_parity_graph
    PS = _product()
    indptr, indices = small_parity_graph(6000, 24, 13)
    N = indptr.size - 1
    tables = _ppr_tables(indptr, indices, 100)
    t = np.random.default_rng(2).permutation(N - 2)[:512].astype(np.uint32)
    cfg = dict(method="ppr", k="100", threshold="0", num_roots="1", add_self_edge="false", include_target_conn="false")
    out = {}
    for name, env in (("warp", dict()), ("redo", dict(SHADOW_WARP_ECAP_MULT=0)), ("generic", dict(SHADOW_NO_WARP_PPR=1))):
        with _Env(**env):
            s = PS.ParallelSampler(indptr, indices, [], 512, 1, True, True, [], 1, "", "", "", 1)
            s.set_ppr_tables(*tables)
            s.shuffle_targets(t)
            b = s.sample_to_device([cfg], [set()])[0]
            out[name] = {f: getattr(b, f).cpu().numpy().copy() for f in ("node_ptr", "rowptr", "indices", "orig_node", "orig_edge", "target", "ppr")}
            out[name]["redo"] = s.last_redo_count()
    assert out["warp"]["redo"] < 512 and out["generic"]["redo"] == 0
    assert out["redo"]["redo"] > 0, "a one-stage staging area should overflow for some subgraphs"
    for name in ("warp", "redo"):
        for f, v in out["generic"].items():
            if f != "redo":
                assert np.array_equal(out[name][f], v), (name, f)


def test_ppr_warp_superbatch_equals_generic_kernel():
    """larger scale (hub rows of a few thousand slots, 8192 subgraphs per launch): fast path == generic kernel, bit for bit,
    on the canonical CSR; plus structural properties (every kept edge is the edge its orig_edge names)"""
    import torch
    from shadow_gnn_b200.synth import powerlaw_graph
    PS = _product()
    indptr, indices = powerlaw_graph(300_000, 9_000_000, 5, dmax=6000)
    N = indptr.size - 1
    P = 8192
    t = np.random.default_rng(1).permutation(N)[:P].astype(np.uint32)
    res = {}
    for name, env in (("warp", dict()), ("generic", dict(SHADOW_NO_WARP_PPR=1))):
        with _Env(**env):
            s = PS.ParallelSampler(indptr, indices, [], P, 1, True, True, [], 1, "", "", "", 1)
            s.preproc_ppr_approximate(t, 150, 0.85, 1e-5, "", "")
            s.shuffle_targets(t)
            for se in ("false", "true"):
                cfg = dict(method="ppr", k="150", threshold="0", num_roots="1", add_self_edge=se, include_target_conn="false")
                b = s.sample_to_device([cfg], [set()])[0]
                res[name, se] = {f: getattr(b, f).clone() for f in ("node_ptr", "rowptr", "indices", "orig_node", "orig_edge", "target", "ppr")}
                if name == "warp":
                    res["redo", se] = s.last_redo_count()
    for se in ("false", "true"):
        for f, v in res["generic", se].items():
            assert torch.equal(res["warp", se][f], v), (se, f)
        assert res["redo", se] < P // 4
    r = res["warp", "false"]
    oe = r["orig_edge"].long() & 0xFFFFFFFF
    on = r["orig_node"].long() & 0xFFFFFFFF
    full_idx = torch.as_tensor(indices.astype(np.int64), device="cuda")
    assert bool((full_idx[oe] == on[r["indices"].long()]).all())
    assert bool((on[r["target"].long().ravel()] == torch.as_tensor(t.astype(np.int64), device="cuda")).all())
