"""CPU tests (-m "not gpu"): pin the oracle (oracle/oracle_sampler.c) against
 (1) the committed golden vectors generated from the unmodified reference (tests/golden/make_golden.py),
 (2) libc's own rand(),
 (3) the live compiled reference in oracle/_ref when it is present (this container)."""
import ctypes
import itertools

import numpy as np
import pytest

from oracle import oracle as O
from shadow_gnn_b200.synth import small_parity_graph
from tests.common import Golden, assert_subgraph_equal

G = Golden()


def test_glibc_rand_matches_libc():
    libc = ctypes.CDLL("libc.so.6")
    for seed in (0, 1, 5, 12345, 2 ** 31 - 1):
        libc.srand(seed)
        want = np.array([libc.rand() for _ in range(2000)], dtype=np.uint32)
        assert np.array_equal(O.glibc_rand_stream(seed, 2000), want)


@pytest.mark.parametrize("ci", range(G.num_cases))
def test_oracle_matches_golden(ci):
    cfg, aug, targets, calls, want = G.case(ci)
    o = O.OracleSampler(G.indptr, G.indices, G.meta["P"], 1, G.meta["seed"])
    o.shuffle_targets(targets)
    if cfg["method"] in ("ppr", "ppr_st"):
        o.set_ppr(G.ppr_ptr, G.ppr_neighs, G.ppr_scores)
    got = []
    for ncall in calls:
        b = o.sample(O.cfg_from_cpp_config(cfg, aug=aug))
        assert b.num_subg == ncall
        got.extend(b.subgraphs())
    assert o.get_idx_root() == 0
    assert len(got) == len(want)
    for i, (a, b) in enumerate(zip(want, got)):
        assert_subgraph_equal(a, b, f"case {ci} subgraph {i}")


def test_oracle_ppr_push_matches_golden():
    """float32 forward push, bit-exact incl. ordering (PS.cpp:237-344)."""
    N = G.indptr.size - 1
    p = G.meta["ppr"]
    nb, sc, ln = O.ppr_push(G.indptr, G.indices, np.arange(N, dtype=np.uint32), p["k"], p["alpha"], p["epsilon"], 2)
    ptr, fn, fs = O.ppr_rows_to_csr(N, np.arange(N, dtype=np.uint32), nb, sc, ln)
    assert np.array_equal(ptr, G.ppr_ptr)
    assert np.array_equal(fn, G.ppr_neighs)
    assert fs.tobytes() == G.ppr_scores.tobytes()
    # entry 0 of every row is the target itself
    assert np.array_equal(fn[ptr[:-1].astype(np.int64)], np.arange(N, dtype=np.uint32))


def test_oracle_fixed_mode_rows_are_sorted_unique():
    """`fixed` mode (row bound repaired): every subgraph row is sorted and duplicate-free and
    the edge set equals the bug-compatible one minus the spurious bug-slot edges."""
    indptr, indices = small_parity_graph(800, 10, 3)
    N = indptr.size - 1
    t = np.random.default_rng(0).permutation(N - 2)[:64].astype(np.uint32)
    outs = []
    for fixed in (False, True):
        o = O.OracleSampler(indptr, indices, 64, 1, 1)
        o.shuffle_targets(t)
        outs.append(o.sample(O.make_cfg("khop", depth=2, budget=8, fixed_mode=fixed)).subgraphs())
    nbug = 0
    for a, b in zip(*outs):
        assert np.array_equal(a["node"], b["node"])
        for r in range(b["node"].size):
            row = b["indices"][b["indptr"][r]:b["indptr"][r + 1]].astype(np.int64)
            assert np.all(np.diff(row) > 0)
            v = int(b["node"][r])
            eb = a["edge_index"][a["indptr"][r]:a["indptr"][r + 1]]
            keep = eb < indptr[v + 1]
            nbug += int((~keep).sum())
            assert np.array_equal(a["indices"][a["indptr"][r]:a["indptr"][r + 1]][keep], row.astype(np.uint32))
    assert nbug > 0, "fixture never exercised the PS.cpp:401 bug slot"


def test_collation_restatement():
    indptr, indices = small_parity_graph(500, 8, 1)
    o = O.OracleSampler(indptr, indices, 8, 1, 1)
    o.shuffle_targets(np.arange(8, dtype=np.uint32))
    sub = o.sample(O.make_cfg("khop", depth=1, budget=4, add_self_edge=True)).subgraphs()
    c = O.cat_to_block_diagonal(sub)
    n = sum(s["node"].size for s in sub)
    assert c["indptr"].size == n + 1 and c["indptr"][-1] == c["indices"].size
    off = 0
    for i, s in enumerate(sub):       # every block only references its own rows
        rows = slice(off, off + s["node"].size)
        cols = c["indices"][c["indptr"][rows.start]:c["indptr"][rows.stop]]
        assert cols.min() >= off and cols.max() < off + s["node"].size
        assert c["target"][i] == off + s["target"][0]
        off += s["node"].size


# ------------------------------------------------------------------------------------------------
ref = O.load_ref()
needs_ref = pytest.mark.skipif(ref is None, reason="oracle/_ref (compiled reference) not present")


def _both(indptr, indices, targets, P, seed, cfg, aug, ppr=None):
    r = ref.ParallelSampler(indptr.tolist(), indices.tolist(), [], P, 1, True, True, [], 1, "", "", "", seed)
    r.shuffle_targets(targets.tolist())
    o = O.OracleSampler(indptr, indices, P, 1, seed)
    o.shuffle_targets(targets)
    if ppr is not None:
        alln = np.arange(indptr.size - 1, dtype=np.uint32)
        r.preproc_ppr_approximate(alln.tolist(), ppr[0], ppr[1], ppr[2], "", "")
        o.preproc_ppr_approximate(alln, *ppr)
    for _ in range(100):
        rs = O.ref_subgraphs(r.parallel_sampler_ensemble([cfg], [set(aug)])[0])
        os_ = o.sample(O.cfg_from_cpp_config(cfg, aug=aug)).subgraphs()
        assert len(rs) == len(os_)
        for i, (a, b) in enumerate(zip(rs, os_)):
            assert_subgraph_equal(a, b, f"{cfg} {aug} subgraph {i}")
        assert r.get_idx_root() == o.get_idx_root()
        if r.get_idx_root() == 0:
            break


@needs_ref
def test_oracle_vs_live_reference_khop(capfd):
    indptr, indices = small_parity_graph(1500, 12, 2, self_loops=25)
    rng = np.random.default_rng(1)
    N = indptr.size - 1
    for depth, budget, se, nr, tc, aug in itertools.product([1, 2], [-1, 6], [False, True], [1, 2], [False, True],
                                                            [(), ("hops",), ("drnls",)]):
        if aug == ("drnls",) and nr == 1:
            continue
        cfg = dict(method="khop", depth=str(depth), budget=str(budget), num_roots=str(nr),
                   add_self_edge="true" if se else "false", include_target_conn="true" if tc else "false")
        _both(indptr, indices, rng.permutation(N - 2)[:45 * nr].astype(np.uint32), 20, 4, cfg, aug)


@needs_ref
def test_oracle_vs_live_reference_ppr(capfd):
    indptr, indices = small_parity_graph(1500, 12, 4, self_loops=25)
    rng = np.random.default_rng(2)
    N = indptr.size - 1
    for k, thr, se, nr in itertools.product([1, 25, 60], [0, 0.01, 0.4], [False, True], [1, 2]):
        cfg = dict(method="ppr", k=str(k), threshold=str(thr), num_roots=str(nr),
                   add_self_edge="true" if se else "false", include_target_conn="false")
        _both(indptr, indices, rng.permutation(N - 2)[:45 * nr].astype(np.uint32), 20, 4, cfg, ("pprs",), ppr=(60, 0.85, 2e-4))
