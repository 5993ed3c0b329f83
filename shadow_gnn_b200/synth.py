"""Synthetic stand-ins for the OGB graphs (no network => no datasets), SURVEY.md 8(d).

`powerlaw_graph` follows the recipe fixed there: out-degree min(floor(pareto(1.6)*c)+1, dmax),
uniform random endpoints, self loops dropped, symmetrised, rows sorted + unique, uint32 CSR --
the invariants `to_undirected_csr` guarantees for the graphs the reference samples from
(para_graph_sampler/graph_engine/frontend/graph_utils.py:19-45).
"""
import numpy as np

# name -> (N, nnz_target, dmax, F, C, n_train, seed)      (public OGB sizes)
PRESETS = {
    "S-arxiv":    (169_343,     2_320_000,    13_161, 128, 40,   90_941, 0),
    "S-products": (2_449_029, 123_700_000,    17_481, 100, 47,  196_615, 1),
    "S-papers":   (111_059_956, 3_230_000_000, 100_000, 128, 172, 1_207_179, 2),
    # the same sizes with COMMUNITY structure (planted partition, see clustered_graph_torch): a PPR scope then induces thousands of edges
    "S-products-c": (2_449_029, 123_700_000,  17_481, 100, 47,  196_615, 1),
    "S-papers-8":   (13_882_495,  403_750_000, 100_000, 128, 172,   150_897, 2),      # one eighth of S-papers (per-GPU share of an 8-GPU box)
}
CLUSTERED = {"S-products-c": (192, 0.75)}        # name -> (community size, probability that an edge stays inside its community)


def _coo_to_csr_sym(src, dst, n):
    keep = src != dst
    src, dst = src[keep], dst[keep]
    key = np.concatenate([src.astype(np.int64) * n + dst, dst.astype(np.int64) * n + src])
    key = np.unique(key)
    rows = (key // n).astype(np.int64)
    indices = (key % n).astype(np.uint32)
    indptr = np.zeros(n + 1, np.int64)
    np.cumsum(np.bincount(rows, minlength=n), out=indptr[1:])
    assert indptr[-1] < 2 ** 32
    return indptr.astype(np.uint32), indices


def powerlaw_graph(n, nnz_target, seed, dmax=None, tail=False):
    """Undirected power-law CSR (uint32).  tail=True appends the disconnected 2-node component
    (n, n+1) that keeps the reference's out-of-bounds read away from sampled rows."""
    rng = np.random.default_rng(seed)
    dmax = dmax or max(8, n // 4)
    raw = rng.pareto(1.6, n)
    c = 1.0
    for _ in range(30):      # tune c so the directed edge count is ~ nnz_target / 2
        deg = np.minimum(np.floor(raw * c) + 1, dmax)
        c *= (nnz_target / 2) / deg.sum()
    deg = np.minimum(np.floor(raw * c) + 1, dmax).astype(np.int64)
    src = np.repeat(np.arange(n, dtype=np.int64), deg)
    dst = rng.integers(0, n, src.size, dtype=np.int64)
    if tail:
        src = np.concatenate([src, [n]])
        dst = np.concatenate([dst, [n + 1]])
        n += 2
    return _coo_to_csr_sym(src, dst, n)


def powerlaw_graph_torch(n, nnz_target, seed, dmax, device, community=None):
    """Same recipe on the GPU with torch ops (setup only, not part of any timed region):
    the S-products graph (124 M edges) takes seconds instead of a minute of host sorting.
    community = (size, p_in): planted partition -- nodes [j * size, (j + 1) * size) form a community and an edge's second endpoint is
    drawn inside the first endpoint's community with probability p_in, uniformly otherwise.  Real co-purchase / citation graphs are
    clustered like this; the uniform recipe is not (a 141-node PPR scope keeps 3 % of the slots it scans there, 20-60 % here)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    u = torch.rand(n, generator=g, device=device, dtype=torch.float64).clamp_min(1e-12)
    raw = u.pow(-1.0 / 1.6) - 1.0          # pareto(1.6), numpy's (Lomax) convention
    c = 1.0
    for _ in range(30):
        deg = torch.clamp(torch.floor(raw * c) + 1, max=dmax)
        c *= (nnz_target / 2) / float(deg.sum())
    deg = torch.clamp(torch.floor(raw * c) + 1, max=dmax).long()
    src = torch.repeat_interleave(torch.arange(n, device=device), deg)
    dst = torch.randint(0, n, (src.numel(),), generator=g, device=device)
    if community is not None:
        size, p_in = community
        inside = torch.rand(src.numel(), generator=g, device=device) < p_in
        local = torch.div(src, size, rounding_mode="floor") * size + torch.randint(0, size, (src.numel(),), generator=g, device=device)
        dst = torch.where(inside, local.clamp_max(n - 1), dst)
        del inside, local
    keep = src != dst
    src, dst = src[keep], dst[keep]
    key = torch.cat([src * n + dst, dst * n + src])
    del src, dst
    key = torch.unique(key)
    rows = torch.div(key, n, rounding_mode="floor")
    indices = (key - rows * n).to(torch.int32)      # bit pattern of uint32 (n < 2^31 here)
    counts = torch.bincount(rows, minlength=n)
    indptr = torch.zeros(n + 1, dtype=torch.int64, device=device)
    torch.cumsum(counts, 0, out=indptr[1:])
    assert int(indptr[-1]) < 2 ** 32
    return indptr, indices


def stratified_csr_torch(n, nnz_target, seed, device, rows_per_chunk=2_000_000):
    """Directed CSR with sorted, duplicate-free rows for SCALE tests (billions of edges: the sort / unique of the recipe above does not fit
    CUB's 2^31-element limit).  Row i gets a power-law degree d_i and the neighbours lo_j + floor(u_j (hi_j - lo_j)), lo_j = floor(j n / d_i), hi_j = lo_{j+1}, u_j ~ U[0,1): one
    stratified draw per slot, so a row is sorted and unique by construction.  Returns (indptr int64 [n+1], indices int32 [nnz])."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    u = torch.rand(n, generator=g, device=device, dtype=torch.float64).clamp_min(1e-12)
    raw = u.pow(-1.0 / 1.6) - 1.0
    c = 1.0
    dmax = min(n // 4, 100_000)
    for _ in range(30):
        deg = torch.clamp(torch.floor(raw * c) + 1, max=dmax)
        c *= nnz_target / float(deg.sum())
    deg = torch.clamp(torch.floor(raw * c) + 1, max=dmax).long()
    del u, raw
    indptr = torch.zeros(n + 1, dtype=torch.int64, device=device)
    torch.cumsum(deg, 0, out=indptr[1:])
    E = int(indptr[-1])
    indices = torch.empty(E, dtype=torch.int32, device=device)
    for r0 in range(0, n, rows_per_chunk):
        r1 = min(n, r0 + rows_per_chunk)
        d = deg[r0:r1]
        e0, e1 = int(indptr[r0]), int(indptr[r1])
        rows = torch.repeat_interleave(torch.arange(r1 - r0, device=device), d)
        j = torch.arange(e1 - e0, device=device, dtype=torch.int64) - (indptr[r0:r1][rows] - e0)
        dd = d[rows].double()
        lo = torch.floor(j.double() * (n / dd)).long()                       # integer strata [lo_j, hi_j): disjoint because n / d >= 1
        hi = torch.floor((j.double() + 1.0) * (n / dd)).long().clamp_(max=n)
        ids = (lo + torch.floor(torch.rand(e1 - e0, generator=g, device=device, dtype=torch.float64) * (hi - lo).double()).long()).clamp_(max=n - 1)
        del lo, hi
        indices[e0:e1] = ids.to(torch.int32)
        del rows, j, dd, ids
    return indptr, indices


def small_parity_graph(n=2000, avg_deg=12, seed=0, self_loops=0):
    """Small graph for parity tests: power-law + optional pre-existing self loops + tail component."""
    indptr, indices = powerlaw_graph(n, n * avg_deg, seed, dmax=max(8, n // 10), tail=True)
    if self_loops:
        rng = np.random.default_rng(seed + 1)
        loops = rng.choice(n, self_loops, replace=False)
        rows = np.repeat(np.arange(indptr.size - 1), np.diff(indptr.astype(np.int64)))
        key = np.unique(np.concatenate([rows.astype(np.int64) * (n + 2) + indices, loops.astype(np.int64) * (n + 2) + loops]))
        r = key // (n + 2)
        indices = (key % (n + 2)).astype(np.uint32)
        ip = np.zeros(n + 3, np.int64)
        np.cumsum(np.bincount(r, minlength=n + 2), out=ip[1:])
        indptr = ip.astype(np.uint32)
    return indptr, indices
