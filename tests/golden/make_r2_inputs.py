"""Deterministic inputs of the round-2 golden cases (numpy Generator streams only: identical here and on the GPU box).  Shared by
tests/golden/make_r2_golden.py (which runs the unmodified reference on them) and tests/test_parity_r2_gpu.py (which runs the CUDA path)."""
import numpy as np

from shadow_gnn_b200.synth import PRESETS, powerlaw_graph

SAGE5_CFG = dict(method="ppr", k="150", threshold="0", num_roots="1", add_self_edge="false", include_target_conn="false")
GCN3_CFG = dict(method="khop", depth="2", budget="10", num_roots="1", add_self_edge="true", include_target_conn="false")


def sage5_inputs(O):
    """graph / 32 targets / PPR tables (CPU oracle push, k=150, eps=1e-5) of the 5-layer SAGE-256 case"""
    indptr, indices = powerlaw_graph(30000, 900000, 11, dmax=1500, tail=True)
    N = indptr.size - 1
    targets = np.random.default_rng(3).permutation(N - 2)[:32].astype(np.uint32)
    nb, sc, ln = O.ppr_push(indptr, indices, targets, 150, 0.85, 1e-5, 8)
    return indptr, indices, targets, O.ppr_rows_to_csr(N, targets, nb, sc, ln)


def gcn3_inputs():
    """the S-arxiv stand-in (N=169,343, nnz~2.32 M, seed 0) and the first 32 targets"""
    N, nnz, dmax, F, C, ntrain, seed = PRESETS["S-arxiv"]
    indptr, indices = powerlaw_graph(N, nnz, seed, dmax)
    targets = np.random.default_rng(seed).permutation(N)[:32].astype(np.uint32)
    return indptr, indices, targets
