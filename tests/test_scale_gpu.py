"""Scale test (-m gpu): full-graph CSR slots beyond 2^31 (papers100M-scale graphs have nnz ~ 3.2e9 < 2^32: uint32 indptr / edge ids whose top
bit is set, PS.h:27-37 `NodeType = uint32`).  A 2.3-billion-edge synthetic graph is built on the GPU, roots are taken from the top of the id
range (their rows start past slot 2^31), and khop / PPR subgraphs are compared bit for bit with the C oracle run on the host copy."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_full_graph_slots_beyond_2_31_vs_oracle():
    from oracle import oracle as O
    import shadow_gnn_b200.ParallelSampler as PS
    from shadow_gnn_b200.synth import stratified_csr_torch
    dev = torch.device("cuda:0")
    free, _ = torch.cuda.mem_get_info(dev)
    import psutil
    if free < 60e9 or psutil.virtual_memory().available < 40e9:
        pytest.skip("needs ~60 GB of free HBM and ~40 GB of host memory for the 2.3e9-edge graph")
    N, nnz = 40_000_000, 2_300_000_000
    indptr64, indices = stratified_csr_torch(N, nnz, 11, dev)
    E = int(indptr64[-1])
    assert 2 ** 31 < E < 2 ** 32
    indptr32 = indptr64.to(torch.int32)                      # uint32 bit pattern
    first_high = int(torch.searchsorted(indptr64, torch.tensor([2 ** 31 + 1], device=dev)))
    assert first_high < N - 1000
    rng = np.random.default_rng(0)
    roots = (first_high + rng.permutation(N - first_high)[:96]).astype(np.uint32)
    torch.cuda.empty_cache()
    s = PS.ParallelSampler.from_device_csr(indptr32, indices, 96, seed=3)
    s.preproc_ppr_approximate(roots, 40, 0.85, 1e-4, "", "")
    ip_h = indptr64.cpu().numpy().astype(np.uint32)
    ix_h = indices.cpu().numpy().view(np.uint32)
    o = O.OracleSampler(ip_h, ix_h, 96, 8, 3)
    o.preproc_ppr_approximate(roots, 40, 0.85, 1e-4)
    for cfg_cpp, cfg_o, aug in (
            (dict(method="khop", depth="2", budget="10", num_roots="1", add_self_edge="true", include_target_conn="false"),
             O.make_cfg("khop", depth=2, budget=10, add_self_edge=True, aug=("hops",)), {"hops"}),
            (dict(method="ppr", k="40", threshold="0.002", num_roots="1", add_self_edge="false", include_target_conn="false"),
             O.make_cfg("ppr", k=40, threshold=0.002), set())):
        s.shuffle_targets(roots); o.shuffle_targets(roots)
        got = s.sample_to_device([cfg_cpp], [aug])[0]
        want = O.cat_to_block_diagonal(o.sample(cfg_o).subgraphs())
        assert np.array_equal(got.orig_node.cpu().numpy().view(np.uint32), want["node"])
        assert np.array_equal(got.rowptr.cpu().numpy(), want["indptr"])
        assert np.array_equal(got.indices.cpu().numpy(), want["indices"])
        oe = got.orig_edge.cpu().numpy().view(np.uint32)
        assert np.array_equal(oe, want["edge_index"])
        real = oe[oe != 0xFFFFFFFF]
        assert real.size and int(real.max()) > 2 ** 31, "the sampled rows must lie beyond slot 2^31"
