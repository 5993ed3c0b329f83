"""Round-2 GPU parity tests (-m gpu), all through the C ABI: the branches and BASELINE configurations that round 1 left unproven.

  * dropedge in training mode (GCN: sym_survive_kernel; SAGE: normalize_rw with a draw) against the reference's own adj_norm_sym /
    adj_norm_rw run with an injected draw (tests/golden/r2_golden.npz) and against the restatement with the kernel's own draw
  * the `.bin` constructor path (read_array_from_bin, PS.cpp:70-86)
  * the PPR `.bin` cache in both directions against files written by the unmodified reference (PS.cpp:94-231)
  * EnsembleAggregator and a 2-branch DeepGNN against reference goldens
  * C3 at its real width: sampler + 5-layer SAGE-256 (F=100, C=47) forward / loss / every gradient against the reference golden
  * C1: S-arxiv-size khop(2,10)+hops, batch 32, CUDA == C oracle bit for bit;  C2: 3-layer GCN-256 + hops on that batch vs the reference golden
  * an 8,192-subgraph PPR launch on a 300k-node graph against the ORACLE (not against another kernel of ours)
Floating point: 1e-3 relative (BASELINE.json north_star); integers: bit-exact.
"""
import os
import shutil

import numpy as np
import pytest
import torch

from tests.common import Golden, assert_subgraph_equal, det_fill, grad_signature, FIELDS

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
Z = np.load(os.path.join(HERE, "golden", "r2_golden.npz"))
RTOL = 1e-3


def close(a, b, what, rtol=RTOL):
    a, b = torch.as_tensor(a).detach().float().cpu(), torch.as_tensor(b).float()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = (a - b).abs().max().item()
    scale = b.abs().max().item() + 1e-6
    assert err <= rtol * scale + 1e-6, f"{what}: max abs err {err:.3e} vs scale {scale:.3e}"
    return err / scale


def _features(n, F, seed):
    return torch.from_numpy(np.random.default_rng(seed).standard_normal((n, F), dtype=np.float32))


def _device_csr(indptr, indices, off=0):
    from shadow_gnn_b200.ops import DeviceCSR
    span = torch.as_tensor(np.stack([indptr[:-1], indptr[1:]], 1).astype(np.int32)).cuda().contiguous()
    return DeviceCSR(span, torch.as_tensor((np.asarray(indices) + off).astype(np.int32)).cuda(), off)


def _vals_in_csr_order(adj):
    span = adj.row_span.cpu().numpy()
    v = adj.val.cpu().numpy()
    return np.concatenate([v[s:e] for s, e in span]) if len(span) else v[:0]


# ------------------------------------------------------------------------------------------------ dropedge
def test_sym_and_rw_dropedge_vs_reference_with_injected_draw():
    """the reference's adj_norm_sym / adj_norm_rw were run with drop_idx as their draw (make_r2_golden.py); here the same draw is written
    into the mask and the normalisation kernels (sym_survive_kernel + sym_scale_kernel, row_normalize_kernel) must reproduce the values"""
    import ctypes as C
    from shadow_gnn_b200._lib import lib, check
    ip, ix, drop = Z["drop_indptr"], Z["drop_indices"], Z["drop_idx"]
    n = ip.size - 1
    adj = _device_csr(ip, ix, off=5)
    p = lambda t: C.c_void_p(t.data_ptr())
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    mask = torch.ones(ix.size, device="cuda")
    mask[torch.as_tensor(drop).cuda()] = 0
    val = torch.empty_like(mask)
    deg = torch.empty(n, device="cuda")
    check(lib.shadow_edge_vals_sym_normalize(p(adj.row_span), p(adj.col), adj.col_off, n, 1, p(mask), p(val), p(deg), st))
    close(val, Z["drop_sym_vals"], "adj_norm_sym with dropedge", rtol=1e-5)
    assert torch.equal(val == 0, torch.as_tensor(Z["drop_sym_vals"] == 0).cuda())
    val2 = mask.clone()
    check(lib.shadow_edge_vals_row_normalize(p(adj.row_span), n, 0, p(val2), st))
    close(val2, Z["drop_rw_vals"], "adj_norm_rw with dropedge", rtol=1e-6)


@pytest.mark.parametrize("kind", ["gcn", "sage"])
def test_train_mode_layer_with_dropedge_vs_restatement(kind):
    """GCN / GraphSAGE forward + backward in TRAIN mode (dropedge 0.3, dropout 0): the layer's own draw is read back from the device and fed
    to the (golden-pinned) restatement of the reference's normalisation"""
    from oracle import layers_ref as R
    from shadow_gnn_b200 import layers as L
    ip, ix = Z["drop_indptr"], Z["drop_indices"]
    n = ip.size - 1
    adj = _device_csr(ip, ix, off=3)
    layer = det_fill((L.GCN if kind == "gcn" else L.GraphSAGE)(24, 32, dropout=0.0, act="elu")).cuda().train()
    x = _features(n, 24, 70).cuda().requires_grad_(True)
    out = layer((x, adj, False, 0.3), None)[0]
    vals = _vals_in_csr_order(adj)
    # the draw, recovered from the device: dropped positions (sym: a position can also die because its mirror was drawn, so recover the draw
    # from a second handle that only masks)
    adj_m = _device_csr(ip, ix, off=3)
    seed, step = L._Dropedge.next(x.device)                  # the layer used the previous step value
    step_prev = step.clone() - 1
    adj_m.mask_only(0.3, seed, step_prev)
    drawn = np.nonzero(_vals_in_csr_order(adj_m) == 0)[0]
    assert 0 < drawn.size <= int(ix.size * 0.3)
    want_vals = R.sym_vals_dropedge(ip, ix, drawn) if kind == "gcn" else R.rw_vals_dropedge(ip, drawn)
    assert np.allclose(vals, want_vals, rtol=1e-5, atol=1e-7), np.abs(vals - want_vals).max()
    # layer output / gradients with that adjacency
    rows = np.repeat(np.arange(n), np.diff(ip))
    A = torch.zeros(n, n); A.index_put_((torch.as_tensor(rows), torch.as_tensor(ix)), torch.as_tensor(want_vals), accumulate=True)
    pr = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in layer.named_parameters()}
    xc = x.detach().cpu().clone().requires_grad_(True)
    if kind == "gcn":
        ref = R.norm_feat(R.ACT["elu"]((A @ xc) @ pr["f_lin.weight"].T + pr["f_lin.bias"]), pr["scale"][0], pr["offset"][0])
    else:
        hs = R.ACT["elu"](xc @ pr["f_lin_self.weight"].T + pr["f_lin_self.bias"])
        hn = R.ACT["elu"]((A @ xc) @ pr["f_lin_neigh.weight"].T + pr["f_lin_neigh.bias"])
        ref = R.norm_feat(hs, pr["scale"][0], pr["offset"][0]) + R.norm_feat(hn, pr["scale"][1], pr["offset"][1])
    close(out, ref.detach(), f"{kind} train-mode forward")
    w = _features(n, 32, 71)
    (out * w.cuda()).sum().backward(); (ref * w).sum().backward()
    close(x.grad, xc.grad, f"{kind} train-mode dx")
    for k, v in layer.named_parameters():
        close(v.grad, pr[k].grad, f"{kind} train-mode d{k}")


# ------------------------------------------------------------------------------------------------ .bin paths
def test_bin_constructor_path_matches_array_constructor(tmp_path):
    """ParallelSampler([], [], [], ..., path_indptr, path_indices, ...) (PS.h:42-47, read_array_from_bin PS.cpp:70-86) == golden case"""
    import shadow_gnn_b200.ParallelSampler as PS
    G = Golden()
    fi, fx = str(tmp_path / "indptr.bin"), str(tmp_path / "indices.bin")
    G.indptr.astype(np.uint32).tofile(fi); G.indices.astype(np.uint32).tofile(fx)
    for ci in (1, 10):
        cfg, aug, targets, calls, want = G.case(ci)
        s = PS.ParallelSampler([], [], [], G.meta["P"], 1, True, True, [], 1, fi, fx, "", G.meta["seed"])
        assert s.num_nodes() == G.indptr.size - 1 and s.num_edges() == G.indices.size
        s.shuffle_targets(targets)
        if cfg["method"] in ("ppr", "ppr_st"):
            s.set_ppr_tables(G.ppr_ptr, G.ppr_neighs, G.ppr_scores)
        got = []
        for ncall in calls:
            vec = s.parallel_sampler_ensemble([cfg], [set(aug)])[0]
            got.extend({f: vec.numpy(f)[p] for f in FIELDS} for p in range(vec.get_num_valid_subg()))
        for i, (a, b) in enumerate(zip(want, got)):
            assert_subgraph_equal(a, b, f".bin ctor case {ci} subgraph {i}")
    with pytest.raises(Exception):
        PS.ParallelSampler([], [], [], 4, 1, True, True, [], 1, str(tmp_path / "missing.bin"), fx, "", 1)


def test_ppr_cache_files_interchange_with_the_reference(tmp_path):
    """(a) files written by the unmodified reference are LOADED by the CUDA path (a sentinel planted in a copy shows up in the table);
    (b) files written by the CUDA path are byte-identical to the reference-written ones; (c) where oracle/_ref travelled along, the
    reference loads the CUDA-written files and samples the same PPR subgraphs from them"""
    import shadow_gnn_b200.ParallelSampler as PS
    from oracle import oracle as O
    G = Golden()
    N, p = G.indptr.size - 1, G.meta["ppr"]
    ref_n, ref_s = os.path.join(HERE, "golden", "ppr_ref_neighs.bin"), os.path.join(HERE, "golden", "ppr_ref_scores.bin")
    alln = np.arange(N, dtype=np.uint32)
    # (a)
    fn, fs = str(tmp_path / "n.bin"), str(tmp_path / "s.bin")
    shutil.copy(ref_n, fn); shutil.copy(ref_s, fs)
    raw = np.fromfile(fs, np.uint32)
    first_len = int(raw[4])
    assert first_len >= 2
    sentinel = np.float32(0.123456)
    raw[5 + first_len - 1] = sentinel.view(np.uint32)            # last score of node 0's row (keeps the row sorted enough: no check on load)
    raw.tofile(fs)
    s = PS.ParallelSampler(G.indptr, G.indices, [], 4, 1, True, True, [], 1, "", "", "", 1)
    s.preproc_ppr_approximate(alln, p["k"], p["alpha"], p["epsilon"], fn, fs)
    nb, sc = s.get_ppr_row(0)
    assert sc[first_len - 1] == sentinel, "the cache file was not used"
    for v in range(1, N):
        nb, sc = s.get_ppr_row(v)
        a, b = int(G.ppr_ptr[v]), int(G.ppr_ptr[v + 1])
        assert np.array_equal(nb, G.ppr_neighs[a:b]) and sc.tobytes() == G.ppr_scores[a:b].tobytes(), v
    # (b)
    wn, ws = str(tmp_path / "wn.bin"), str(tmp_path / "ws.bin")
    s2 = PS.ParallelSampler(G.indptr, G.indices, [], 4, 1, True, True, [], 1, "", "", "", 1)
    s2.preproc_ppr_approximate(alln, p["k"], p["alpha"], p["epsilon"], wn, ws)
    assert open(wn, "rb").read() == open(ref_n, "rb").read(), "neighbour cache differs from the reference-written file"
    assert open(ws, "rb").read() == open(ref_s, "rb").read(), "score cache differs from the reference-written file"
    # (c)
    ref = O.load_ref()
    if ref is not None:
        cfg = dict(method="ppr", k="30", threshold="0.002", num_roots="1", add_self_edge="true", include_target_conn="false")
        t = np.random.default_rng(4).permutation(N - 2)[:48].astype(np.uint32)
        r = ref.ParallelSampler(G.indptr.tolist(), G.indices.tolist(), [], 16, 1, True, True, [], 1, "", "", "", 1)
        devnull = os.open(os.devnull, os.O_WRONLY); saved = os.dup(1); os.dup2(devnull, 1)
        try:
            r.preproc_ppr_approximate([], 30, p["alpha"], p["epsilon"], wn, ws)          # no targets: everything must come from the files
        finally:
            os.dup2(saved, 1); os.close(devnull)
        r.shuffle_targets(t.tolist())
        s2.shuffle_targets(t)
        s2.set_num_sampler_per_batch(16)
        for _ in range(3):
            want = O.ref_subgraphs(r.parallel_sampler_ensemble([cfg], [set()])[0])
            vec = s2.parallel_sampler_ensemble([cfg], [set()])[0]
            assert vec.get_num_valid_subg() == len(want) == 16
            for i, w in enumerate(want):
                assert_subgraph_equal(w, {f: vec.numpy(f)[i] for f in FIELDS}, f"ref on CUDA-written cache, subgraph {i}")


# ------------------------------------------------------------------------------------------------ ensemble
def test_ensemble_aggregator_vs_reference_golden():
    from shadow_gnn_b200 import layers as L
    ens = det_fill(L.EnsembleAggregator(16, 16, 3, dropout=0.0, act="leakyrelu", type_dropout="none")).cuda().eval()
    Xs = [_features(9, 16, 30 + i).cuda().requires_grad_(True) for i in range(3)]
    y = ens(list(Xs))
    close(y, Z["ens_out"], "EnsembleAggregator forward")
    (y * _features(9, 16, 40).cuda()).sum().backward()
    for i, X in enumerate(Xs):
        close(X.grad, Z[f"ens_dx{i}"], f"EnsembleAggregator dX{i}")
    for k, v in ens.named_parameters():
        close(v.grad, Z[f"ens_g_{k}"], f"EnsembleAggregator d{k}")


def _check_signature(tag, named_grads):
    worst = 0.0
    for k, v in grad_signature(named_grads).items():
        w = Z[f"{tag}_g|{k}"]
        scale = float(Z[f"{tag}_g|{k.split('|')[0]}|norm"]) + 1e-12
        err = float(np.abs(np.asarray(v, np.float64) - w).max())
        assert err <= 2e-3 * scale + 1e-7, (tag, k, err, scale)
        worst = max(worst, err / scale)
    return worst


def test_two_branch_deepgnn_vs_reference_golden():
    from shadow_gnn_b200.models import DeepGNN
    arch = dict(num_layers=2, num_cls_layers=1, heads=1, branch_sharing=False, dim=16, act="relu", layer_norm="norm_feat", feature_augment_ops="sum",
                aggr="sage", residue="none", pooling="center", loss="softmax", ensemble_act="leakyrelu")
    m = det_fill(DeepGNN(12, 12, 5, 0, arch, [], 2, dict(dropout=0.0, dropedge=0.0, lr=0.01, ensemble_dropout="none"), "node")).cuda().eval()
    xs, adjs, tg, sizes = [], [], [], []
    for b in range(2):
        ip, ix = Z[f"ens2_b{b}_indptr"], Z[f"ens2_b{b}_indices"]
        xs.append(_features(ip.size - 1, 12, 50 + b).cuda())
        adjs.append(_device_csr(ip, ix, off=11 * b))
        tg.append(torch.as_tensor(Z[f"ens2_b{b}_target"]).cuda()); sizes.append(torch.as_tensor(Z[f"ens2_b{b}_size_subg"]).cuda())
    preds, _ = m(0, xs, adjs, tg, torch.stack(sizes, 0), [{}, {}], 0.0)
    close(preds, Z["ens2_preds"], "2-branch DeepGNN predictions")
    (preds * _features(6, 5, 60).cuda()).sum().backward()
    _check_signature("ens2", [(k, (v.grad if v.grad is not None else torch.zeros_like(v)).cpu().numpy()) for k, v in m.named_parameters()])


# ------------------------------------------------------------------------------------------------ C3 at width
ARCH5 = dict(num_layers=5, num_cls_layers=1, heads=1, branch_sharing=False, dim=256, act="relu", layer_norm="norm_feat", feature_augment_ops="sum",
             aggr="sage", residue="none", pooling="center", loss="softmax", ensemble_act="leakyrelu")


def _sample_block(s, cfg, aug=()):
    b = s.sample_to_device([cfg], [set(aug)])[0]
    return b


def test_c3_sampler_and_sage5x256_vs_reference_golden(capsys):
    """BASELINE config C3 at its real width (5 x 256, F=100, C=47, PPR k=150, batch 32): CUDA sampler == golden batch bit for bit, then the
    model's predictions, loss and every gradient against the unmodified reference (3xTF32 error budget over 5 layers, reported)"""
    import shadow_gnn_b200.ParallelSampler as PS
    from oracle import oracle as O
    from shadow_gnn_b200.models import DeepGNN
    from shadow_gnn_b200.ops import DeviceCSR
    from tests.golden.make_r2_inputs import sage5_inputs, SAGE5_CFG
    indptr, indices, targets, tables = sage5_inputs(O)
    s = PS.ParallelSampler(indptr, indices, [], 32, 1, True, True, [], 1, "", "", "", 1)
    s.set_ppr_tables(*tables); s.shuffle_targets(targets)
    b = _sample_block(s, SAGE5_CFG)
    assert np.array_equal(b.rowptr.cpu().numpy(), Z["sage5_indptr"]) and np.array_equal(b.indices.cpu().numpy(), Z["sage5_indices"])
    assert np.array_equal(b.orig_node.cpu().numpy().view(np.uint32), Z["sage5_node"]) and np.array_equal(b.target.cpu().numpy().ravel(), Z["sage5_target"])
    n = b.total_nodes
    model = det_fill(DeepGNN(100, 100, 47, 0, ARCH5, [], 1, dict(dropout=0.0, dropedge=0.0, lr=0.002, ensemble_dropout="none"), "node")).cuda().eval()
    x = _features(n, 100, 21).cuda().requires_grad_(True)
    adj = DeviceCSR(b.row_span, b.indices_raw, 0)
    sizes = torch.diff(b.node_ptr).long().view(1, -1)
    preds, _ = model(0, [x], [adj], [b.target.long().ravel()], sizes, [{}], 0.0)
    e_fwd = close(preds, Z["sage5_preds"], "5x256 SAGE predictions")
    loss = model._loss(preds, torch.as_tensor(Z["sage5_labels"]).cuda().long())
    assert abs(loss.item() - float(Z["sage5_loss"])) <= RTOL * float(Z["sage5_loss"])
    loss.backward()
    e_bwd = _check_signature("sage5", [(k, (v.grad if v.grad is not None else torch.zeros_like(v)).cpu().numpy()) for k, v in model.named_parameters()] +
                             [("input_x", x.grad.cpu().numpy())])
    with capsys.disabled():
        print(f"\n[5x256 SAGE vs reference] max rel err: predictions {e_fwd:.2e}, gradient signatures {e_bwd:.2e}")


# ------------------------------------------------------------------------------------------------ C1 / C2
def test_c1_arxiv_size_khop_hops_vs_oracle():
    """BASELINE config C1: S-arxiv (N=169,343, nnz 2.32 M), khop(depth 2, budget 10) + hops, batch 32, glibc replay: CUDA == C oracle, every
    field of every subgraph, over 12 consecutive calls (the rand() stream position carries over) -- with and without the inserted self edge"""
    from tests.golden.make_r2_inputs import gcn3_inputs
    from tests.test_sampler_gpu import _oracle_vs_cuda
    indptr, indices, _ = gcn3_inputs()
    N = indptr.size - 1
    for se in ("false", "true"):
        cfg = dict(method="khop", depth="2", budget="10", num_roots="1", add_self_edge=se, include_target_conn="false")
        t = np.random.default_rng(9).permutation(N)[:384].astype(np.uint32)
        assert _oracle_vs_cuda(indptr, indices, t, 32, 1, cfg, ("hops",)) == 384


def test_c2_arxiv_gcn3_hops_vs_reference_golden(capsys):
    """BASELINE config C2: the first 32-root khop(2,10)+hops batch of S-arxiv through the CUDA sampler (== golden batch), hop one-hot on device,
    3-layer GCN-256 (elu) with the hops embedding summed into the features: predictions / loss / gradients vs the unmodified reference"""
    import shadow_gnn_b200.ParallelSampler as PS
    from shadow_gnn_b200.minibatch import hop2onehot
    from shadow_gnn_b200.models import DeepGNN
    from shadow_gnn_b200.ops import DeviceCSR
    from tests.golden.make_r2_inputs import gcn3_inputs, GCN3_CFG
    indptr, indices, targets = gcn3_inputs()
    s = PS.ParallelSampler(indptr, indices, [], 32, 1, True, True, [], 1, "", "", "", 1)
    s.shuffle_targets(targets)
    b = _sample_block(s, GCN3_CFG, ("hops",))
    assert np.array_equal(b.rowptr.cpu().numpy(), Z["gcn3_indptr"]) and np.array_equal(b.indices.cpu().numpy(), Z["gcn3_indices"])
    assert np.array_equal(b.orig_node.cpu().numpy().view(np.uint32), Z["gcn3_node"])
    assert np.array_equal(b.hop.cpu().numpy().view(np.uint32).astype(np.int64), Z["gcn3_hop"])
    onehot = hop2onehot(b.hop, 7)
    assert np.array_equal(onehot.cpu().numpy(), Z["gcn3_hop_onehot"])
    arch = dict(ARCH5, num_layers=3, aggr="gcn", act="elu")
    model = det_fill(DeepGNN(128, 128, 40, 0, arch, [("hops", 7)], 1, dict(dropout=0.0, dropedge=0.0, lr=0.002, ensemble_dropout="none"), "node")).cuda().eval()
    x = _features(b.total_nodes, 128, 22).cuda()
    adj = DeviceCSR(b.row_span, b.indices_raw, 0)
    preds, _ = model(0, [x], [adj], [b.target.long().ravel()], torch.diff(b.node_ptr).long().view(1, -1), [{"hops": onehot}], 0.0)
    e_fwd = close(preds, Z["gcn3_preds"], "3-layer GCN predictions")
    loss = model._loss(preds, torch.as_tensor(Z["gcn3_labels"]).cuda().long())
    assert abs(loss.item() - float(Z["gcn3_loss"])) <= RTOL * float(Z["gcn3_loss"])
    loss.backward()
    e_bwd = _check_signature("gcn3", [(k, (v.grad if v.grad is not None else torch.zeros_like(v)).cpu().numpy()) for k, v in model.named_parameters()])
    with capsys.disabled():
        print(f"\n[3-layer GCN-256 + hops vs reference] max rel err: predictions {e_fwd:.2e}, gradient signatures {e_bwd:.2e}")


# ------------------------------------------------------------------------------------------------ large launch vs the oracle
@pytest.mark.parametrize("se", ["false", "true"])
def test_superbatch_8192_ppr_vs_oracle(se):
    """8,192 subgraphs per launch on a 300k-node graph with hub rows (S-products-like, PPR k=150): the one-warp fast path (+ redo hand-over)
    against the C ORACLE (OpenMP), every array of the block-diagonal batch, bit for bit"""
    import shadow_gnn_b200.ParallelSampler as PS
    from oracle import oracle as O
    from shadow_gnn_b200.synth import powerlaw_graph
    indptr, indices = powerlaw_graph(300_000, 9_000_000, 5, dmax=6000)
    N, P = indptr.size - 1, 8192
    t = np.random.default_rng(1).permutation(N)[:P].astype(np.uint32)
    threads = os.cpu_count() or 1
    nb, sc, ln = O.ppr_push(indptr, indices, t, 150, 0.85, 1e-5, threads)
    tables = O.ppr_rows_to_csr(N, t, nb, sc, ln)
    cfg = dict(method="ppr", k="150", threshold="0", num_roots="1", add_self_edge=se, include_target_conn="false")
    o = O.OracleSampler(indptr, indices, P, threads, 1)
    o.set_ppr(*tables); o.shuffle_targets(t)
    w = o.sample(O.cfg_from_cpp_config(cfg))
    s = PS.ParallelSampler(indptr, indices, [], P, 1, True, True, [], 1, "", "", "", 1)
    s.set_ppr_tables(*tables); s.shuffle_targets(t)
    b = s.sample_to_device([cfg], [set()])[0]
    assert b.num_subg == P == w.num_subg
    node_ptr, edge_ptr = b.node_ptr.cpu().numpy().astype(np.int64), b.edge_ptr.cpu().numpy().astype(np.int64)
    assert np.array_equal(node_ptr, w.node_ptr) and np.array_equal(edge_ptr, w.edge_ptr)
    assert np.array_equal(b.orig_node.cpu().numpy().view(np.uint32), w.node)
    assert np.array_equal(b.orig_edge.cpu().numpy().view(np.uint32), w.edge_index)
    nn_, ne_ = np.diff(node_ptr), np.diff(edge_ptr)
    assert np.array_equal(b.indices.cpu().numpy().astype(np.int64) - np.repeat(node_ptr[:-1], ne_), w.indices.astype(np.int64))
    rowptr = b.rowptr.cpu().numpy().astype(np.int64)
    keep = np.ones(w.indptr.size, bool); keep[w.indptr_ptr[1:] - 1] = False                 # the oracle stores n+1 entries per subgraph
    assert np.array_equal(rowptr[:-1] - np.repeat(edge_ptr[:-1], nn_), w.indptr[keep].astype(np.int64))
    assert np.array_equal(w.indptr[w.indptr_ptr[1:] - 1].astype(np.int64), ne_)
    assert np.array_equal(b.target.cpu().numpy().ravel().astype(np.int64) - node_ptr[:-1], w.target.astype(np.int64))
    assert b.ppr.cpu().numpy().tobytes() == w.ppr.tobytes()
    assert s.last_redo_count() < P // 4
    assert s.last_sym(), "powerlaw_graph is symmetric with strictly ascending rows: the upper-triangle variant should have run"


# ------------------------------------------------------------------------------------------------ graphed trainer with real pooling segments
@pytest.mark.parametrize("pooling", ["mean", "max"])
def test_graphed_trainer_pooling_matches_eager(pooling):
    """the whole-step CUDA graph loads the batch's real subgraph sizes, so non-center pooling (21 of the 61 reference configs) trains exactly
    like the eager DeepGNN.step (round-1 advisor finding: sizes were hard-coded to 1)"""
    from shadow_gnn_b200 import minibatch as MB
    from shadow_gnn_b200.models import DeepGNN
    from shadow_gnn_b200.train import GraphedTrainer
    from shadow_gnn_b200.synth import small_parity_graph
    indptr, indices = small_parity_graph(2000, 12, 9)
    N = indptr.size - 1
    torch.manual_seed(0)
    label = torch.randint(0, 4, (N,))
    feat = torch.randn(N, 16)
    train = np.arange(0, 256, dtype=np.int64)
    cfg = {"batch_size": 32, "configs": [{"method": "ppr", "k": [20], "threshold": [0.0], "epsilon": [1e-4]}]}
    arch = dict(num_layers=2, num_cls_layers=1, heads=1, branch_sharing=False, dim=32, act="relu", layer_norm="norm_feat", feature_augment_ops="sum",
                aggr="sage", residue="sum", pooling=pooling, loss="softmax", ensemble_act="leakyrelu")
    res = []
    for graphed in (False, True):
        torch.manual_seed(1); np.random.seed(1)
        mb = MB.MinibatchShallowExtractor("toy", None, {m: (indptr, indices) for m in range(3)}, {0: train, 1: train[:64], 2: train[:64]}, cfg, set(), None,
                                          feat, label, 16, True, 1, seed_cpp=1, num_subg_per_batch=128)
        model = DeepGNN(16, 16, 4, 0, arch, [], 1, dict(dropout=0.0, dropedge=0.0, lr=0.01, ensemble_dropout="none"), "node").cuda()
        mb.epoch_start_reset(0, MB.TRAIN); mb.shuffle_entity(MB.TRAIN)
        tr = GraphedTrainer(model, mb, row_cap=32 * 21, edge_cap=32 * 21 * 21) if graphed else None
        losses = []
        while not mb.is_end_epoch(MB.TRAIN):
            losses.append(float(tr.step()) if graphed else float(model.step(MB.TRAIN, "running", mb.one_batch(MB.TRAIN))["loss"].detach()))
        if graphed:
            assert tr.graph_steps == 8 and tr.eager_steps == 0
        res.append((losses, [p.detach().clone() for p in model.parameters()]))
    for a, b in zip(res[0][0], res[1][0]):
        assert abs(a - b) <= 1e-3 * max(abs(a), 1e-3) + 1e-5, (res[0][0], res[1][0])
    for a, b in zip(res[0][1], res[1][1]):
        close(a, b.cpu(), f"parameters after one epoch ({pooling} pooling): graphed vs eager")


@pytest.mark.parametrize("case", ["hops", "two_branches"])
def test_graphed_trainer_aug_and_ensemble_match_eager(case):
    """the whole-step CUDA graph with feature augmentation (every arxiv config: `feature_augment: hops`) and with two ensemble branches
    (EnsembleAggregator, shaDow/layers.py:236-296) trains exactly like the eager DeepGNN.step"""
    from shadow_gnn_b200 import minibatch as MB
    from shadow_gnn_b200.models import DeepGNN
    from shadow_gnn_b200.train import GraphedTrainer
    from shadow_gnn_b200.synth import small_parity_graph
    indptr, indices = small_parity_graph(2000, 12, 9)
    N = indptr.size - 1
    torch.manual_seed(0)
    label = torch.randint(0, 4, (N,))
    feat = torch.randn(N, 16)
    train = np.arange(0, 256, dtype=np.int64)
    if case == "hops":
        cfg = {"batch_size": 32, "configs": [{"method": "khop", "depth": [2], "budget": [5]}]}
        aug, E, aggr = {"hops"}, 1, "gcn"
    else:
        cfg = {"batch_size": 32, "configs": [{"method": "ppr", "k": [20], "threshold": [0.0], "epsilon": [1e-4]}, {"method": "khop", "depth": [1], "budget": [8]}]}
        aug, E, aggr = set(), 2, "sage"
    arch = dict(num_layers=2, num_cls_layers=1, heads=1, branch_sharing=False, dim=32, act="relu", layer_norm="norm_feat", feature_augment_ops="sum",
                aggr=aggr, residue="sum", pooling="center", loss="softmax", ensemble_act="leakyrelu")
    res = []
    for graphed in (False, True):
        torch.manual_seed(1); np.random.seed(1)
        mb = MB.MinibatchShallowExtractor("toy", None, {m: (indptr, indices) for m in range(3)}, {0: train, 1: train[:64], 2: train[:64]}, cfg, aug, None,
                                          feat, label, 16, True, 1, seed_cpp=1, num_subg_per_batch=128)
        augs = [(k, mb.get_aug_dim(k)) for k in sorted(aug)]
        model = DeepGNN(16, 16, 4, 0, arch, augs, E, dict(dropout=0.0, dropedge=0.0, lr=0.01, ensemble_dropout="none"), "node").cuda()
        mb.epoch_start_reset(0, MB.TRAIN); mb.shuffle_entity(MB.TRAIN)
        tr = GraphedTrainer(model, mb, row_cap=32 * 40, edge_cap=32 * 40 * 24) if graphed else None
        losses = []
        while not mb.is_end_epoch(MB.TRAIN):
            losses.append(float(tr.step()) if graphed else float(model.step(MB.TRAIN, "running", mb.one_batch(MB.TRAIN))["loss"].detach()))
        if graphed:
            assert tr.graph_steps == 8 and tr.eager_steps == 0
        res.append((losses, [p.detach().clone() for p in model.parameters()]))
    for a, b in zip(res[0][0], res[1][0]):
        assert abs(a - b) <= 1e-3 * max(abs(a), 1e-3) + 1e-5, (res[0][0], res[1][0])
    for a, b in zip(res[0][1], res[1][1]):
        close(a, b.cpu(), f"parameters after one epoch ({case}): graphed vs eager")
