"""GPU parity tests (-m gpu) of the CUDA layers / model / optimizer / minibatch.
Floating point: tolerance 1e-3 relative (BASELINE.json north_star) against (1) golden outputs of the unmodified reference layers
(tests/golden/layers_golden.npz) with the reference's own state_dict loaded into our modules, (2) the fp32 torch restatement."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "layers_golden.npz"))
RTOL = 1e-3


def close(a, b, what):
    a, b = a.detach().float().cpu(), torch.as_tensor(b).float()
    err = (a - b).abs().max().item()
    scale = b.abs().max().item() + 1e-6
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert err <= RTOL * scale + 1e-5, f"{what}: max abs err {err:.3e} vs scale {scale:.3e}"


def device_adj(self_edge):
    from shadow_gnn_b200.ops import DeviceCSR
    ip, ix = Z[f"adj{int(self_edge)}_indptr"], Z[f"adj{int(self_edge)}_indices"]
    span = torch.as_tensor(np.stack([ip[:-1], ip[1:]], 1).astype(np.int32)).cuda()
    # place the batch behind an offset, as a slice of a sampler super-batch would be
    off = 7
    col = torch.as_tensor((ix + off).astype(np.int32)).cuda()
    return DeviceCSR(span.contiguous(), col, off), torch.as_tensor(Z[f"adj{int(self_edge)}_target"]).cuda(), torch.as_tensor(Z[f"adj{int(self_edge)}_size_subg"]).cuda()


def load_golden_state(module, name):
    pre = f"{name}_p_"
    sd = {k[len(pre):]: torch.tensor(Z[k]) for k in Z.files if k.startswith(pre)}
    module.load_state_dict(sd, strict=True)          # the reference's parameter names / shapes must fit ours exactly
    return module.cuda().eval()


def run_case(name, module, fwd):
    module = load_golden_state(module, name)
    x = torch.tensor(Z[f"{name}_x"], device="cuda", requires_grad=True)
    out = fwd(module, x)
    close(out, Z[f"{name}_out"], f"{name} forward")
    (out * torch.tensor(Z[f"{name}_w"], device="cuda")).sum().backward()
    close(x.grad, Z[f"{name}_dx"], f"{name} d/dx")
    for pn, p in module.named_parameters():
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        close(g, Z[f"{name}_g_{pn}"], f"{name} d/d{pn}")


@pytest.mark.parametrize("name,cls,kw,se", [
    ("sage", "GraphSAGE", dict(act="relu"), False), ("sage_elu", "GraphSAGE", dict(act="elu"), True), ("gcn", "GCN", dict(act="elu"), True),
    ("gin", "GIN", dict(act="relu", eps=0.1), False), ("gat", "GAT", dict(act="relu", mulhead=4), True),
    ("gatscat", "GATScatter", dict(act="relu", mulhead=2), True)])
def test_layer_vs_reference_golden(name, cls, kw, se):
    from shadow_gnn_b200 import layers as L
    adj, _, sizes = device_adj(se)
    run_case(name, getattr(L, cls)(12, 16, **kw), lambda m, x: m((x, adj, False, 0.0), sizes)[0])


def test_stacked_layers_reuse_normalised_adjacency():
    from shadow_gnn_b200 import layers as L
    for name, cls in (("sage2", L.GraphSAGE), ("gin2", L.GIN)):
        adj, _, sizes = device_adj(False)
        run_case(name, torch.nn.Sequential(cls(12, 16), cls(16, 16)), lambda m, x: m[1](m[0]((x, adj, False, 0.0), sizes), sizes)[0])


@pytest.mark.parametrize("res,pool", [("none", "center"), ("cat", "center"), ("max", "max"), ("sum", "mean"), ("cat", "sum"), ("none", "sort")])
def test_respool_vs_reference_golden(res, pool):
    from shadow_gnn_b200 import layers as L
    _, tgt, sizes = device_adj(False)
    rp = L.ResPool(16, 16, 3, res, pool, dropout=0.0, act="relu", args_pool={"k": 5} if pool == "sort" else {})
    if res == "none" and pool == "center":
        x = torch.tensor(Z["respool_none_center_x"], device="cuda")
        close(rp([x[:, :16], x[:, 16:32] * 0.5 + 0.1, x[:, 32:48] - 0.2], tgt, sizes), Z["respool_none_center_out"], "center/none")
        return
    run_case(f"respool_{res}_{pool}", rp, lambda m, x: m([x[:, :16], x[:, 16:32] * 0.5 + 0.1, x[:, 32:48] - 0.2], tgt, sizes))


def _arch(aggr, act, heads):
    return dict(num_layers=3, num_cls_layers=1, heads=heads, branch_sharing=False, dim=16, act=act, layer_norm="norm_feat",
                feature_augment_ops="sum", aggr=aggr, residue="none", pooling="center", loss="softmax", ensemble_act="leakyrelu")


@pytest.mark.parametrize("name,aggr,se,act,heads", [("model_sage", "sage", False, "relu", 1), ("model_gat", "gat", True, "elu", 2)])
def test_deepgnn_vs_reference_golden(name, aggr, se, act, heads):
    from shadow_gnn_b200.models import DeepGNN
    adj, tgt, sizes = device_adj(se)
    model = DeepGNN(12, 12, 5, 0, _arch(aggr, act, heads), [], 1, dict(dropout=0.0, dropedge=0.0, lr=0.01, ensemble_dropout="none"), "node")
    run_case(name, model, lambda m, x: m(1, [x], [adj], [tgt], sizes.view(1, -1), [{}], 0.0)[0])


def test_spmm_and_norms_vs_dense_restatement_random():
    """larger random problem incl. duplicate edges, odd feature widths, dropedge: CUDA ops vs the dense fp32 restatement"""
    from oracle import layers_ref as R
    from shadow_gnn_b200 import ops
    g = torch.Generator().manual_seed(0)
    n, e = 700, 9000
    rows = torch.sort(torch.randint(0, n, (e,), generator=g)).values
    cols = torch.randint(0, n, (e,), generator=g)
    indptr = torch.zeros(n + 1, dtype=torch.int64); indptr[1:] = torch.cumsum(torch.bincount(rows, minlength=n), 0)
    A = R.dense_counts(indptr.numpy(), cols.numpy(), n)
    span = torch.stack([indptr[:-1], indptr[1:]], 1).to(torch.int32).cuda().contiguous()
    for F_ in (100, 256, 37):
        adj = ops.DeviceCSR(span, cols.to(torch.int32).cuda(), 0)
        x = torch.randn(n, F_, generator=g).cuda().requires_grad_(True)
        adj.normalize_rw(0.0)
        y = ops.spmm(adj, x)
        close(y, R.adj_rw(A) @ x.detach().cpu(), f"spmm rw F={F_}")
        w = torch.randn(n, F_, generator=g).cuda()
        (y * w).sum().backward()
        close(x.grad, R.adj_rw(A).T @ w.cpu(), f"spmm^T F={F_}")
        close(adj.to_dense(), R.adj_rw(A), "rw values")
    adj = ops.DeviceCSR(span, cols.to(torch.int32).cuda(), 0).normalize_gin(0.3, seed=5, step=torch.ones(1, dtype=torch.int32, device='cuda'))
    D = adj.to_dense().cpu()
    dropped = (A > 0) & (D == 0)
    assert 0 < dropped.sum() <= int(e * 0.3)                         # draws with replacement: at most int(e*p) distinct edges dropped
    close(D.sum(1), torch.where(D.sum(1) > 0, A.sum(1), torch.zeros(n)), "gin rescale keeps the row sums")
    for act in ("relu", "elu", "tanh", "leakyrelu", "I"):
        z = torch.randn(n, 64, generator=g).cuda().requires_grad_(True)
        sc, of = torch.randn(64, generator=g).cuda().requires_grad_(True), torch.randn(64, generator=g).cuda().requires_grad_(True)
        out = ops.act_norm(z, sc, of, act)
        zc, scc, ofc = (t.detach().cpu().requires_grad_(True) for t in (z, sc, of))
        ref = R.norm_feat(R.ACT[act](zc), scc, ofc)
        close(out, ref, f"act_norm {act}")
        w = torch.randn(n, 64, generator=g)
        (out * w.cuda()).sum().backward(); (ref * w).sum().backward()
        close(z.grad, zc.grad, f"act_norm {act} dz"); close(sc.grad, scc.grad, f"act_norm {act} dscale"); close(of.grad, ofc.grad, f"act_norm {act} doffset")


def test_flat_adam_clip_matches_torch():
    from shadow_gnn_b200.ops import FlatAdamClip
    torch.manual_seed(0)
    m1 = torch.nn.Sequential(torch.nn.Linear(20, 30), torch.nn.Linear(30, 5)).cuda()
    m2 = torch.nn.Sequential(torch.nn.Linear(20, 30), torch.nn.Linear(30, 5)).cuda()
    m2.load_state_dict(m1.state_dict())
    ours = FlatAdamClip(list(m1.parameters()), lr=0.01, max_norm=5.0)
    ref = torch.optim.Adam(m2.parameters(), lr=0.01)
    for it in range(6):
        x = torch.randn(64, 20, device="cuda") * (30.0 if it % 2 else 1.0)       # some steps clip, some do not
        ours.zero_grad(); ref.zero_grad()
        (m1(x) ** 2).sum().backward(); (m2(x) ** 2).sum().backward()
        torch.nn.utils.clip_grad_norm_(m2.parameters(), 5)
        ours.step(); ref.step()
        for a, b in zip(m1.parameters(), m2.parameters()):
            close(a, b.detach().cpu(), f"param after step {it}")


def test_minibatch_epoch_matches_oracle_collation():
    """device MinibatchShallowExtractor: every batch == cat_to_block_diagonal of the oracle's subgraphs for the same targets"""
    from oracle import oracle as O
    from shadow_gnn_b200 import minibatch as MB
    from shadow_gnn_b200.synth import small_parity_graph
    indptr, indices = small_parity_graph(1500, 10, 2)
    N = indptr.size - 1
    feat = torch.randn(N, 20)
    label = torch.randint(0, 5, (N,))
    train = np.arange(0, 230, dtype=np.int64)
    cfg = {"batch_size": 32, "configs": [{"method": "khop", "depth": [2], "budget": [5], "add_self_edge": [True]}]}
    np.random.seed(3)
    mb = MB.MinibatchShallowExtractor("toy", None, {0: (indptr, indices), 1: (indptr, indices), 2: (indptr, indices)},
                                      {0: train, 1: train[:40], 2: train[:40]}, cfg, {"hops"}, None, feat, label, 20, True, 1, seed_cpp=4,
                                      num_subg_per_batch=100)
    mb.epoch_start_reset(0, MB.TRAIN)
    mb.shuffle_entity(MB.TRAIN)
    o = O.OracleSampler(indptr, indices, 96, 1, 4)            # 96 = whole batches per call
    o.shuffle_targets(mb.entity_epoch[MB.TRAIN].astype(np.uint32))
    want = []
    while True:
        want.extend(o.sample(O.make_cfg("khop", depth=2, budget=5, add_self_edge=True, aug=("hops",))).subgraphs())
        if o.get_idx_root() == 0:
            break
    seen = 0
    while not mb.is_end_epoch(MB.TRAIN):
        b = mb.one_batch(MB.TRAIN, ret_raw_idx=True)
        bs = b.target_ens[0].numel()
        col = O.cat_to_block_diagonal(want[seen:seen + bs])
        adj = b.adj_ens[0]
        span = adj.row_span.cpu().numpy().astype(np.int64)
        assert np.array_equal(span[:, 1] - span[:, 0], np.diff(col["indptr"]))
        colidx = np.concatenate([adj.col.cpu().numpy()[s:e] - adj.col_off for s, e in span])
        assert np.array_equal(colidx, col["indices"])
        assert np.array_equal(b.target_ens[0].cpu().numpy(), col["target"])
        assert np.array_equal(b.size_subg_ens[0].cpu().numpy(), col["size_subg"])
        assert np.array_equal(b.idx_raw[0].cpu().numpy().view(np.uint32), col["node"])
        assert torch.equal(b.feat_ens[0].cpu(), feat[col["node"].astype(np.int64)])
        assert torch.equal(b.label.cpu(), label[mb.entity_epoch[MB.TRAIN][seen:seen + bs]])
        hop = np.concatenate([w["hop"] for w in want[seen:seen + bs]]).astype(np.int64)
        oh = b.feat_aug_ens[0]["hops"].cpu().numpy()
        assert oh.shape == (hop.size, 7) and np.array_equal(oh.argmax(1)[hop <= 5], hop[hop <= 5] + 1)
        seen += bs
    assert seen == train.size
    mb.epoch_end_reset(MB.TRAIN)


def test_train_step_decreases_loss_and_is_deterministic_in_eval():
    from shadow_gnn_b200 import minibatch as MB
    from shadow_gnn_b200.models import DeepGNN
    from shadow_gnn_b200.synth import small_parity_graph
    torch.manual_seed(0); np.random.seed(0)
    indptr, indices = small_parity_graph(2000, 12, 9)
    N = indptr.size - 1
    label = torch.randint(0, 4, (N,))
    feat = torch.randn(N, 16) + torch.nn.functional.one_hot(label, 4).float().repeat(1, 4) * 1.5
    train = np.arange(0, 512, dtype=np.int64)
    cfg = {"batch_size": 32, "configs": [{"method": "ppr", "k": [20], "threshold": [0.0], "epsilon": [1e-4]}]}
    mb = MB.MinibatchShallowExtractor("toy", None, {m: (indptr, indices) for m in range(3)}, {0: train, 1: train[:64], 2: train[:64]}, cfg, set(), None,
                                      feat, label, 16, True, 1, seed_cpp=1, num_subg_per_batch=256)
    arch = dict(num_layers=3, num_cls_layers=1, heads=1, branch_sharing=False, dim=32, act="relu", layer_norm="norm_feat", feature_augment_ops="sum",
                aggr="sage", residue="none", pooling="center", loss="softmax", ensemble_act="leakyrelu")
    model = DeepGNN(16, 16, 4, 0, arch, [], 1, dict(dropout=0.1, dropedge=0.05, lr=0.01, ensemble_dropout="none"), "node").cuda()
    losses = []
    for ep in range(3):
        mb.epoch_start_reset(ep, MB.TRAIN); mb.shuffle_entity(MB.TRAIN)
        tot = 0.0
        while not mb.is_end_epoch(MB.TRAIN):
            tot += float(model.step(MB.TRAIN, "running", mb.one_batch(MB.TRAIN))["loss"])
        mb.epoch_end_reset(MB.TRAIN)
        losses.append(tot)
    assert losses[-1] < 0.7 * losses[0], losses
    mb.epoch_start_reset(0, MB.VALID); mb.shuffle_entity(MB.VALID)
    b = mb.one_batch(MB.VALID)
    r1 = model.step(MB.VALID, "running", b)["preds"]
    r2 = model.step(MB.VALID, "running", b)["preds"]
    assert torch.equal(r1, r2)


def test_graphed_trainer_matches_eager_steps():
    """the captured whole-step graph (padded static batch) gives the same parameters as the eager DeepGNN.step (dropout = dropedge = 0)"""
    from shadow_gnn_b200 import minibatch as MB
    from shadow_gnn_b200.models import DeepGNN
    from shadow_gnn_b200.train import GraphedTrainer
    from shadow_gnn_b200.synth import small_parity_graph
    indptr, indices = small_parity_graph(2000, 12, 9)
    N = indptr.size - 1
    torch.manual_seed(0)
    label = torch.randint(0, 4, (N,))
    feat = torch.randn(N, 16)
    train = np.arange(0, 400, dtype=np.int64)            # 12 full batches + a short one (eager fallback)
    cfg = {"batch_size": 32, "configs": [{"method": "ppr", "k": [20], "threshold": [0.0], "epsilon": [1e-4]}]}
    arch = dict(num_layers=2, num_cls_layers=1, heads=1, branch_sharing=False, dim=32, act="relu", layer_norm="norm_feat", feature_augment_ops="sum",
                aggr="sage", residue="none", pooling="center", loss="softmax", ensemble_act="leakyrelu")
    params = []
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
            assert tr.graph_steps == 12 and tr.eager_steps == 1
        params.append((losses, [p.detach().clone() for p in model.parameters()]))
    for a, b in zip(params[0][0], params[1][0]):
        assert abs(a - b) <= 1e-3 * max(abs(a), 1e-3) + 1e-5, (params[0][0], params[1][0])
    for a, b in zip(params[0][1], params[1][1]):
        close(a, b.cpu(), "parameters after one epoch: graphed vs eager")


def test_aug_onehots_match_reference_formulas():
    """hop / ppr / drnl one-hot encodings on device vs the numpy formulas of EntityEncoding (frontend/graph.py:134-172)"""
    from shadow_gnn_b200 import minibatch as MB
    hop = np.array([0, 1, 2, 5, 6, 7, 254, 255, 4294967295], dtype=np.uint32)
    want = np.zeros((hop.size, 7))
    for i in [-1, 0, 1, 2, 3, 4, 5]:
        want[np.where(hop.astype(np.int64) == i)[0], i + 1] = 1
    want[np.where(hop >= 255)[0], 0] = 1
    got = MB.hop2onehot(torch.as_tensor(hop.view(np.int32)).cuda(), 7).cpu().numpy()
    assert np.array_equal(got, want)
    ppr = np.array([-1.0, 0.0, 0.1, 0.25, 0.3, 1.0, 1.5], dtype=np.float32)
    for dim in (1, 3):
        cf = [0.25 ** i for i in range(dim)] + [0]
        want = np.zeros((ppr.size, dim))
        for i in range(dim):
            want[np.where(np.logical_and(ppr <= cf[i], ppr >= cf[i + 1])), i] = 1
        assert np.array_equal(MB.ppr2onehot(torch.as_tensor(ppr).cuda(), dim).cpu().numpy(), want)
    drnl = np.array([0, 1, 7, 25, 26, 200, 255, 4294967295], dtype=np.uint32)
    d = drnl.astype(np.int64).copy(); d[d >= 255] = 0; d[d > 25] = 0
    want = np.zeros((d.size, 26)); want[np.arange(d.size), d] = 1
    assert np.array_equal(MB.drnl2onehot(torch.as_tensor(drnl.view(np.int32)).cuda(), 26).cpu().numpy(), want)


@pytest.mark.parametrize("M,N,K", [(4832, 256, 256), (4832, 256, 100), (77, 47, 256), (1, 1, 1), (130, 70, 33), (64, 64, 32), (65, 129, 257)])
def test_tensor_core_linear_products_match_fp64(M, N, K):
    """shadow_gemm_tf32x3_f32 (csrc/gemm.cu): the three products of nn.Linear -- forward x W^T + b, dgrad dZ W, wgrad dZ^T x (split along
    the batch, atomics) -- are fp32-accurate (3xTF32 error compensation), checked against an fp64 product.  Tolerance: 2e-5 of the
    row-by-column magnitude (one plain TF32 pass would be ~5e-4)."""
    from shadow_gnn_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    x = torch.randn(M, K, device="cuda", generator=g)
    w = torch.randn(N, K, device="cuda", generator=g)
    b = torch.randn(N, device="cuda", generator=g)
    dz = torch.randn(M, N, device="cuda", generator=g)

    def check(got, want64, a64, b64, what):
        scale = (a64.abs() @ b64.abs()).clamp_min(1e-30)
        err = ((got.double() - want64).abs() / scale).max().item()
        assert err < 2e-5, f"{what}: scaled error {err:.3e}"
    xd, wd, dzd = x.double(), w.double(), dz.double()
    check(ops.gemm(x, w, bias=b) - b, xd @ wd.t(), xd, wd.t(), "forward")
    check(ops.gemm(x, w), xd @ wd.t(), xd, wd.t(), "forward (no bias)")
    check(ops.gemm(dz, w, b_kn=True), dzd @ wd, dzd, wd, "dgrad")
    for split in (1, max(1, M // 160), 7):
        gacc = torch.zeros(N, K, device="cuda")
        ops.gemm(dz, x, trans_a=True, b_kn=True, out=gacc, accumulate=True, split_k=split)
        check(gacc, dzd.t() @ xd, dzd.t().abs(), xd.abs(), f"wgrad split {split}")


@pytest.mark.parametrize("M,N,K", [(4832, 256, 256), (4832, 256, 100), (672, 256, 256), (33, 64, 128), (1024, 48, 256)])
def test_tcgen05_linear_products_match_fp64(M, N, K):
    """csrc/gemm_umma.cu (tcgen05 + TMEM + TMA, fp32 emulated by 9 bf16 products): forward incl. the bias (C operand with row stride 0),
    dgrad and the batched wgrad slices against fp64; the emulation keeps ~fp32 accuracy (tolerance 2e-5 of the |a|.|b| magnitude)"""
    import ctypes as C
    from shadow_gnn_b200._lib import lib, check as chk
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    x = torch.randn(M, K, device="cuda", generator=g)
    w = torch.randn(N, K, device="cuda", generator=g)
    b = torch.randn(N, device="cuda", generator=g)
    dz = torch.randn(M, N, device="cuda", generator=g)

    def close(got, want64, mag, what):
        err = ((got.double() - want64).abs() / mag.clamp_min(1e-30)).max().item()
        assert err < 2e-5, f"{what}: scaled error {err:.3e}"
    xd, wd, dzd = x.double(), w.double(), dz.double()
    Z = torch.full((M, N), float("nan"), device="cuda")
    chk(lib.shadow_linear_umma_fwd_f32(p(x), K, p(w), K, p(b), p(Z), N, M, N, K, st))
    close(Z, xd @ wd.t() + b.double(), xd.abs() @ wd.abs().t() + b.double().abs(), "forward + bias")
    chk(lib.shadow_linear_umma_fwd_f32(p(x), K, p(w), K, None, p(Z), N, M, N, K, st))
    close(Z, xd @ wd.t(), xd.abs() @ wd.abs().t(), "forward")
    dX = torch.full((M, K), float("nan"), device="cuda")
    chk(lib.shadow_linear_umma_dgrad_f32(p(dz), N, p(w), K, p(dX), K, M, K, N, st))
    close(dX, dzd @ wd, dzd.abs() @ wd.abs(), "dgrad")
    if M % 16 == 0:
        part = torch.full((16, N, K), float("nan"), device="cuda")
        chk(lib.shadow_linear_umma_wgrad_f32(p(dz), N, p(x), K, p(part), M // 16, 16, N, K, st))
        close(part.sum(0), dzd.t() @ xd, dzd.abs().t() @ xd.abs(), "wgrad")
    torch.cuda.synchronize()


def test_tensor_core_linear_pair_launch_matches_single_launches():
    """the two-problems-per-launch variant (GraphSAGE self / neighbour branch) gives bit-identical results to two single launches
    for forward and dgrad, and the same sums (up to atomics order) for wgrad"""
    from shadow_gnn_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(5)
    M, N, K = 1000, 256, 100
    x0, x1 = torch.randn(M, K, device="cuda", generator=g), torch.randn(M, K, device="cuda", generator=g)
    w0, w1 = torch.randn(N, K, device="cuda", generator=g), torch.randn(N, K, device="cuda", generator=g)
    b0, b1 = torch.randn(N, device="cuda", generator=g), torch.randn(N, device="cuda", generator=g)
    d0, d1 = torch.randn(M, N, device="cuda", generator=g), torch.randn(M, N, device="cuda", generator=g)
    old = ops._LINEAR
    ops._LINEAR = "tf32x3"
    try:
        Z0, Z1 = ops._linear_fwd_pair(x0, w0, b0, x1, w1, b1)
        assert torch.equal(Z0, ops.gemm(x0, w0, bias=b0)) and torch.equal(Z1, ops.gemm(x1, w1, bias=b1))
        e0, e1 = ops._linear_dgrad_pair(d0, w0, d1, w1)
        assert torch.equal(e0, ops.gemm(d0, w0, b_kn=True)) and torch.equal(e1, ops.gemm(d1, w1, b_kn=True))
        p0, p1 = torch.nn.Parameter(w0.clone()), torch.nn.Parameter(w1.clone())
        ops._accum_wgrad_pair(p0, d0, x0, p1, d1, x1)
        for got, d, x in ((p0.grad, d0, x0), (p1.grad, d1, x1)):
            want = d.double().t() @ x.double()
            assert ((got.double() - want).abs() / (d.double().abs().t() @ x.double().abs())).max().item() < 2e-5
    finally:
        ops._LINEAR = old


@pytest.mark.parametrize("M,N,K,act,norm,nbranch,mode", [(128, 256, 32, "I", False, 1, 0), (4832, 256, 256, "relu", True, 2, 2), (4832, 256, 100, "elu", True, 1, 1),
                                                         (1000, 64, 256, "tanh", True, 1, 0), (77, 48, 36, "leakyrelu", True, 2, 2), (300, 256, 256, "I", False, 2, 0)])
def test_hand_written_tcgen05_linear_with_fused_epilogue_matches_fp64(M, N, K, act, norm, nbranch, mode):
    """csrc/linear_tc.cu (shadow_linear_tc_f32): Z = X W^T + b, out = norm_feat(act(Z)) (layers.py:329-338,474-483) against fp64 torch:
    pre-activation, output in all three output modes (store / add / atomic add of two branches), saved mean and rstd.  Tolerance: 2e-5
    relative to the tensor's max (3xTF32 = fp32-level accuracy; the layer budget is 1e-3)."""
    import torch.nn.functional as Fn
    from shadow_gnn_b200 import ops
    torch.manual_seed(1)
    dev = torch.device("cuda")
    f = {"relu": torch.relu, "I": lambda v: v, "elu": Fn.elu, "tanh": torch.tanh, "leakyrelu": lambda v: Fn.leaky_relu(v, 0.2)}[act]
    X = [torch.randn(M, K, device=dev) for _ in range(nbranch)]
    W = [torch.randn(N, K, device=dev) / K ** 0.5 for _ in range(nbranch)]
    b = [torch.randn(N, device=dev) for _ in range(nbranch)]
    sc = [torch.rand(N, device=dev) + 0.5 for _ in range(nbranch)]
    of = [torch.randn(N, device=dev) for _ in range(nbranch)]
    Zt = [torch.empty(M, N, device=dev) for _ in range(nbranch)]
    mean = [torch.empty(M, device=dev) for _ in range(nbranch)]
    rstd = [torch.empty(M, device=dev) for _ in range(nbranch)]
    base = torch.randn(M, N, device=dev) if mode == 1 else torch.zeros(M, N, device=dev)
    out0 = base.clone()
    outs = [out0] * nbranch if mode == 2 else [out0] + [torch.empty(M, N, device=dev) for _ in range(nbranch - 1)]
    assert ops._tc_ok(N, K, *X, *W, *Zt, *outs)
    ops._linear_tc([(X[i], W[i], b[i], sc[i] if norm else None, of[i] if norm else None, Zt[i], outs[i], mean[i], rstd[i]) for i in range(nbranch)],
                   M, N, K, ops.ACT_ID[act], norm, mode)
    torch.cuda.synchronize()
    rel = lambda a, w: float((a.double() - w).abs().max() / (w.abs().max() + 1e-12))
    total = base.double().clone()
    for i in range(nbranch):
        z = X[i].double() @ W[i].double().t() + b[i].double()
        o = f(z)
        if norm:
            m, v = o.mean(1, keepdim=True), o.var(1, unbiased=False, keepdim=True) + 1e-9
            assert rel(mean[i], m[:, 0]) < 2e-5 and rel(rstd[i], torch.rsqrt(v)[:, 0]) < 2e-5
            o = (o - m) * sc[i].double() * torch.rsqrt(v) + of[i].double()
        assert rel(Zt[i], z) < 2e-5
        if mode == 2:
            total += o
        else:
            assert rel(outs[i], o + (base.double() if (mode == 1 and i == 0) else 0)) < 2e-5
    if mode == 2:
        assert rel(out0, total) < 2e-5


def test_tcgen05_linear_rejects_misaligned_shapes():
    from shadow_gnn_b200 import ops
    x = torch.randn(64, 256, device="cuda")
    assert not ops._tc_ok(47, 256, x) and not ops._tc_ok(256, 30, x) and not ops._tc_ok(512, 256, x)
    assert not ops._tc_ok(256, 256, x[:, 1:])        # not contiguous / not 16-byte aligned


def test_presplit_weight_planes_equal_in_kernel_split_and_follow_checkpoints(monkeypatch):
    """the TF32 planes of the weights that FlatAdamClip refreshes after every optimizer step (ops.WeightPlanes; read by csrc/linear_tc.cu instead
    of splitting every weight tile in every CTA) give bit-identical training to the in-kernel split, and a load_state_dict after the optimizer
    exists is followed by the planes (stale planes would silently compute with the old weights)"""
    from shadow_gnn_b200 import minibatch as MB, ops
    from shadow_gnn_b200.models import DeepGNN
    from shadow_gnn_b200.synth import small_parity_graph
    indptr, indices = small_parity_graph(2000, 12, 9)
    N = indptr.size - 1
    torch.manual_seed(0)
    label = torch.randint(0, 4, (N,))
    feat = torch.randn(N, 16)
    train = np.arange(0, 256, dtype=np.int64)
    cfg = {"batch_size": 32, "configs": [{"method": "ppr", "k": [20], "threshold": [0.0], "epsilon": [1e-4]}]}
    arch = dict(num_layers=3, num_cls_layers=1, heads=1, branch_sharing=False, dim=32, act="relu", layer_norm="norm_feat", feature_augment_ops="sum",
                aggr="sage", residue="none", pooling="center", loss="softmax", ensemble_act="leakyrelu")
    runs = []
    for presplit in ("1", "0"):
        monkeypatch.setenv("SHADOW_LTC_PRESPLIT", presplit)
        torch.manual_seed(1); np.random.seed(1)
        mb = MB.MinibatchShallowExtractor("toy", None, {m: (indptr, indices) for m in range(3)}, {0: train, 1: train[:64], 2: train[:64]}, cfg, set(), None,
                                          feat, label, 16, True, 1, seed_cpp=1, num_subg_per_batch=128)
        model = DeepGNN(16, 16, 4, 0, arch, [], 1, dict(dropout=0.0, dropedge=0.0, lr=0.01, ensemble_dropout="none"), "node").cuda()
        mb.epoch_start_reset(0, MB.TRAIN); mb.shuffle_entity(MB.TRAIN)
        losses = []
        while not mb.is_end_epoch(MB.TRAIN):
            losses.append(model.step(MB.TRAIN, "running", mb.one_batch(MB.TRAIN))["loss"].detach().clone())
        runs.append((torch.stack(losses), [p.detach().clone() for p in model.parameters()], model, mb))
    # the two forward passes are bit-identical; spmm backward and the weight-gradient split accumulate with fp32 atomics (order varies run to
    # run), so losses after the first step and the parameters agree to rounding
    assert torch.equal(runs[0][0][0], runs[1][0][0])
    assert torch.allclose(runs[0][0], runs[1][0], rtol=1e-4, atol=1e-6)
    for a, b in zip(runs[0][1], runs[1][1]):
        close(a, b.cpu(), "parameters: pre-split planes vs in-kernel split")
    # checkpoint restore with live planes: the forward pass must see the restored weights
    monkeypatch.setenv("SHADOW_LTC_PRESPLIT", "1")
    model, mb = runs[0][2], runs[0][3]
    assert model.optimizer.planes is not None and model.optimizer.planes.fresh
    mb.epoch_start_reset(0, MB.VALID); mb.shuffle_entity(MB.VALID)
    b = mb.one_batch(MB.VALID)
    before = model.step(MB.VALID, "running", b)["preds"].clone()
    sd = {k: (v * 0.5 if v.dim() == 2 else v.clone()) for k, v in model.state_dict().items()}
    model.load_state_dict(sd)
    assert not model.optimizer.planes.fresh
    after = model.step(MB.VALID, "running", b)["preds"]
    monkeypatch.setenv("SHADOW_LTC_PRESPLIT", "0")
    want = model.step(MB.VALID, "running", b)["preds"]
    assert torch.equal(after, want) and not torch.equal(after, before)


@pytest.mark.parametrize("M,N_out,K_in,two", [(128, 256, 256, False), (4832, 256, 256, True), (4832, 256, 100, True), (1000, 64, 32, False), (77, 48, 36, True)])
def test_tcgen05_weight_gradient_matches_fp64_and_is_deterministic(M, N_out, K_in, two):
    """csrc/linear_tc.cu wgrad_tc_kernel + wgrad_finish_kernel: grad += dZ^T X (layers.py:421,451-452 backward) against fp64 torch, accumulated on
    top of an existing gradient; two runs give bit-identical results (slices are summed in slice order, no atomics)."""
    from shadow_gnn_b200 import ops
    torch.manual_seed(2)
    dev = torch.device("cuda")
    k = 2 if two else 1
    dZ = [torch.randn(M, N_out, device=dev) for _ in range(k)]
    X = [torch.randn(M, K_in, device=dev) for _ in range(k)]
    results = []
    for _ in range(2):
        W = [torch.nn.Parameter(torch.zeros(N_out, K_in, device=dev)) for _ in range(k)]
        base = [torch.randn(N_out, K_in, device=dev) for _ in range(k)]
        for w, b in zip(W, base):
            w.grad = b.clone()
        assert ops._wgrad_tc(W, dZ, X)
        torch.cuda.synchronize()
        results.append([w.grad.clone() for w in W])
    for i in range(k):
        want = base[i].double() * 0 + dZ[i].double().t() @ X[i].double()
        got = results[1][i].double() - base[i].double()
        assert float((got - want).abs().max() / want.abs().max()) < 2e-5
    base0 = None
    # determinism: same inputs, same bits (the base gradients differ between the two runs, so compare the increments of a third and fourth run)
    incs = []
    for _ in range(2):
        W = [torch.nn.Parameter(torch.zeros(N_out, K_in, device=dev)) for _ in range(k)]
        for w in W:
            w.grad = torch.zeros_like(w)
        assert ops._wgrad_tc(W, dZ, X)
        incs.append([w.grad.clone() for w in W])
    for a, b in zip(*incs):
        assert torch.equal(a, b)


@pytest.mark.parametrize("D,H,K,act", [(256, 4, 256, "relu"), (256, 4, 100, "elu"), (64, 4, 32, "relu"), (128, 1, 64, "tanh"), (256, 1, 128, "leakyrelu"), (256, 8, 256, "relu")])
def test_fused_gat_layer_equals_composed_path(D, H, K, act, monkeypatch):
    """the fused GAT node (csrc/gat.cu + the tcgen05 Linear: layers.py:607-629 in 4 forward / 6 backward launches) against the composed path
    (separate Linear, logits, aggregation, per-head norm ops) that the reference goldens pin at small dims: output, input gradient and every
    parameter gradient, on the golden batch (self edges, PS.cpp:401 bug edges) with some edges dropped"""
    from shadow_gnn_b200 import layers as L
    torch.manual_seed(3)
    adj, _, _ = device_adj(True)
    n = adj.n
    x0 = torch.randn(n, K, device="cuda")
    res = []
    for fused in ("0", "1"):
        monkeypatch.setenv("SHADOW_GAT_FUSED", fused)
        torch.manual_seed(4)
        layer = L.GAT(K, D, dropout=0.0, act=act, norm="norm_feat", mulhead=H).cuda()
        with torch.no_grad():
            layer.scale.uniform_(0.5, 1.5); layer.offset.normal_(); layer.attention.normal_()
        a, _, _ = device_adj(True)
        a.val = torch.ones(a.col.numel(), device="cuda")
        a.val[::7] = 0.0                                   # dropped edges
        a.normed = "gat"
        x = x0.clone().requires_grad_(True)
        out = layer((x, a, True, 0.0), None)[0]
        (out * torch.linspace(0.5, 1.5, out.numel(), device="cuda").view_as(out)).sum().backward()
        res.append((out.detach(), x.grad.detach(), {k: p.grad.detach().clone() for k, p in layer.named_parameters()}))
    close(res[1][0], res[0][0].cpu(), "fused GAT output")
    close(res[1][1], res[0][1].cpu(), "fused GAT input gradient")
    for k in res[0][2]:
        close(res[1][2][k], res[0][2][k].cpu(), f"fused GAT grad of {k}")


def test_link_task_epoch_matches_oracle_and_trains():
    """link prediction through the device minibatch (shaDow/minibatch.py:281-304,373-377; fe/samplers_ensemble.py:184-210): positives + drawn
    negatives, label 1 / 0, two roots per subgraph, include_target_conn off.  Every batch equals the block-diagonal collation of the oracle's
    two-root subgraphs for the same edge order; negatives are non-edges; a link DeepGNN takes training steps on those batches."""
    from oracle import oracle as O
    from shadow_gnn_b200 import minibatch as MB
    from shadow_gnn_b200.models import DeepGNN
    from shadow_gnn_b200.synth import small_parity_graph
    indptr, indices = small_parity_graph(1200, 10, 5)
    N = indptr.size - 1
    rows = np.repeat(np.arange(N), np.diff(indptr.astype(np.int64)))
    und = np.stack([rows, indices.astype(np.int64)], 1)
    und = und[und[:, 0] < und[:, 1]]
    rng = np.random.default_rng(0)
    pos = und[rng.permutation(und.shape[0])[:150]]
    torch.manual_seed(5); np.random.seed(5)
    feat = torch.randn(N, 12)
    es = {0: {"pos": pos}, 1: {"pos": pos[:20], "neg": pos[20:40][:, ::-1].copy()}, 2: {"pos": pos[:20]}}
    cfg = {"batch_size": 16, "configs": [{"method": "khop", "depth": [2], "budget": [4], "add_self_edge": [True], "include_target_conn": [True]}]}
    mb = MB.MinibatchShallowExtractor("toy", None, {m: (indptr, indices) for m in range(3)}, es, cfg, {"drnls"}, None, feat, None, 12, True, 1, seed_cpp=7,
                                      num_subg_per_batch=64)
    assert mb.prediction_task == "link"
    mb.epoch_start_reset(0, MB.TRAIN)
    mb.shuffle_entity(MB.TRAIN)
    edges = mb.entity_epoch[MB.TRAIN]
    lab = mb.label_epoch[MB.TRAIN].cpu().numpy()[:, 0]
    assert edges.shape == (300, 2) and lab.sum() == 150
    key = lambda e: e[:, 0] * N + e[:, 1]
    posk = set(key(pos).tolist()) | set(key(pos[:, ::-1]).tolist())
    negs = edges[lab == 0]
    assert len(set(key(negs).tolist())) == 150 and not (set(key(negs).tolist()) & posk) and np.all(negs[:, 0] != negs[:, 1])
    assert set(key(edges[lab == 1]).tolist()) == set(key(pos).tolist())
    o = O.OracleSampler(indptr, indices, 64, 1, 7)
    o.shuffle_targets(edges.reshape(-1).astype(np.uint32))
    want = []
    while True:
        want.extend(o.sample(O.make_cfg("khop", num_roots=2, depth=2, budget=4, add_self_edge=True, include_target_conn=False, aug=("drnls",))).subgraphs())
        if o.get_idx_root() == 0:
            break
    arch = dict(num_layers=2, num_cls_layers=1, heads=1, branch_sharing=False, dim=32, act="relu", layer_norm="norm_feat", feature_augment_ops="sum",
                aggr="sage", residue="none", pooling="center", loss="sigmoid", ensemble_act="leakyrelu")
    model = DeepGNN(12, 12, 1, 0, arch, [("drnls", mb.get_aug_dim("drnls"))], 1, dict(dropout=0.0, dropedge=0.0, lr=0.01, ensemble_dropout="none"), "link").cuda()
    seen, losses = 0, []
    while not mb.is_end_epoch(MB.TRAIN):
        b = mb.one_batch(MB.TRAIN, ret_raw_idx=True)
        bs = b.label.shape[0]
        col = O.cat_to_block_diagonal(want[seen:seen + bs])
        adj = b.adj_ens[0]
        span = adj.row_span.cpu().numpy().astype(np.int64)
        assert np.array_equal(span[:, 1] - span[:, 0], np.diff(col["indptr"]))
        colidx = np.concatenate([adj.col.cpu().numpy()[s:e] - adj.col_off for s, e in span])
        assert np.array_equal(colidx, col["indices"])
        tgt = b.target_ens[0].cpu().numpy()
        assert tgt.size == 2 * bs and np.array_equal(tgt, col["target"])
        node = b.idx_raw[0].cpu().numpy().view(np.uint32)
        assert np.array_equal(node, col["node"]) and np.array_equal(node[tgt].reshape(-1, 2), edges[seen:seen + bs])
        assert np.array_equal(b.label.cpu().numpy()[:, 0], lab[seen:seen + bs])
        losses.append(float(model.step(MB.TRAIN, "running", b)["loss"].detach()))
        seen += bs
    assert seen == 300 and all(np.isfinite(losses))
    mb.epoch_end_reset(MB.TRAIN)
    mb.epoch_start_reset(0, MB.VALID); mb.shuffle_entity(MB.VALID)          # given negatives are used as they are
    assert mb.entity_epoch[MB.VALID].shape == (40, 2) and int(mb.label_epoch[MB.VALID].sum()) == 20


def test_p2p_exchange_kernels_single_rank():
    """the peer-memory gradient exchange (csrc/layers.cu: p2p_wait_zero_kernel, p2p_reduce_sqnorm_kernel + adam_clip_kernel) with a world of one
    rank -- its own IPC-shareable buffer is the only 'peer' -- must equal the local optimizer step bit for bit over several steps; the
    multi-GPU behaviour (flags, peer reads, replicas identical) is checked by scripts/check_p2p.py under torchrun (profiles/r2_p2p_exchange_check_*.txt)"""
    import ctypes as C
    from shadow_gnn_b200._lib import lib, check
    from shadow_gnn_b200.ops import _RawDev
    dev = torch.device("cuda")
    n = 100_003
    ptrs = []
    for nbytes in (n * 4, 256):
        ptr, h = C.c_void_p(), (C.c_ubyte * 64)()
        check(lib.shadow_p2p_alloc(nbytes, C.byref(ptr), h))
        ptrs.append(ptr.value)
    gp, fp = (C.c_uint64 * 1)(ptrs[0]), (C.c_uint64 * 1)(ptrs[1])
    grad = torch.as_tensor(_RawDev(ptrs[0], n, "<f4"), device=dev)
    state = torch.zeros(516, dtype=torch.int32, device=dev)
    torch.manual_seed(0)
    p0 = torch.randn(n, device=dev)
    pa, pb = p0.clone(), p0.clone()
    ma, va, mb_, vb = (torch.zeros(n, device=dev) for _ in range(4))
    gsum, sa, sb = torch.zeros(n, device=dev), torch.zeros(1, device=dev), torch.zeros(1, device=dev)
    ta, tb = torch.zeros(1, dtype=torch.int32, device=dev), torch.zeros(1, dtype=torch.int32, device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    P = lambda t: C.c_void_p(t.data_ptr())
    for it in range(4):
        check(lib.shadow_p2p_zero_grad_f32(gp, fp, 1, 0, n, P(state), st))
        assert float(grad.abs().max()) == 0.0
        g = torch.randn(n, device=dev) * (10.0 if it == 0 else 0.01)          # clipped on the first step, not afterwards
        grad.copy_(g)
        check(lib.shadow_p2p_adam_clip_step_f32(gp, fp, 1, 0, P(pa), P(gsum), P(ma), P(va), n, 1.0, 5.0, 0.01, 0.9, 0.999, 1e-8, P(ta), P(sa), P(state), st))
        check(lib.shadow_adam_clip_step_f32(P(pb), P(g), P(mb_), P(vb), n, 1.0, 5.0, 0.01, 0.9, 0.999, 1e-8, P(tb), P(sb), st))
        torch.cuda.synchronize()
        assert int(state[2]) == 0 and int(state[0]) == it + 1
        assert torch.equal(gsum, g)
        close(pa, pb.cpu(), f"parameters after step {it}: fused exchange vs local step")
    check(lib.shadow_p2p_free(C.c_void_p(ptrs[0]))); check(lib.shadow_p2p_free(C.c_void_p(ptrs[1])))
