"""Scratch harness for csrc/linear_tc.cu: accuracy vs fp64 and timing vs the warp-level 3xTF32 kernel (the pytest lives in tests/test_layers_gpu.py)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from shadow_gnn_b200._lib import lib, LinearBranch, check
from shadow_gnn_b200 import ops

dev = torch.device("cuda:0")
torch.manual_seed(0)
p = lambda t: None if t is None else t.data_ptr()
st = lambda: torch.cuda.current_stream().cuda_stream


def act_ref(z, act):
    return {0: torch.relu, 1: lambda v: v, 2: torch.nn.functional.elu, 3: torch.tanh, 4: lambda v: torch.nn.functional.leaky_relu(v, 0.2)}[act](z)


def run(M, N, K, act, do_norm, nbranch, out_mode):
    Xs = [torch.randn(M, K, device=dev) for _ in range(nbranch)]
    Ws = [torch.randn(N, K, device=dev) / K ** 0.5 for _ in range(nbranch)]
    bs = [torch.randn(N, device=dev) for _ in range(nbranch)]
    sc = [torch.rand(N, device=dev) + 0.5 for _ in range(nbranch)]
    of = [torch.randn(N, device=dev) for _ in range(nbranch)]
    Zs = [torch.empty(M, N, device=dev) for _ in range(nbranch)]
    mean = [torch.empty(M, device=dev) for _ in range(nbranch)]
    rstd = [torch.empty(M, device=dev) for _ in range(nbranch)]
    base = torch.randn(M, N, device=dev) if out_mode == 1 else torch.zeros(M, N, device=dev)
    out = base.clone()
    outs = [out] * nbranch if out_mode == 2 else [out] + [torch.zeros(M, N, device=dev) for _ in range(nbranch - 1)]
    arr = (LinearBranch * 2)()
    for b in range(nbranch):
        arr[b] = LinearBranch(p(Xs[b]), p(Ws[b]), p(bs[b]), p(sc[b]), p(of[b]), p(Zs[b]), p(outs[b]), p(mean[b]), p(rstd[b]))
    call = lambda: check(lib.shadow_linear_tc_f32(arr, nbranch, K, K, N, N, M, N, K, act, int(do_norm), out_mode, st()))
    call()
    torch.cuda.synchronize()
    err = 0.0
    want_sum = base.double().clone()
    for b in range(nbranch):
        z = Xs[b].double() @ Ws[b].double().t() + bs[b].double()
        o = act_ref(z, act)
        if do_norm:
            m = o.mean(1, keepdim=True); v = o.var(1, unbiased=False, keepdim=True) + 1e-9
            o = (o - m) * sc[b].double() * torch.rsqrt(v) + of[b].double()
        err = max(err, float((Zs[b].double() - z).abs().max() / z.abs().max()))
        if out_mode == 2:
            want_sum += o
        else:
            want = o + (base.double() if out_mode == 1 and b == 0 else 0)
            err = max(err, float((outs[b].double() - want).abs().max() / want.abs().max()))
    if out_mode == 2:
        err = max(err, float((out.double() - want_sum).abs().max() / want_sum.abs().max()))
    # timing
    for _ in range(3): call()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20): call()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 20
    print(f"M {M} N {N} K {K} act {act} norm {do_norm} branches {nbranch} mode {out_mode}: max rel err {err:.2e}  {us:.1f} us/launch", flush=True)
    return err


bad = 0
if len(sys.argv) > 1:      # ncu target: one configuration
    run(128, 256, 32, 1, False, 1, 0)
    run(4832, 256, 256, 0, True, 2, 2)
    os._exit(0)
for args in [(128, 256, 32, 1, False, 1, 0), (4832, 256, 256, 1, False, 1, 0), (4832, 256, 256, 0, True, 1, 0), (4832, 256, 100, 2, True, 1, 1), (4832, 256, 256, 0, True, 2, 2),
             (4832, 256, 256, 1, False, 2, 0), (1000, 64, 256, 3, True, 1, 0), (77, 48, 36, 4, True, 2, 2), (9664, 256, 256, 0, True, 2, 2)]:
    try:
        bad += run(*args) > 2e-5
    except Exception as e:
        print("FAILED", args, e); bad += 1
# reference timing: warp-level kernel + separate act_norm
M, N, K = 4832, 256, 256
x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev); b = torch.randn(N, device=dev)
for _ in range(3): ops.gemm(x, w, bias=b)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(20): ops.gemm(x, w, bias=b)
g.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
g.replay()
e1.record(); torch.cuda.synchronize()
print(f"gemm_tf32x3 (mma.sync) {M}x{N}x{K}: {e0.elapsed_time(e1) * 1e3 / 20:.1f} us")
print("BAD" if bad else "ALL OK")
