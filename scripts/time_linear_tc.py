"""Scratch: per-launch time of shadow_linear_tc_f32 (graph replay of 20 launches) under the SHADOW_LTC_DEBUG bits."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from shadow_gnn_b200._lib import lib, LinearBranch, check
dev = torch.device("cuda:0")
p = lambda t: None if t is None else t.data_ptr()
st = lambda: torch.cuda.current_stream().cuda_stream


def time_it(M, N, K, act, norm, nb, mode, dbg):
    os.environ["SHADOW_LTC_DEBUG"] = str(dbg)
    X = [torch.randn(M, K, device=dev) for _ in range(nb)]; W = [torch.randn(N, K, device=dev) for _ in range(nb)]
    b = [torch.randn(N, device=dev) for _ in range(nb)]; Z = [torch.empty(M, N, device=dev) for _ in range(nb)]
    out = torch.zeros(M, N, device=dev); mean = [torch.empty(M, device=dev) for _ in range(nb)]; rstd = [torch.empty(M, device=dev) for _ in range(nb)]
    arr = (LinearBranch * 2)()
    for i in range(nb):
        arr[i] = LinearBranch(p(X[i]), p(W[i]), p(b[i]), p(b[i]), p(b[i]), p(Z[i]), p(out), p(mean[i]), p(rstd[i]))
    call = lambda: check(lib.shadow_linear_tc_f32(arr, nb, K, K, N, N, M, N, K, act, int(norm), mode, st()))
    for _ in range(3): call()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20): call()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    print(f"M {M} K {K} norm {norm} nb {nb} mode {mode} debug {dbg}: {e0.elapsed_time(e1) * 1e3 / 20:.1f} us", flush=True)


for dbg in (0, 1, 2, 3):
    time_it(128, 256, 32, 1, False, 1, 0, dbg)
    time_it(4832, 256, 256, 0, True, 2, 2, dbg)
# empty-kernel reference: a trivial torch op in a graph
x = torch.zeros(1024, device=dev)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(20): x.add_(1)
g.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
print(f"tiny torch kernel: {e0.elapsed_time(e1) * 1e3 / 20:.1f} us")
