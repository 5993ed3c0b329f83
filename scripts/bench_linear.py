"""Exploration (scratch): time the three products of one Linear (M x 256 x 256 and M x 256 x 100) for the three implementations --
tcgen05 (csrc/gemm_umma.cu), warp-level 3xTF32 (csrc/gemm.cu), fp32 SIMT library GEMM -- eagerly and inside a CUDA graph."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from shadow_gnn_b200 import ops

M = int(sys.argv[1]) if len(sys.argv) > 1 else 4832
dev = torch.device("cuda:0")
res = {}
for K in (256, 100):
    N = 256
    x = torch.randn(M, K, device=dev); w = torch.nn.Parameter(torch.randn(N, K, device=dev)); b = torch.nn.Parameter(torch.randn(N, device=dev))
    dz = torch.randn(M, N, device=dev)
    w.grad = torch.zeros_like(w)
    for mode in ("umma", "tf32x3", "cublas"):
        ops._LINEAR = mode; ops._TC_LINEAR = mode != "cublas"
        fns = {"fwd": lambda: ops._linear_fwd(x, w, b), "dgrad": lambda: ops._linear_dgrad(dz, w), "wgrad": lambda: ops._accum_wgrad(w, dz, x),
               "fwd2": lambda: ops._linear_fwd_pair(x, w, b, x, w, b), "dgrad2": lambda: ops._linear_dgrad_pair(dz, w, dz, w),
               "wgrad2": lambda: ops._accum_wgrad_pair(w, dz, x, w, dz, x)}
        for name, fn in fns.items():
            for _ in range(5):
                fn()
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for _ in range(100):
                fn()
            ev1.record(); torch.cuda.synchronize()
            eager = ev0.elapsed_time(ev1) * 10      # us per call
            g = torch.cuda.CUDAGraph()
            s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                fn()
            torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
            with torch.cuda.graph(g):
                for _ in range(20):
                    fn()
            g.replay(); torch.cuda.synchronize()
            ev0.record()
            for _ in range(10):
                g.replay()
            ev1.record(); torch.cuda.synchronize()
            graphed = ev0.elapsed_time(ev1) * 1000 / 200
            res[f"K{K}_{mode}_{name}"] = (round(eager, 2), round(graphed, 2))
            print(f"M={M} K={K} {mode:7s} {name:6s} eager {eager:8.2f} us   graphed {graphed:8.2f} us", flush=True)
