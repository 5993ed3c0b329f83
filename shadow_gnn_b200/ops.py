"""torch.autograd bindings of the layer kernels in libshadow_b200.so (csrc/layers.cu) + the device adjacency handle.

Every op here fails loudly without the CUDA library / a CUDA tensor: there is no eager-PyTorch fallback.
"""
import ctypes as C
import os
import weakref

import torch

from ._lib import lib, check, LinearBranch

ACT_ID = {"relu": 0, "I": 1, "elu": 2, "tanh": 3, "leakyrelu": 4}


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream(t):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _req(t, dtype, name):
    if not (t.is_cuda and t.dtype == dtype and t.is_contiguous()):
        raise TypeError(f"{name}: expected a contiguous CUDA {dtype} tensor, got {t.dtype} on {t.device} (no CPU fallback)")
    return t


class DeviceCSR:
    """Adjacency of one batch in the RAW layout of include/shadow_b200.h.

    row_span [n,2] int32 : [start,end) of every row inside `col` / `val`
    col      [*]   int32 : batch-global column ids of the whole sampler super-batch; local id = col - col_off
    val      [*]   fp32  : edge values aligned with `col` (None = all ones), written by normalize_*()
    """

    def __init__(self, row_span, col, col_off, val=None, row_ord=None):
        self.row_span = _req(row_span, torch.int32, "row_span")
        self.col = _req(col, torch.int32, "col")
        self.col_off = int(col_off)
        self.val = val
        self.n = row_span.shape[0]
        self._row_ord = row_ord
        self.normed = None          # None | "rw" | "sym" | "gin" | "gat"

    @property
    def shape(self):
        return (self.n, self.n)

    @property
    def row_ord(self):
        if self._row_ord is None:
            lens = (self.row_span[:, 1] - self.row_span[:, 0])
            ro = torch.zeros(self.n + 1, dtype=torch.int32, device=lens.device)
            torch.cumsum(lens, 0, out=ro[1:])
            self._row_ord = ro
        return self._row_ord

    @property
    def num_edges(self):
        return int(self.row_ord[-1])

    @classmethod
    def from_scipy(cls, adj, device):
        """compat path: a scipy CSR block-diagonal batch as the reference's one_batch produces (minibatch.py:468)"""
        import numpy as np
        indptr = torch.as_tensor(np.asarray(adj.indptr, dtype=np.int64))
        span = torch.stack([indptr[:-1], indptr[1:]], 1).to(torch.int32).contiguous().to(device)
        col = torch.as_tensor(np.asarray(adj.indices, dtype=np.int64)).to(torch.int32).to(device)
        return cls(span, col, 0)

    def _vals(self):
        if self.val is None or self.val.numel() < self.col.numel():
            self.val = torch.empty(self.col.numel(), dtype=torch.float32, device=self.col.device)
        return self.val

    def _fill_and_drop(self, dropedge, seed, step):
        if dropedge > 0 and step is None:
            raise ValueError("dropedge needs a device-side step counter (layers._Dropedge.next())")
        val = self._vals()
        check(lib.shadow_edge_vals_fill(_p(self.row_span), self.n, _p(val), 1.0, _stream(val)))
        if dropedge > 0:
            check(lib.shadow_edge_vals_dropedge(_p(self.row_span), _p(self.row_ord), self.n, float(dropedge), seed & 0xFFFFFFFF,
                                                _p(step), _p(val), _stream(val)))
        return val

    def normalize_rw(self, dropedge=0.0, seed=0, step=None):
        """adj_norm_rw (graph_utils.py:81-95)"""
        val = self._fill_and_drop(dropedge, seed, step)
        check(lib.shadow_edge_vals_row_normalize(_p(self.row_span), self.n, 0, _p(val), _stream(val)))
        self.normed = "rw"
        return self

    def normalize_gin(self, dropedge=0.0, seed=0, step=None):
        """GIN dropedge + rescale (layers.py:512-522)"""
        val = self._fill_and_drop(dropedge, seed, step)
        check(lib.shadow_edge_vals_row_normalize(_p(self.row_span), self.n, 1, _p(val), _stream(val)))
        self.normed = "gin"
        return self

    def mask_only(self, dropedge=0.0, seed=0, step=None):
        """GAT / GATScatter: values stay 1, dropped edges become 0 (layers.py:584-600)"""
        self._fill_and_drop(dropedge, seed, step)
        self.normed = "gat"
        return self

    def normalize_sym(self, dropedge=0.0, seed=0, step=None):
        """adj_norm_sym (graph_utils.py:109-145); the self loops were added by the sampler"""
        val = self._fill_and_drop(dropedge, seed, step)
        deg = torch.empty(self.n, dtype=torch.float32, device=val.device)
        if dropedge > 0:
            mask = val.clone()
            check(lib.shadow_edge_vals_sym_normalize(_p(self.row_span), _p(self.col), self.col_off, self.n, 1, _p(mask), _p(val), _p(deg), _stream(val)))
        else:
            check(lib.shadow_edge_vals_sym_normalize(_p(self.row_span), _p(self.col), self.col_off, self.n, 0, _p(val), _p(val), _p(deg), _stream(val)))
        self.normed = "sym"
        return self

    def to_dense(self):
        """[n,n] dense matrix with duplicates summed (tests)"""
        n = self.n
        out = torch.zeros(n, n, device=self.col.device)
        lens = (self.row_span[:, 1] - self.row_span[:, 0]).long()
        rows = torch.repeat_interleave(torch.arange(n, device=out.device), lens)
        pos = torch.cat([torch.arange(int(s), int(e), device=out.device) for s, e in self.row_span.tolist()]) if n else rows
        vals = self.val[pos] if self.val is not None else torch.ones(pos.numel(), device=out.device)
        out.index_put_((rows, (self.col[pos] - self.col_off).long()), vals, accumulate=True)
        return out


class _SpMM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X, adj):
        X = _req(X.contiguous(), torch.float32, "spmm input")
        Y = torch.empty((adj.n, X.shape[1]), dtype=torch.float32, device=X.device)
        check(lib.shadow_spmm_csr_fwd_f32(_p(adj.row_span), _p(adj.col), adj.col_off, _p(adj.val), _p(X), _p(Y), adj.n, X.shape[1], 0.0, _stream(X)))
        ctx.adj, ctx.nsrc = adj, X.shape[0]
        return Y

    @staticmethod
    def backward(ctx, dY):
        adj = ctx.adj
        dY = dY.contiguous()
        dX = torch.zeros((ctx.nsrc, dY.shape[1]), dtype=torch.float32, device=dY.device)
        check(lib.shadow_spmm_csr_bwd_f32(_p(adj.row_span), _p(adj.col), adj.col_off, _p(adj.val), _p(dY), _p(dX), adj.n, dY.shape[1], _stream(dY)))
        return dX, None


def spmm(adj, X):
    """torch.sparse.mm(adj, X) of the reference (layers.py:326-327)"""
    return _SpMM.apply(X, adj)


class _ActNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, Z, scale, offset, act, do_norm):
        Z = _req(Z.contiguous(), torch.float32, "act_norm input")
        n, D = Z.shape
        out = torch.empty_like(Z)
        mean = torch.empty(n, dtype=torch.float32, device=Z.device)
        rstd = torch.empty(n, dtype=torch.float32, device=Z.device)
        sc = scale.contiguous() if do_norm else None
        of = offset.contiguous() if do_norm else None
        check(lib.shadow_act_norm_fwd_f32(_p(Z), D, _p(sc), _p(of), _p(out), D, _p(mean), _p(rstd), n, D, act, int(do_norm), 0, _stream(Z)))
        ctx.save_for_backward(Z, sc if do_norm else Z.new_empty(0), mean, rstd)
        ctx.act, ctx.do_norm = act, do_norm
        return out

    @staticmethod
    def backward(ctx, dOut):
        Z, sc, mean, rstd = ctx.saved_tensors
        n, D = Z.shape
        dOut = dOut.contiguous()
        dZ = torch.empty_like(Z)
        dscale = torch.zeros(D, dtype=torch.float32, device=Z.device) if ctx.do_norm else None
        doffset = torch.zeros(D, dtype=torch.float32, device=Z.device) if ctx.do_norm else None
        check(lib.shadow_act_norm_bwd_f32(_p(dOut), D, _p(Z), D, _p(sc) if ctx.do_norm else None, _p(mean), _p(rstd), _p(dZ), D, _p(dscale), _p(doffset),
                                          None, n, D, ctx.act, int(ctx.do_norm), _stream(Z)))
        return dZ, dscale, doffset, None, None


def act_norm(Z, scale, offset, act, do_norm=True):
    """norm_feat(act(Z)) (layers.py:329-338); scale/offset are 1-D of length Z.shape[1]"""
    return _ActNorm.apply(Z, scale, offset, ACT_ID[act], bool(do_norm))


def _grad_of(p):
    """the parameter's gradient buffer (created on demand); fused ops accumulate into it in place"""
    if p.grad is None:
        p.grad = torch.zeros_like(p)
    return p.grad


# The dense Linear runs on the tensor cores.  SHADOW_LINEAR selects the implementation (all are parity-tested):
#   "tc"     (default) the hand-written tcgen05 kernel of csrc/linear_tc.cu: TMA-staged operands, tcgen05.mma.kind::tf32 (3xTF32, fp32-accurate)
#            with the accumulator in TMEM, bias + activation + norm_feat fused into the tcgen05.ld epilogue, both GraphSAGE branches per
#            launch; also the input-gradient product (on a transposed copy of the weight).  Weight gradients, and shapes that miss its
#            alignment rules (e.g. the 47-class classifier), take the warp-level kernel below
#   "tf32x3" the warp-level mma.sync kernel of csrc/gemm.cu: error-compensated 3xTF32, fp32-accurate; 13.6-22.9 us per
#            4,832 x 256 x {100,256} product inside the captured step (the fp32 SIMT library GEMM: 13.8-26.5 us)
#   "umma"   the tcgen05 / TMEM / TMA kernel assembled from CUTLASS templates (csrc/gemm_umma.cu, fp32 emulated with 9 bf16 products);
#            shapes that miss the 16-byte TMA alignment (e.g. the 47-class classifier) take the kernel above.  Correct, but its
#            load -> transform -> MMA -> epilogue pipeline is latency-bound at K = 256 with one tile per SM (24.8-32.7 us): not the default
#   "cublas" the fp32 SIMT library GEMM (the former path, kept for A/B timing)
_LINEAR = os.environ.get("SHADOW_LINEAR", "tc")
_TC_LINEAR = _LINEAR != "cublas"
_SPLITK = 16


def _al(*xs):
    return all(int(x) % 4 == 0 for x in xs)


def gemm(A, B, *, trans_a=False, b_kn=False, bias=None, out=None, accumulate=False, split_k=1):
    """C (+)= op(A) op(B) (+ bias) on the tensor cores (include/shadow_b200.h: shadow_gemm_tf32x3_f32).
    trans_a: A is stored [K, M];  b_kn: B is stored [K, N] (default: [N, K], a Linear weight)."""
    A = _req(A.contiguous(), torch.float32, "gemm A")
    B = _req(B.contiguous(), torch.float32, "gemm B")
    K, M = (A.shape if trans_a else (A.shape[1], A.shape[0]))
    N = B.shape[1] if b_kn else B.shape[0]
    assert (B.shape[0] if b_kn else B.shape[1]) == K, "gemm: inner dimensions differ"
    if out is None:
        assert not accumulate
        out = torch.empty((M, N), dtype=torch.float32, device=A.device)
    check(lib.shadow_gemm_tf32x3_f32(_p(A), A.stride(0), int(trans_a), _p(B), B.stride(0), int(b_kn), _p(out), out.stride(0),
                                     _p(bias.contiguous()) if bias is not None else None, M, N, K, int(accumulate), int(split_k), _stream(A)))
    return out


def _linear_fwd(x, w, b):
    """x W^T + b"""
    if not _TC_LINEAR:
        return torch.addmm(b.detach(), x, w.detach().t()) if b is not None else x @ w.detach().t()
    w = w.detach()
    b = b.detach() if b is not None else None
    M, K = x.shape
    N = w.shape[0]
    if _LINEAR == "umma" and M > 0 and _al(K, N) and w.is_contiguous() and x.data_ptr() % 16 == 0 and w.data_ptr() % 16 == 0 and \
            (b is None or (b.is_contiguous() and b.data_ptr() % 16 == 0)):
        Z = torch.empty((M, N), dtype=torch.float32, device=x.device)
        check(lib.shadow_linear_umma_fwd_f32(_p(x), K, _p(w), K, _p(b), _p(Z), N, M, N, K, _stream(x)))
        return Z
    return gemm(x, w, bias=b)


def _linear_dgrad(dZ, w):
    """dZ W"""
    if not _TC_LINEAR:
        return dZ @ w.detach()
    w = w.detach()
    M, K = dZ.shape              # reduction over the Linear's output features
    N = w.shape[1]
    if _LINEAR == "umma" and M > 0 and _al(K, N) and w.is_contiguous() and dZ.is_contiguous() and dZ.data_ptr() % 16 == 0 and w.data_ptr() % 16 == 0:
        dX = torch.empty((M, N), dtype=torch.float32, device=dZ.device)
        check(lib.shadow_linear_umma_dgrad_f32(_p(dZ), K, _p(w), N, _p(dX), N, M, N, K, _stream(dZ)))
        return dX
    return gemm(dZ, w, b_kn=True)


_WG_SCRATCH = {}


def _wgrad_tc(ws, dZs, xs):
    """W.grad += dZ^T x for one or two (W, dZ, x) triples of the same shape on the tcgen05 kernel (csrc/linear_tc.cu: wgrad_tc_kernel +
    wgrad_finish_kernel: slices of 128 rows, summed in slice order -- deterministic).  False when the shapes miss its alignment rules."""
    if _LINEAR != "tc" or os.environ.get("SHADOW_WGRAD_TC", "1") == "0":
        return False
    n, N_out = dZs[0].shape
    K_in = xs[0].shape[1]
    gs = [_grad_of(w) for w in ws]
    ok = n > 0 and 8 <= N_out <= 256 and 8 <= K_in <= 256 and N_out % 4 == 0 and K_in % 4 == 0 and \
        all(t.is_contiguous() and t.dtype == torch.float32 and t.data_ptr() % 16 == 0 for t in (*dZs, *xs, *gs)) and \
        all(dz.shape == dZs[0].shape for dz in dZs) and all(x.shape == xs[0].shape for x in xs)
    if not ok:
        return False
    need = int(lib.shadow_wgrad_tc_scratch_floats(n, N_out, K_in))
    key = (dZs[0].device.index, N_out, K_in)
    sc = _WG_SCRATCH.get(key)
    if sc is None or sc[0].numel() < need:
        sc = _WG_SCRATCH[key] = [torch.empty(need, dtype=torch.float32, device=dZs[0].device) for _ in range(2)]
    two = len(ws) == 2
    check(lib.shadow_wgrad_tc_f32(_p(dZs[0]), _p(xs[0]), _p(gs[0]), _p(sc[0]), _p(dZs[1]) if two else None, _p(xs[1]) if two else None,
                                  _p(gs[1]) if two else None, _p(sc[1]) if two else None, n, N_out, K_in, _stream(dZs[0])))
    return True


def _accum_wgrad(w, dZ, x):
    """W.grad += dZ^T x.  The product is [D_out, n] x [n, D_in] with n ~ 4,800 and a 256 x 256 result: a handful of output tiles, so it is
    split along n (tcgen05 path: 16 batched slices + a sum; warp-MMA path: ~160 rows per CTA, fp32 atomics into the gradient)."""
    if _wgrad_tc([w], [dZ], [x]):
        return
    g = _grad_of(w)
    n = x.shape[0]
    if not _TC_LINEAR:
        if n % _SPLITK == 0 and n >= 64 * _SPLITK:
            part = torch.bmm(dZ.view(_SPLITK, n // _SPLITK, dZ.shape[1]).transpose(1, 2), x.view(_SPLITK, n // _SPLITK, x.shape[1]))
            g.add_(part.sum(0))
        else:
            g.addmm_(dZ.t(), x)
        return
    N_out, K_in = dZ.shape[1], x.shape[1]
    if _LINEAR == "umma" and n % _SPLITK == 0 and n >= 64 * _SPLITK and _al(N_out, K_in) and dZ.is_contiguous() and x.is_contiguous() and \
            dZ.data_ptr() % 16 == 0 and x.data_ptr() % 16 == 0:
        part = torch.empty((_SPLITK, N_out, K_in), dtype=torch.float32, device=x.device)
        check(lib.shadow_linear_umma_wgrad_f32(_p(dZ), N_out, _p(x), K_in, _p(part), n // _SPLITK, _SPLITK, N_out, K_in, _stream(x)))
        g.add_(part.sum(0))
    else:
        gemm(dZ, x, trans_a=True, b_kn=True, out=g, accumulate=True, split_k=max(1, n // 160))


def _pair_ok(*ts):
    return _LINEAR in ("tf32x3", "tc") and all(t.is_contiguous() and t.dtype == torch.float32 for t in ts)


def _tc_ok(N, K, *ts):
    """alignment rules of csrc/linear_tc.cu (TMA descriptors, 128-bit epilogue stores)"""
    return _LINEAR == "tc" and 8 <= N <= 256 and N % 4 == 0 and K % 4 == 0 and K >= 4 and \
        all(t is None or (t.is_contiguous() and t.dtype == torch.float32 and t.data_ptr() % 16 == 0) for t in ts)


class WeightPlanes:
    """The two TF32 planes (head = what a TF32 tensor core reads, exact fp32 remainder) of every parameter of one flat buffer, and of the
    TRANSPOSE of every 2-D parameter, kept in step with the weights by their owner (FlatAdamClip refreshes them right after each optimizer
    step: 2 launches per step instead of a split of every weight tile in every CTA of every Linear launch).
    Weights changed behind the owner's back (checkpoint restore, manual edits) must be followed by `mark_stale()`; DeepGNN does that from
    a load_state_dict hook."""
    _owners = []

    def __init__(self, flat, params, offsets):
        self.flat, self.n = flat, flat.numel()
        self.hi, self.lo, self.t_hi, self.t_lo = (torch.zeros_like(flat) for _ in range(4))
        rows = [(offsets[id(p)], p.shape[0], p.shape[1], offsets[id(p)]) for p in params if p.dim() == 2]
        self.table = torch.tensor(rows if rows else [(0, 0, 0, 0)], dtype=torch.int64, device=flat.device)
        self.ntab = len(rows)
        self.fresh = False
        WeightPlanes._owners = [r for r in WeightPlanes._owners if r() is not None and r().flat.data_ptr() != flat.data_ptr()] + [weakref.ref(self)]

    def mark_stale(self):
        self.fresh = False

    def refresh(self):
        check(lib.shadow_tf32_split_f32(_p(self.flat), self.n, _p(self.hi), _p(self.lo), _stream(self.flat)))
        check(lib.shadow_tf32_split_transpose_f32(_p(self.flat), _p(self.table), self.ntab, _p(self.t_hi), _p(self.t_lo), _stream(self.flat)))
        self.fresh = True

    @staticmethod
    def of(w, transposed=False):
        """(hi, lo) planes of the weight `w` (a view into a registered flat buffer) or of its transpose; None when `w` is not registered"""
        if os.environ.get("SHADOW_LTC_PRESPLIT", "1") == "0":
            return None
        for r in WeightPlanes._owners:
            o = r()
            if o is None:
                continue
            off = (w.data_ptr() - o.flat.data_ptr()) // 4
            if 0 <= off < o.n and w.is_contiguous() and w.device == o.flat.device:
                if not o.fresh:
                    o.refresh()
                a, b = (o.t_hi, o.t_lo) if transposed else (o.hi, o.lo)
                shape = (w.shape[1], w.shape[0]) if transposed else tuple(w.shape)
                return a[off:off + w.numel()].view(shape), b[off:off + w.numel()].view(shape)
        return None


def _linear_tc(branches, M, N, K, act, do_norm, out_mode, w_lo=None):
    """branches: 1 or 2 tuples (X, W, bias, scale, offset, Z, out, mean, rstd) of the same shape -> one launch of shadow_linear_tc_f32;
    w_lo: the weights' remainder planes when W holds their TF32 heads (WeightPlanes)"""
    arr = (LinearBranch * 2)()
    for i, b in enumerate(branches):
        arr[i] = LinearBranch(*[None if t is None else t.data_ptr() for t in b], None if w_lo is None else w_lo[i].data_ptr())
    x, w, out, Z = branches[0][0], branches[0][1], branches[0][6], branches[0][5]
    check(lib.shadow_linear_tc_f32(arr, len(branches), x.stride(0), w.stride(0), Z.stride(0) if Z is not None else N, out.stride(0), M, N, K, act,
                                   int(do_norm), out_mode, _stream(x)))


def _dgrad_tc_pair(dZs, ws):
    """[dZ W for each pair] on the tcgen05 kernel: the product reduces over the Linear's OUTPUT features, so the K-major B operand is W^T"""
    M, N_out = dZs[0].shape
    K_in = ws[0].shape[1]
    planes = [WeightPlanes.of(w.detach(), transposed=True) for w in ws]
    if all(pl is not None for pl in planes):
        wts, w_lo = [pl[0] for pl in planes], [pl[1] for pl in planes]
    else:
        wts, w_lo = [w.detach().t().contiguous() for w in ws], None
    outs = [torch.empty((M, K_in), dtype=torch.float32, device=dZs[0].device) for _ in dZs]
    if not _tc_ok(K_in, N_out, *dZs, *wts, *outs) or (w_lo is not None and not _tc_ok(K_in, N_out, *w_lo)):
        return None
    _linear_tc([(dz, wt, None, None, None, None, o, None, None) for dz, wt, o in zip(dZs, wts, outs)], M, K_in, N_out, ACT_ID["I"], False, 0, w_lo)
    return outs


def _linear_fwd_pair(x0, w0, b0, x1, w1, b1):
    """(x0 W0^T + b0, x1 W1^T + b1), one launch when both products have the same shape"""
    if _pair_ok(x0, x1, w0, w1) and x0.shape == x1.shape and w0.shape == w1.shape:
        w0, w1, b0, b1 = w0.detach(), w1.detach(), b0.detach(), b1.detach()
        M, K = x0.shape
        N = w0.shape[0]
        Z0 = torch.empty((M, N), dtype=torch.float32, device=x0.device)
        Z1 = torch.empty_like(Z0)
        check(lib.shadow_gemm_tf32x3_pair_f32(_p(x0), _p(x1), K, 0, _p(w0), _p(w1), K, 0, _p(Z0), _p(Z1), N, _p(b0), _p(b1), M, N, K, 0, 1, _stream(x0)))
        return Z0, Z1
    return _linear_fwd(x0, w0, b0), _linear_fwd(x1, w1, b1)


def _linear_dgrad_pair(dZ0, w0, dZ1, w1):
    if _pair_ok(dZ0, dZ1, w0, w1) and dZ0.shape == dZ1.shape and w0.shape == w1.shape:
        w0, w1 = w0.detach(), w1.detach()
        M, K = dZ0.shape
        N = w0.shape[1]
        d0 = torch.empty((M, N), dtype=torch.float32, device=dZ0.device)
        d1 = torch.empty_like(d0)
        check(lib.shadow_gemm_tf32x3_pair_f32(_p(dZ0), _p(dZ1), K, 0, _p(w0), _p(w1), N, 1, _p(d0), _p(d1), N, None, None, M, N, K, 0, 1, _stream(dZ0)))
        return d0, d1
    return _linear_dgrad(dZ0, w0), _linear_dgrad(dZ1, w1)


# ---- weight gradients off the critical path ----
# W.grad of a layer is needed by nobody until the backward pass is over (optimizer / gradient exchange), while its input gradient is what
# the next node of the backward pass waits for.  The weight-gradient launches therefore go to a side stream that forks from the current
# stream where dZ exists and joins it again (a) when the autograd engine finishes the pass (queue_callback) and (b) wherever gradients are
# read earlier (the in-graph gradient exchange).  One wave of 76 CTAs each: wgrad and dgrad / SpMM^T / act-norm backward share the 148 SMs.
# Inside a CUDA-graph capture the fork / join become graph edges.  SHADOW_WGRAD_OVERLAP=0 keeps everything on one stream.
_WG_OVERLAP = os.environ.get("SHADOW_WGRAD_OVERLAP", "1") != "0"
_WG_SIDE = {}                  # device index -> [side stream, tensors kept alive until the join, pending]
_ANB_POOL, _ANB_POOL_NEXT = {}, {}


def join_wgrad():
    """make the current stream wait for every weight-gradient launch issued so far (no-op when none is pending)"""
    for st in _WG_SIDE.values():
        if st[2]:
            torch.cuda.current_stream(st[0].device).wait_stream(st[0])
            st[1].clear()
            st[2] = False
    _ANB_POOL_NEXT.clear()


def wgrad_stream(device):
    """the side stream that currently holds un-joined weight-gradient launches of `device`, or None"""
    st = _WG_SIDE.get(torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device())
    return st[0] if st is not None and st[2] else None


def _wgrad_async(fn, keep):
    if not _WG_OVERLAP or not keep[0].is_cuda:
        fn()
        return False
    dev = keep[0].device
    st = _WG_SIDE.get(dev.index)
    if st is None:
        st = _WG_SIDE[dev.index] = [torch.cuda.Stream(device=dev), [], False]
    if not st[2]:
        st[2] = True
        try:
            torch.autograd.Variable._execution_engine.queue_callback(join_wgrad)        # end of this backward pass
        except RuntimeError:                               # not inside a backward pass (direct call): run in line
            st[2] = False
            fn()
            return False
    st[0].wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(st[0]):
        fn()
    st[1].extend(keep)
    return True


def _accum_wgrad_pair(w0, dZ0, x0, w1, dZ1, x1):
    if w0.shape == w1.shape and _wgrad_tc([w0, w1], [dZ0, dZ1], [x0, x1]):
        return
    if _pair_ok(dZ0, dZ1, x0, x1) and dZ0.shape == dZ1.shape and x0.shape == x1.shape:
        g0, g1 = _grad_of(w0), _grad_of(w1)
        n, N_out, K_in = x0.shape[0], dZ0.shape[1], x0.shape[1]
        if g0.is_contiguous() and g1.is_contiguous():
            check(lib.shadow_gemm_tf32x3_pair_f32(_p(dZ0), _p(dZ1), N_out, 1, _p(x0), _p(x1), K_in, 1, _p(g0), _p(g1), K_in, None, None, N_out, K_in, n, 1,
                                                  max(1, n // 160), _stream(x0)))
            return
    _accum_wgrad(w0, dZ0, x0)
    _accum_wgrad(w1, dZ1, x1)


def _act_norm_fwd_raw(Z, scale, offset, idx, act, do_norm, out=None, accumulate=False):
    n, D = Z.shape
    out = torch.empty_like(Z) if out is None else out
    mean = torch.empty(n, dtype=torch.float32, device=Z.device)
    rstd = torch.empty(n, dtype=torch.float32, device=Z.device)
    sc = scale.detach()[idx] if do_norm else None
    of = offset.detach()[idx] if do_norm else None
    check(lib.shadow_act_norm_fwd_f32(_p(Z), D, _p(sc), _p(of), _p(out), D, _p(mean), _p(rstd), n, D, act, int(do_norm), int(accumulate), _stream(Z)))
    return out, mean, rstd


def _act_norm_bwd_raw(dOut, Z, scale, offset, bias, idx, mean, rstd, act, do_norm):
    """dZ; dscale[idx] / doffset[idx] / dbias are accumulated straight into the parameters' .grad"""
    n, D = Z.shape
    dZ = torch.empty_like(Z)
    sc = scale.detach()[idx] if do_norm else None
    ds = _grad_of(scale)[idx] if do_norm else None
    do = _grad_of(offset)[idx] if do_norm else None
    db = _grad_of(bias) if bias is not None else None
    check(lib.shadow_act_norm_bwd_f32(_p(dOut), D, _p(Z), D, _p(sc), _p(mean), _p(rstd), _p(dZ), D, _p(ds), _p(do), _p(db), n, D, act, int(do_norm), _stream(Z)))
    return dZ


_ANB_SCRATCH = {}


def _act_norm_bwd_pair(dOut, Zs, scale, offset, biases, idxs, means, rstds, act, do_norm):
    """dZ of one or two branches that share dOut (csrc/layers.cu: act_norm_bwd_pair_kernel + colsum_finish_kernel, deterministic);
    dscale[idx] / doffset[idx] / dbias are accumulated straight into the parameters' .grad"""
    n, D = Zs[0].shape
    dev = dOut.device
    key = (dev.index, D)
    if key not in _ANB_SCRATCH:
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        _ANB_SCRATCH[key] = torch.empty(8 * sms * 2 * 3 * D, dtype=torch.float32, device=dev)      # up to 8 CTAs per SM (SHADOW_ANB_PAIR_MULT)
    scratch = _ANB_SCRATCH[key]
    dZs = [torch.empty_like(Z) for Z in Zs]
    nb = len(Zs)
    if nb == 2 and do_norm and _WG_OVERLAP and dOut.is_cuda and all(b is not None for b in biases):
        # column sums (dscale / doffset / dbias) finish on the weight-gradient side stream: every call of a backward pass takes its own
        # partial-sum buffer from a rotating pool (the next layer's kernel must not overwrite partials the side stream still reads)
        pool = _ANB_POOL.setdefault(key, [])
        idx = _ANB_POOL_NEXT.get(key, 0)                   # calls since the last join
        while len(pool) <= idx:
            pool.append(torch.empty_like(scratch))
        part = pool[idx]
        check(lib.shadow_act_norm_bwd_pair_nofinish_f32(_p(dOut), D, _p(Zs[0]), _p(Zs[1]), D, _p(scale.detach()[idxs[0]]), _p(scale.detach()[idxs[1]]),
                                                        _p(means[0]), _p(rstds[0]), _p(means[1]), _p(rstds[1]), _p(dZs[0]), _p(dZs[1]), D, n, D, act, 1,
                                                        _p(part), part.numel(), _stream(dOut)))
        nparts = int(lib.shadow_act_norm_bwd_pair_nparts(n))
        gs, go = _grad_of(scale), _grad_of(offset)
        dsts = (gs[idxs[0]], go[idxs[0]], _grad_of(biases[0]), gs[idxs[1]], go[idxs[1]], _grad_of(biases[1]))
        if _wgrad_async(lambda: check(lib.shadow_colsum_finish_f32(_p(part), nparts, D, *[_p(t) for t in dsts], _stream(part))), (part,)):
            _ANB_POOL_NEXT[key] = idx + 1
        return dZs
    sc = [scale.detach()[i] if do_norm else None for i in idxs]
    ds = [_grad_of(scale)[i] if do_norm else None for i in idxs]
    do = [_grad_of(offset)[i] if do_norm else None for i in idxs]
    db = [_grad_of(b) if b is not None else None for b in biases]
    g = lambda lst, i: _p(lst[i]) if i < nb else None
    check(lib.shadow_act_norm_bwd_pair_f32(_p(dOut), D, g(Zs, 0), g(Zs, 1), D, g(sc, 0), g(sc, 1), g(means, 0), g(rstds, 0), g(means, 1), g(rstds, 1),
                                           g(dZs, 0), g(dZs, 1), D, g(ds, 0), g(do, 0), g(db, 0), g(ds, 1), g(do, 1), g(db, 1), n, D, act, int(do_norm),
                                           _p(scratch), scratch.numel(), _stream(dOut)))
    return dZs


class _LinearActNorm(torch.autograd.Function):
    """out = norm_feat_idx(act(x W^T + b)) with ONE fused backward kernel; parameter gradients (W, b, scale[idx], offset[idx]) are
    accumulated in place into `.grad` (the flat bucket of FlatAdamClip), so autograd launches no bias-sum / select / accumulate kernels"""

    @staticmethod
    def forward(ctx, x, lin_w, lin_b, scale, offset, idx, act, do_norm):
        x = _req(x.contiguous(), torch.float32, "linear input")
        M, K = x.shape
        N = lin_w.shape[0]
        w = lin_w.detach()
        if M > 0 and _tc_ok(N, K, x, w):
            Z = torch.empty((M, N), dtype=torch.float32, device=x.device)
            out = torch.empty_like(Z)
            mean = torch.empty(M, dtype=torch.float32, device=x.device)
            rstd = torch.empty(M, dtype=torch.float32, device=x.device)
            pl = WeightPlanes.of(w)
            _linear_tc([(x, pl[0] if pl else w, lin_b.detach() if lin_b is not None else None, scale.detach()[idx] if do_norm else None,
                         offset.detach()[idx] if do_norm else None, Z, out, mean, rstd)], M, N, K, act, do_norm, 0, [pl[1]] if pl else None)
        else:
            Z = _linear_fwd(x, lin_w, lin_b)
            out, mean, rstd = _act_norm_fwd_raw(Z, scale, offset, idx, act, do_norm)
        ctx.save_for_backward(x, Z, mean, rstd)
        ctx.p = (lin_w, lin_b, scale, offset, idx, act, do_norm)
        return out

    @staticmethod
    def backward(ctx, dOut):
        x, Z, mean, rstd = ctx.saved_tensors
        lin_w, lin_b, scale, offset, idx, act, do_norm = ctx.p
        if Z.shape[1] <= 256 and Z.shape[1] % 4 == 0 and Z.shape[0] > 0:
            dZ = _act_norm_bwd_pair(dOut.contiguous(), [Z], scale, offset, [lin_b], [idx], [mean], [rstd], act, do_norm)[0]
        else:
            dZ = _act_norm_bwd_raw(dOut.contiguous(), Z, scale, offset, lin_b, idx, mean, rstd, act, do_norm)
        _accum_wgrad(lin_w, dZ, x)
        dX = None
        if ctx.needs_input_grad[0]:
            r = _dgrad_tc_pair([dZ], [lin_w]) if _LINEAR == "tc" and dZ.shape[0] > 0 else None
            dX = r[0] if r is not None else _linear_dgrad(dZ, lin_w)
        return dX, None, None, None, None, None, None, None


def linear_act_norm(x, lin, scale, offset, idx, act, do_norm=True):
    return _LinearActNorm.apply(x, lin.weight, lin.bias, scale, offset, idx, ACT_ID[act], bool(do_norm))


class _SageLayer(torch.autograd.Function):
    """GraphSAGE layer body (layers.py:474-483 of the reference) as one autograd node:
    fwd  spmm, 2 GEMMs, 2 fused act+norm launches (the second accumulates into the first's output)
    bwd  2 fused act+norm backward launches, 4 GEMMs (weight grads in place), spmm^T accumulating into dX"""

    @staticmethod
    def forward(ctx, x, adj, ws, bs, wn, bn, scale, offset, act, do_norm):
        x = _req(x.contiguous(), torch.float32, "sage input")
        agg = torch.empty_like(x)
        check(lib.shadow_spmm_csr_fwd_f32(_p(adj.row_span), _p(adj.col), adj.col_off, _p(adj.val), _p(x), _p(agg), adj.n, x.shape[1], 0.0, _stream(x)))
        M, K = x.shape
        N = ws.shape[0]
        w0, w1 = ws.detach(), wn.detach()
        if M > 0 and ws.shape == wn.shape and _tc_ok(N, K, x, agg, w0, w1):     # both branches in one tcgen05 launch, epilogue fused
            Zs = torch.empty((M, N), dtype=torch.float32, device=x.device)
            Zn = torch.empty_like(Zs)
            out = torch.zeros_like(Zs)
            mean_s, rstd_s, mean_n, rstd_n = (torch.empty(M, dtype=torch.float32, device=x.device) for _ in range(4))
            sc, of = scale.detach(), offset.detach()
            p0, p1 = WeightPlanes.of(w0), WeightPlanes.of(w1)
            pre = p0 is not None and p1 is not None
            _linear_tc([(x, p0[0] if pre else w0, bs.detach(), sc[0] if do_norm else None, of[0] if do_norm else None, Zs, out, mean_s, rstd_s),
                        (agg, p1[0] if pre else w1, bn.detach(), sc[1] if do_norm else None, of[1] if do_norm else None, Zn, out, mean_n, rstd_n)],
                       M, N, K, act, do_norm, 2, [p0[1], p1[1]] if pre else None)
        else:
            Zs, Zn = _linear_fwd_pair(x, ws, bs, agg, wn, bn)
            out, mean_s, rstd_s = _act_norm_fwd_raw(Zs, scale, offset, 0, act, do_norm)
            _, mean_n, rstd_n = _act_norm_fwd_raw(Zn, scale, offset, 1, act, do_norm, out=out, accumulate=True)
        ctx.save_for_backward(x, agg, Zs, Zn, mean_s, rstd_s, mean_n, rstd_n)
        ctx.p = (adj, ws, bs, wn, bn, scale, offset, act, do_norm)
        return out

    @staticmethod
    def backward(ctx, dOut):
        x, agg, Zs, Zn, mean_s, rstd_s, mean_n, rstd_n = ctx.saved_tensors
        adj, ws, bs, wn, bn, scale, offset, act, do_norm = ctx.p
        dOut = dOut.contiguous()
        if Zs.shape[1] <= 256 and Zs.shape[1] % 4 == 0 and Zs.shape[0] > 0:       # both branches in one launch: dOut is read once, column sums without atomics
            dZs, dZn = _act_norm_bwd_pair(dOut, [Zs, Zn], scale, offset, [bs, bn], [0, 1], [mean_s, mean_n], [rstd_s, rstd_n], act, do_norm)
        else:
            dZs = _act_norm_bwd_raw(dOut, Zs, scale, offset, bs, 0, mean_s, rstd_s, act, do_norm)
            dZn = _act_norm_bwd_raw(dOut, Zn, scale, offset, bn, 1, mean_n, rstd_n, act, do_norm)
        if ctx.needs_input_grad[0]:
            _wgrad_async(lambda: _accum_wgrad_pair(ws, dZs, x, wn, dZn, agg), (dZs, dZn, x, agg))
        else:                                              # first layer: nothing left to overlap with
            _accum_wgrad_pair(ws, dZs, x, wn, dZn, agg)
            return (None,) * 10
        r = _dgrad_tc_pair([dZs, dZn], [ws, wn]) if _LINEAR == "tc" and ws.shape == wn.shape and dZs.shape[0] > 0 else None
        dX, dAgg = r if r is not None else _linear_dgrad_pair(dZs, ws, dZn, wn)
        check(lib.shadow_spmm_csr_bwd_f32(_p(adj.row_span), _p(adj.col), adj.col_off, _p(adj.val), _p(dAgg), _p(dX), adj.n, dX.shape[1], _stream(dX)))
        return (dX,) + (None,) * 9


def sage_layer(x, adj, lin_self, lin_neigh, scale, offset, act, do_norm=True):
    return _SageLayer.apply(x, adj, lin_self.weight, lin_self.bias, lin_neigh.weight, lin_neigh.bias, scale, offset, ACT_ID[act], bool(do_norm))


class _GATAgg(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a_self, a_neigh, H, adj, heads):
        n, D = H.shape
        d = D // heads
        a_self, a_neigh, H = a_self.contiguous(), a_neigh.contiguous(), _req(H.contiguous(), torch.float32, "gat input")
        out = torch.empty_like(H)
        rowmax = torch.empty(n * heads, dtype=torch.float32, device=H.device)
        denom = torch.empty(n * heads, dtype=torch.float32, device=H.device)
        check(lib.shadow_gat_fwd_f32(_p(adj.row_span), _p(adj.col), adj.col_off, _p(adj.val), _p(a_self), _p(a_neigh), _p(H), _p(out), _p(rowmax), _p(denom),
                                     n, heads, d, _stream(H)))
        ctx.save_for_backward(a_self, a_neigh, H, out, rowmax, denom)
        ctx.adj, ctx.heads = adj, heads
        return out

    @staticmethod
    def backward(ctx, dOut):
        a_self, a_neigh, H, out, rowmax, denom = ctx.saved_tensors
        adj, heads = ctx.adj, ctx.heads
        n, D = H.shape
        dOut = dOut.contiguous()
        dH = torch.zeros_like(H)
        da_self = torch.zeros_like(a_self)
        da_neigh = torch.zeros_like(a_neigh)
        check(lib.shadow_gat_bwd_f32(_p(adj.row_span), _p(adj.col), adj.col_off, _p(adj.val), _p(a_self), _p(a_neigh), _p(H), _p(out), _p(rowmax), _p(denom),
                                     _p(dOut), _p(dH), _p(da_self), _p(da_neigh), n, heads, D // heads, _stream(H)))
        return da_self, da_neigh, dH, None, None


_GAT_SCRATCH = {}


def gat_layer_supported(x, w, heads):
    """shapes the fused GAT node covers (csrc/gat.cu + csrc/linear_tc.cu); everything else takes the composed path of layers.GAT"""
    D, K = w.shape
    return _LINEAR == "tc" and D % heads == 0 and bool(lib.shadow_gat_supported(heads, D // heads)) and x.dim() == 2 and x.shape[0] > 0 and \
        _tc_ok(D, K, x, w.detach()) and _tc_ok(K, D) and os.environ.get("SHADOW_GAT_FUSED", "1") != "0"


class _GATLayer(torch.autograd.Function):
    """GAT layer body (layers.py:607-629 of the reference) as ONE autograd node:
    fwd  both Linear + activation in one tcgen05 launch, attention logits, softmax aggregation of all heads, per-head norm_feat of both branches
    bwd  the three mirrored kernels, the activation / logit backward with its column sums, both weight gradients in one launch, and the
         input gradient dZ_self W_self + dZ_neigh W_neigh in one launch (both products reduce-add into the same tensor)"""

    @staticmethod
    def forward(ctx, x, adj, w0, b0, w1, b1, att, scale, offset, act, do_norm, heads):
        x = _req(x.contiguous(), torch.float32, "gat input")
        M, K = x.shape
        D = w0.shape[0]
        d = D // heads
        dev = x.device
        h_self = torch.empty((M, D), dtype=torch.float32, device=dev)
        h_neigh = torch.empty_like(h_self)
        p0, p1 = WeightPlanes.of(w0.detach()), WeightPlanes.of(w1.detach())
        pre = p0 is not None and p1 is not None
        _linear_tc([(x, p0[0] if pre else w0.detach(), b0.detach(), None, None, None, h_self, None, None),
                    (x, p1[0] if pre else w1.detach(), b1.detach(), None, None, None, h_neigh, None, None)], M, D, K, act, False, 0, [p0[1], p1[1]] if pre else None)
        s_self, s_neigh, a_self, a_neigh, rowmax, denom = (torch.empty((M, heads), dtype=torch.float32, device=dev) for _ in range(6))
        attc = att.detach().contiguous()
        check(lib.shadow_gat_logits_fwd_f32(_p(h_self), _p(h_neigh), _p(attc), _p(s_self), _p(s_neigh), _p(a_self), _p(a_neigh), M, heads, d, _stream(x)))
        agg = torch.empty_like(h_self)
        check(lib.shadow_gat_agg_fwd_f32(_p(adj.row_span), _p(adj.col), adj.col_off, _p(adj.val), _p(a_self), _p(a_neigh), _p(h_neigh), _p(agg), _p(rowmax), _p(denom),
                                         M, heads, d, _stream(x)))
        out = torch.empty_like(h_self)
        mean = torch.empty((2, M, heads), dtype=torch.float32, device=dev)
        rstd = torch.empty_like(mean)
        sc, of = scale.detach().contiguous(), offset.detach().contiguous()
        check(lib.shadow_gat_headnorm_fwd_f32(_p(h_self), _p(agg), _p(sc), _p(of), _p(out), _p(mean), _p(rstd), M, heads, d, int(do_norm), _stream(x)))
        ctx.save_for_backward(x, h_self, h_neigh, s_self, s_neigh, a_self, a_neigh, agg, rowmax, denom, mean, rstd)
        ctx.p = (adj, w0, b0, w1, b1, att, scale, offset, act, do_norm, heads)
        return out

    @staticmethod
    def backward(ctx, dOut):
        x, h_self, h_neigh, s_self, s_neigh, a_self, a_neigh, agg, rowmax, denom, mean, rstd = ctx.saved_tensors
        adj, w0, b0, w1, b1, att, scale, offset, act, do_norm, heads = ctx.p
        M, K = x.shape
        D = w0.shape[0]
        d = D // heads
        dev = x.device
        dOut = dOut.contiguous()
        key = (dev.index, heads, d)
        if key not in _GAT_SCRATCH:
            _GAT_SCRATCH[key] = torch.empty(int(lib.shadow_gat_scratch_floats(heads, d)), dtype=torch.float32, device=dev)
        scratch = _GAT_SCRATCH[key]
        st = _stream(x)
        dh_self, dagg = torch.empty_like(h_self), torch.empty_like(h_self)
        assert scale.is_contiguous() and offset.is_contiguous() and att.is_contiguous()
        check(lib.shadow_gat_headnorm_bwd_f32(_p(dOut), _p(h_self), _p(agg), _p(scale.detach()), _p(mean), _p(rstd), _p(dh_self), _p(dagg),
                                              _p(_grad_of(scale)) if do_norm else None, _p(_grad_of(offset)) if do_norm else None, M, heads, d, int(do_norm),
                                              _p(scratch), scratch.numel(), st))
        dh_neigh = torch.zeros_like(h_self)
        da_self = torch.zeros((M, heads), dtype=torch.float32, device=dev)
        da_neigh = torch.zeros((M, heads), dtype=torch.float32, device=dev)
        check(lib.shadow_gat_agg_bwd_f32(_p(adj.row_span), _p(adj.col), adj.col_off, _p(adj.val), _p(a_self), _p(a_neigh), _p(h_neigh), _p(agg), _p(rowmax), _p(denom),
                                         _p(dagg), _p(dh_neigh), _p(da_self), _p(da_neigh), M, heads, d, st))
        dZs, dZn = torch.empty_like(h_self), torch.empty_like(h_self)
        check(lib.shadow_gat_pre_bwd_f32(_p(dh_self), _p(dh_neigh), _p(h_self), _p(h_neigh), _p(s_self), _p(s_neigh), _p(da_self), _p(da_neigh), _p(att.detach()),
                                         _p(dZs), _p(dZn), _p(_grad_of(att)), _p(_grad_of(b0)), _p(_grad_of(b1)), M, heads, d, act, _p(scratch), scratch.numel(), st))
        if ctx.needs_input_grad[0]:
            _wgrad_async(lambda: _accum_wgrad_pair(w0, dZs, x, w1, dZn, x), (dZs, dZn, x))
        else:
            _accum_wgrad_pair(w0, dZs, x, w1, dZn, x)
        dX = None
        if ctx.needs_input_grad[0]:
            dX = torch.zeros((M, K), dtype=torch.float32, device=dev)
            planes = [WeightPlanes.of(w.detach(), transposed=True) for w in (w0, w1)]
            if all(pl is not None for pl in planes):
                wts, w_lo = [pl[0] for pl in planes], [pl[1] for pl in planes]
            else:
                wts, w_lo = [w.detach().t().contiguous() for w in (w0, w1)], None
            _linear_tc([(dZs, wts[0], None, None, None, None, dX, None, None), (dZn, wts[1], None, None, None, None, dX, None, None)], M, K, D, ACT_ID["I"], False, 2, w_lo)
        return (dX,) + (None,) * 11


def gat_layer(x, adj, lin_self, lin_neigh, attention, scale, offset, act, heads, do_norm=True):
    return _GATLayer.apply(x, adj, lin_self.weight, lin_self.bias, lin_neigh.weight, lin_neigh.bias, attention, scale, offset, ACT_ID[act], bool(do_norm), heads)


def gat_aggregate(adj, a_self, a_neigh, H, heads):
    """GAT._aggregate_attention for all heads at once (layers.py:560-582)"""
    return _GATAgg.apply(a_self, a_neigh, H, adj, heads)


class _SegPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X, seg, seg_off, mode):
        X = _req(X.contiguous(), torch.float32, "pool input")
        S, F = seg.numel() - 1, X.shape[1]
        out = torch.empty((S, F), dtype=torch.float32, device=X.device)
        argmax = torch.empty((S, F), dtype=torch.int32, device=X.device) if mode == 2 else None
        check(lib.shadow_segment_pool_fwd_f32(_p(X), _p(seg), seg_off, S, F, mode, _p(out), _p(argmax), _stream(X)))
        ctx.seg, ctx.seg_off, ctx.mode, ctx.argmax, ctx.n = seg, seg_off, mode, argmax, X.shape[0]
        return out

    @staticmethod
    def backward(ctx, dOut):
        dOut = dOut.contiguous()
        S, F = dOut.shape
        dX = torch.zeros((ctx.n, F), dtype=torch.float32, device=dOut.device)
        check(lib.shadow_segment_pool_bwd_f32(_p(dOut), _p(ctx.seg), ctx.seg_off, S, F, ctx.mode, _p(ctx.argmax), _p(dX), _stream(dOut)))
        return dX, None, None, None


def segment_pool(X, sizes_subg, mode):
    """F.embedding_bag(arange, X, offsets, mode) of ResPool (layers.py:168-184): sizes_subg [B] rows per subgraph"""
    seg = torch.zeros(sizes_subg.numel() + 1, dtype=torch.int32, device=X.device)
    torch.cumsum(sizes_subg.to(torch.int32), 0, out=seg[1:])
    return _SegPool.apply(X, seg, 0, {"sum": 0, "mean": 1, "max": 2}[mode])


class _RawDev:
    """zero-copy torch view of device memory owned by the library (`__cuda_array_interface__`)"""

    def __init__(self, ptr, count, typestr):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": typestr, "data": (int(ptr), False), "version": 2, "strides": None}


class _P2PGrad:
    """The flat gradient buffer of one rank in IPC-shareable device memory + every peer's buffer opened in this process (NVLink peer access),
    for the fused exchange of csrc/layers.cu (p2p_reduce_sqnorm_kernel / p2p_wait_zero_kernel).  Handles travel through torch.distributed."""

    def __init__(self, n, dev):
        import torch.distributed as dist
        self.world, self.rank, self.n = dist.get_world_size(), dist.get_rank(), n
        own = []
        for nbytes in (n * 4, 256):                          # gradient buffer, flag array
            ptr, h = C.c_void_p(), (C.c_ubyte * 64)()
            if lib.shadow_p2p_alloc(nbytes, C.byref(ptr), h) != 0:
                own = None
                break
            own.append((ptr.value, bytes(h)))
        everyone = [None] * self.world
        dist.all_gather_object(everyone, None if own is None else (own[0][1], own[1][1]))      # every rank gets here whatever happened above
        if any(e is None for e in everyone):
            raise RuntimeError("shadow_p2p_alloc failed on some rank")
        self._own, self._opened = own, []
        gp, fp = [], []
        for r, (hg, hf) in enumerate(everyone):
            if r == self.rank:
                gp.append(own[0][0]); fp.append(own[1][0])
                continue
            ptrs = []
            for h in (hg, hf):
                ptr = C.c_void_p()
                check(lib.shadow_p2p_open((C.c_ubyte * 64).from_buffer_copy(h), C.byref(ptr)))
                ptrs.append(ptr.value)
                self._opened.append(ptr.value)
            gp.append(ptrs[0]); fp.append(ptrs[1])
        self.grad_ptrs = (C.c_uint64 * self.world)(*gp)
        self.flag_ptrs = (C.c_uint64 * self.world)(*fp)
        self.grad = torch.as_tensor(_RawDev(own[0][0], n, "<f4"), device=dev)
        self.state = torch.zeros(516, dtype=torch.int32, device=dev)
        # (the caller's all-reduce of the success flags is the barrier: nobody launches the exchange before every peer has opened every buffer)

    def error(self):
        """True when a poll of the exchange gave up waiting for a peer (the results of that step are garbage)"""
        return bool(int(self.state[2]))


class FlatAdamClip:
    """clip_grad_norm_(params, max_norm) + torch.optim.Adam.step (models.py:223-224) as one fused pass over a flat fp32 buffer.
    Parameters and gradients of the module are re-pointed into two flat buffers: the gradient buffer is also what the
    data-parallel run all-reduces (one NCCL call per step)."""

    def __init__(self, params, lr, max_norm=5.0, betas=(0.9, 0.999), eps=1e-8):
        self.params = [p for p in params]
        dev = self.params[0].device
        # every parameter starts on a 16-byte boundary (TMA descriptors of csrc/linear_tc.cu); the padding floats stay 0 in all four buffers
        n = sum((p.numel() + 3) // 4 * 4 for p in self.params)
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        # data-parallel run on GPUs: the gradient buffer is shared with the peers over NVLink and the exchange is fused into step()
        # (SHADOW_P2P=0: a local buffer, the caller all-reduces it with NCCL -- parallel.allreduce_flat_gradients / GraphedTrainer buckets)
        import torch.distributed as dist
        self.p2p = None
        if dev.type == "cuda" and dist.is_available() and dist.is_initialized() and 1 < dist.get_world_size() <= 16 and \
                os.environ.get("SHADOW_P2P", "1") != "0":
            try:
                self.p2p = _P2PGrad(n, dev)
            except Exception as e:                           # no peer access / IPC on this box: every rank must fall back together
                self.p2p, self._p2p_error = None, repr(e)
            ok = torch.tensor([1 if self.p2p is not None else 0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok) == 0:
                self.p2p = None
            else:
                self.gsum = torch.zeros(n, dtype=torch.float32, device=dev)
        self.grad = self.p2p.grad if self.p2p is not None else torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        self.offsets = {}                                    # id(parameter) -> first float of the parameter in the flat buffers
        for p in self.params:
            k = p.numel()
            self.offsets[id(p)] = off
            self.flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + k].view_as(p)
            p.grad = self.grad[off:off + k].view_as(p)
            off += (k + 3) // 4 * 4
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.planes = WeightPlanes(self.flat, self.params, self.offsets) if self.flat.is_cuda and _LINEAR == "tc" else None
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.sqnorm = torch.zeros(1, dtype=torch.float32, device=dev)
        self.lr, self.max_norm, self.betas, self.eps = lr, max_norm, betas, eps

    def zero_grad(self):
        join_wgrad()                                        # (a backward pass that died half-way leaves its side-stream work un-joined)
        if self.p2p is not None:                            # peers may still be reading the previous step's gradients
            p = self.p2p
            check(lib.shadow_p2p_zero_grad_f32(p.grad_ptrs, p.flag_ptrs, p.world, p.rank, self.grad.numel(), _p(p.state), _stream(self.grad)))
        else:
            self.grad.zero_()

    def step(self, grad_scale=1.0):
        if self.p2p is not None:                            # barrier + one-shot all-reduce over peer memory + norm, then Adam on the sum
            p = self.p2p
            join_wgrad()
            check(lib.shadow_p2p_adam_clip_step_f32(p.grad_ptrs, p.flag_ptrs, p.world, p.rank, _p(self.flat), _p(self.gsum), _p(self.exp_avg),
                                                    _p(self.exp_avg_sq), self.flat.numel(), 1.0 / p.world, self.max_norm, self.lr, self.betas[0],
                                                    self.betas[1], self.eps, _p(self.step_dev), _p(self.sqnorm), _p(p.state), _stream(self.flat)))
            if self.planes is not None:
                self.planes.refresh()
            return
        join_wgrad()
        check(lib.shadow_adam_clip_step_f32(_p(self.flat), _p(self.grad), _p(self.exp_avg), _p(self.exp_avg_sq), self.flat.numel(), grad_scale,
                                            self.max_norm, self.lr, self.betas[0], self.betas[1], self.eps, _p(self.step_dev), _p(self.sqnorm),
                                            _stream(self.flat)))
        if self.planes is not None:
            self.planes.refresh()

    def state_dict(self):
        return {"exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq, "step": self.step_dev, "lr": self.lr}

    def load_state_dict(self, sd):
        self.exp_avg.copy_(sd["exp_avg"]); self.exp_avg_sq.copy_(sd["exp_avg_sq"]); self.step_dev.copy_(sd["step"])
