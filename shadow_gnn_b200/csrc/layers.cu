// layers.cu -- message-passing kernels of the shaDow layers (shaDow/layers.py) on the block-diagonal batch the sampler
// leaves in HBM.  fp32 throughout (the reference is fp32; layer parity budget 1e-3 rel).  All kernels take the RAW edge
// layout of include/shadow_b200.h: row_span[i] = [start,end) into col[] / val[], columns are batch-global ids minus col_off.
// At training-batch sizes (n ~ 4,800 rows, e ~ 10^4 edges, D = 256) these kernels are L2-resident and launch-bound; they are
// written to be captured in one CUDA graph per step.  Dense Linear layers stay on cuBLAS (torch.nn.functional.linear).
#include <algorithm>

#include <cstring>

#include "common.cuh"

#define LAYER_BLOCK 256

// ------------------------------------------------------------------------------------------------
// adjacency values: dropedge + normalisation   (frontend/graph_utils.py:67-145, layers.py:512-522,584-600)
// ------------------------------------------------------------------------------------------------
// vals[p] = 1 for every edge of rows [0,n)
__global__ void fill_edge_vals_kernel(const int2 *__restrict__ row_span, int n, float *__restrict__ val, float x) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < n; i += gridDim.x * wpb) {
    const int2 sp = row_span[i];
    for (int p = sp.x + lane; p < sp.y; p += 32) val[p] = x;
  }
}
// adj_norm_rw / GAT / GIN dropedge: `num_drop` indices drawn WITH replacement, idx = floor(u * e) over the batch's e edges in
// CSR order (graph_utils.py:86-88).  The batch's edges are not contiguous in the raw layout, so the draw is over the
// ordinal position and mapped through edge_ord (exclusive prefix of the row lengths).
__global__ void dropedge_kernel(const int2 *__restrict__ row_span, const int *__restrict__ row_ord /*[n+1] exclusive prefix of row lengths*/,
                                int n, float p, uint32_t seed, const uint32_t *__restrict__ step_dev, float *__restrict__ val) {
  const int e = row_ord[n];
  const int num_drop = (int)((double)e * (double)p);                              // int(e * dropedge), graph_utils.py:86
  const uint32_t step = *step_dev;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < num_drop; t += gridDim.x * blockDim.x) {
    const uint32_t r = philox4x32_10_x((uint32_t)t, step, 0x5eed0001u, 0u, seed, 0x243F6A88u);
    int k = (int)(((unsigned long long)r * (unsigned long long)e) >> 32);          // floor(u * e), u in [0,1)
    int lo = 0, hi = n;                                                           // row whose ordinal range holds k
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (row_ord[mid] <= k) lo = mid; else hi = mid; }
    val[row_span[lo].x + (k - row_ord[lo])] = 0.f;
  }
}
// mode 0: rw   vals /= clamp(rowsum(vals), 1)                         (adj_norm_rw, graph_utils.py:89-94)
// mode 1: gin  vals *= deg_orig / clamp(rowsum(vals), 1)              (layers.py:514-522)
__global__ void row_normalize_kernel(const int2 *__restrict__ row_span, int n, int mode, float *__restrict__ val) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < n; i += gridDim.x * wpb) {
    const int2 sp = row_span[i];
    float s = 0.f;
    for (int p = sp.x + lane; p < sp.y; p += 32) s += val[p];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    const float deg = fmaxf(s, 1.f);
    const float f = (mode == 0) ? 1.f / deg : (float)(sp.y - sp.x) / deg;
    for (int p = sp.x + lane; p < sp.y; p += 32) val[p] = (mode == 0) ? val[p] / deg : val[p] * f;
  }
}
// adj_norm_sym (graph_utils.py:109-145): optional symmetric survival (edge kept iff both directions were kept), then
// D^-1/2 A D^-1/2 with D = clip(rowsum, 1).  Two kernels: survive+degree, then scale.
__global__ void sym_survive_kernel(const int2 *__restrict__ row_span, const int *__restrict__ col, int col_off, int n,
                                   const float *__restrict__ mask /*after dropedge*/, float *__restrict__ val, float *__restrict__ deg) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < n; i += gridDim.x * wpb) {
    const int2 sp = row_span[i];
    float s = 0.f;
    for (int p = sp.x + lane; p < sp.y; p += 32) {
      float v = mask[p];
      if (v != 0.f) {                                 // reverse edge (j -> i): binary search in the sorted row j
        const int j = col[p] - col_off;
        const int2 sj = row_span[j];
        int lo = sj.x, hi = sj.y;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (col[mid] - col_off < i) lo = mid + 1; else hi = mid; }
        v = (lo < sj.y && col[lo] - col_off == i) ? mask[lo] : 0.f;
      }
      val[p] = v; s += v;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (lane == 0) deg[i] = fmaxf(s, 1.f);
  }
}
__global__ void row_degree_kernel(const int2 *__restrict__ row_span, int n, const float *__restrict__ val, float *__restrict__ deg) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < n; i += gridDim.x * wpb) {
    const int2 sp = row_span[i];
    float s = 0.f;
    for (int p = sp.x + lane; p < sp.y; p += 32) s += val[p];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (lane == 0) deg[i] = fmaxf(s, 1.f);
  }
}
__global__ void sym_scale_kernel(const int2 *__restrict__ row_span, const int *__restrict__ col, int col_off, int n,
                                 const float *__restrict__ deg, float *__restrict__ val) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < n; i += gridDim.x * wpb) {
    const int2 sp = row_span[i];
    const float di = rsqrtf(deg[i]);
    for (int p = sp.x + lane; p < sp.y; p += 32) val[p] = di * val[p] * rsqrtf(deg[col[p] - col_off]);
  }
}

// ------------------------------------------------------------------------------------------------
// CSR SpMM (torch.sparse.mm on uncoalesced COO, layers.py:326-327,433,475,523): duplicates are summed, nothing is deduped
//   fwd : Y[i,:]  = sum_p val[p] * X[col[p],:]          warp per row, lanes across the feature dimension (float4)
//   bwd : dX[c,:] += val[p] * dY[i,:]                   same traversal, red.global.add (A^T without building a CSC)
// ------------------------------------------------------------------------------------------------
// NV = float4 chunks per lane (F <= 128*NV): the whole output row lives in registers, col[]/val[] are read once, and the
// edge loop is unrolled by 4 so that a row of the batch with ~150 in-scope neighbours is not a 150-deep latency chain.
// accumulate val[p] * X[col[p],:] for p = p_begin, p_begin + p_stride, ... < p_end in chunks of `chunk` consecutive edges
// (chunk <= 32): one coalesced load brings the (col, val) pairs of a chunk, they are broadcast by shuffle so that the X-row loads
// depend on nothing in memory and up to UB of them are in flight per lane
template <int NV>
__device__ __forceinline__ void spmm_accumulate(float4 (&acc)[NV], const int *__restrict__ col, const int col_off, const float *__restrict__ val,
                                                const float *__restrict__ X, const int F, const int F4, const int p_begin, const int p_end,
                                                const int chunk, const int p_stride, const int lane) {
  constexpr int UB = NV <= 2 ? 8 : (NV == 4 ? 4 : 2);           // X-row loads in flight per lane = UB * NV float4
  for (int p0 = p_begin; p0 < p_end; p0 += p_stride) {
    const int cnt = min(chunk, p_end - p0);
    const int pl = p0 + lane;
    const int c_l = lane < cnt ? col[pl] - col_off : 0;
    const float w_l = lane < cnt ? (val ? val[pl] : 1.f) : 0.f;
    for (int u0 = 0; u0 < cnt; u0 += UB) {
      float4 x[UB][NV];
      float w[UB];
#pragma unroll
      for (int u = 0; u < UB; u++) {
        const int c = __shfl_sync(0xffffffffu, c_l, (u0 + u) & 31);
        w[u] = (u0 + u < cnt) ? __shfl_sync(0xffffffffu, w_l, (u0 + u) & 31) : 0.f;
        const float4 *xr = reinterpret_cast<const float4 *>(X + (size_t)c * F);
#pragma unroll
        for (int k = 0; k < NV; k++) {
          const int f = lane + 32 * k;
          x[u][k] = (f < F4 && w[u] != 0.f) ? xr[f] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int u = 0; u < UB; u++)
#pragma unroll
        for (int k = 0; k < NV; k++) { acc[k].x += w[u] * x[u][k].x; acc[k].y += w[u] * x[u][k].y; acc[k].z += w[u] * x[u][k].z; acc[k].w += w[u] * x[u][k].w; }
    }
  }
}

// Y = beta*Y + A X.  One warp per row, the whole output row in registers -- except for LONG rows: in a shaDow batch the root's row holds
// ~140 in-scope neighbours while the other rows hold one or two, and a single warp walking 140 X rows eight at a time is the critical
// path of the whole launch (~18 dependent L2 round trips).  Rows longer than SPMM_LONG edges are therefore left to the end of the CTA's
// row group and split over all of its warps (8 consecutive edges per warp and round), partial rows summed through shared memory in a
// fixed order (deterministic).  NV <= 2 (F <= 256); wider rows keep the one-warp path.
#define SPMM_LONG 24
template <int NV>
__global__ void __launch_bounds__(LAYER_BLOCK) spmm_fwd_vec_kernel(const int2 *__restrict__ row_span, const int *__restrict__ col, int col_off,
                                                                   const float *__restrict__ val, const float *__restrict__ X, float *__restrict__ Y,
                                                                   int n, int F, float beta /*Y = beta*Y + A X*/) {
  constexpr bool SPLIT = NV <= 2;
  __shared__ float4 part[SPLIT ? LAYER_BLOCK / 32 : 1][SPLIT ? 32 * NV : 1];
  __shared__ int long_rows[LAYER_BLOCK / 32];
  __shared__ int n_long;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int F4 = F >> 2;
  for (int base = blockIdx.x * wpb; base < n; base += gridDim.x * wpb) {
    if (SPLIT) { if (threadIdx.x == 0) n_long = 0; __syncthreads(); }
    const int i = base + warp;
    if (i < n) {
      const int2 sp = row_span[i];
      if (SPLIT && sp.y - sp.x > SPMM_LONG) { if (lane == 0) long_rows[atomicAdd(&n_long, 1)] = i; }
      else {
        float4 acc[NV];
#pragma unroll
        for (int k = 0; k < NV; k++) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        spmm_accumulate<NV>(acc, col, col_off, val, X, F, F4, sp.x, sp.y, 32, 32, lane);
#pragma unroll
        for (int k = 0; k < NV; k++) {
          const int f = lane + 32 * k;
          if (f < F4) {
            float4 *y = reinterpret_cast<float4 *>(Y + (size_t)i * F) + f;
            float4 r = acc[k];
            if (beta != 0.f) { const float4 o = *y; r.x += beta * o.x; r.y += beta * o.y; r.z += beta * o.z; r.w += beta * o.w; }
            *y = r;
          }
        }
      }
    }
    if (SPLIT) {
      __syncthreads();
      const int nl = n_long;
      for (int q = 0; q < nl; q++) {
        // rows are taken in ascending order so that the result does not depend on which warp registered a row first
        int r = 0;                                       // q-th smallest entry of long_rows (nl <= 8)
        for (int a = 0; a < nl; a++) { int rank = 0; for (int b2 = 0; b2 < nl; b2++) rank += long_rows[b2] < long_rows[a]; if (rank == q) r = long_rows[a]; }
        const int2 sp = row_span[r];
        float4 acc[NV];
#pragma unroll
        for (int k = 0; k < NV; k++) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        spmm_accumulate<NV>(acc, col, col_off, val, X, F, F4, sp.x + warp * 8, sp.y, 8, 8 * wpb, lane);
#pragma unroll
        for (int k = 0; k < NV; k++) part[warp][lane + 32 * k] = acc[k];
        __syncthreads();
        for (int f = threadIdx.x; f < F4; f += blockDim.x) {
          float4 rsum = part[0][f];
          for (int w2 = 1; w2 < wpb; w2++) { const float4 o = part[w2][f]; rsum.x += o.x; rsum.y += o.y; rsum.z += o.z; rsum.w += o.w; }
          float4 *y = reinterpret_cast<float4 *>(Y + (size_t)r * F) + f;
          if (beta != 0.f) { const float4 o = *y; rsum.x += beta * o.x; rsum.y += beta * o.y; rsum.z += beta * o.z; rsum.w += beta * o.w; }
          *y = rsum;
        }
        __syncthreads();
      }
    }
  }
}
__global__ void __launch_bounds__(LAYER_BLOCK) spmm_fwd_scalar_kernel(const int2 *__restrict__ row_span, const int *__restrict__ col, int col_off,
                                                                      const float *__restrict__ val, const float *__restrict__ X, float *__restrict__ Y,
                                                                      int n, int F, float beta) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < n; i += gridDim.x * wpb) {
    const int2 sp = row_span[i];
    for (int f = lane; f < F; f += 32) {
      float acc = 0.f;
      for (int p = sp.x; p < sp.y; p++) acc += (val ? val[p] : 1.f) * X[(size_t)(col[p] - col_off) * F + f];
      Y[(size_t)i * F + f] = acc + (beta != 0.f ? beta * Y[(size_t)i * F + f] : 0.f);
    }
  }
}
// backward: dX[c,:] += val[p] * dY[i,:] with 128-bit vector reductions (red.global.add.v4.f32, sm_90+).  The dY row lives in registers
// (F <= 256: two float4 per lane); a row longer than 32 edges (the root's ~140 in-scope neighbours) is not walked by one warp: it is queued
// and all warps of the CTA take 32-edge pieces of it afterwards -- the kernel's duration used to be that one warp's serial walk.
template <bool VEC>
__global__ void __launch_bounds__(LAYER_BLOCK) spmm_bwd_kernel(const int2 *__restrict__ row_span, const int *__restrict__ col, int col_off,
                                                               const float *__restrict__ val, const float *__restrict__ dY, float *__restrict__ dX,
                                                               int n, int F) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5, warp = threadIdx.x >> 5;
  __shared__ int long_rows[LAYER_BLOCK / 32];
  __shared__ int n_long;
  if (VEC) {
    const int F4 = F >> 2;
    const bool small = F4 <= 64;
    auto walk = [&](const int i, const int p_lo, const int p_hi) {      // edges [p_lo, p_hi) of row i
      float4 g0 = make_float4(0.f, 0.f, 0.f, 0.f), g1 = g0;
      if (small) {
        if (lane < F4) g0 = reinterpret_cast<const float4 *>(dY + (size_t)i * F)[lane];
        if (lane + 32 < F4) g1 = reinterpret_cast<const float4 *>(dY + (size_t)i * F)[lane + 32];
      }
      for (int p0 = p_lo; p0 < p_hi; p0 += 32) {
        const int pl = p0 + lane;
        const int c_l = pl < p_hi ? col[pl] - col_off : 0;
        const float w_l = pl < p_hi ? (val ? val[pl] : 1.f) : 0.f;
        const int cnt = min(32, p_hi - p0);
        for (int u = 0; u < cnt; u++) {                        // shuffles outside the feature loop: every lane takes part
          const int c = __shfl_sync(0xffffffffu, c_l, u);
          const float w = __shfl_sync(0xffffffffu, w_l, u);
          if (w == 0.f) continue;
          float4 *dst = reinterpret_cast<float4 *>(dX + (size_t)c * F);
          if (small) {
            if (lane < F4) atomicAdd(dst + lane, make_float4(w * g0.x, w * g0.y, w * g0.z, w * g0.w));
            if (lane + 32 < F4) atomicAdd(dst + lane + 32, make_float4(w * g1.x, w * g1.y, w * g1.z, w * g1.w));
          } else {
            for (int f = lane; f < F4; f += 32) {
              const float4 g = reinterpret_cast<const float4 *>(dY + (size_t)i * F)[f];
              atomicAdd(dst + f, make_float4(w * g.x, w * g.y, w * g.z, w * g.w));
            }
          }
        }
      }
    };
    for (int base = blockIdx.x * wpb; base < n; base += gridDim.x * wpb) {
      if (threadIdx.x == 0) n_long = 0;
      __syncthreads();
      const int i = base + warp;
      if (i < n) {
        const int2 sp = row_span[i];
        if (sp.y - sp.x > 32) { if (lane == 0) long_rows[atomicAdd(&n_long, 1)] = i; }
        else walk(i, sp.x, sp.y);
      }
      __syncthreads();
      const int nl = n_long;
      for (int r = 0; r < nl; r++) {
        const int il = long_rows[r];
        const int2 sp = row_span[il];
        for (int p = sp.x + 32 * warp; p < sp.y; p += 32 * wpb) walk(il, p, min(p + 32, sp.y));
      }
      __syncthreads();
    }
  } else {
    for (int i = blockIdx.x * wpb + warp; i < n; i += gridDim.x * wpb) {
      const int2 sp = row_span[i];
      for (int f = lane; f < F; f += 32) {
        const float g = dY[(size_t)i * F + f];
        for (int p = sp.x; p < sp.y; p++) {
          const float w = val ? val[p] : 1.f;
          if (w != 0.f) atomicAdd(dX + (size_t)(col[p] - col_off) * F + f, w * g);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// activation + norm_feat (layers.py:329-338, F_ACT layers.py:26-39), one warp per row:
//   a = act(z);  out = (a - mean(a)) * scale * rsqrt(var_biased(a) + 1e-9) + offset   [+ out_prev if accumulate]
// ------------------------------------------------------------------------------------------------
enum { ACT_RELU = 0, ACT_I = 1, ACT_ELU = 2, ACT_TANH = 3, ACT_LRELU = 4 };
__device__ __forceinline__ float act_f(float z, int act) {
  switch (act) {
    case ACT_RELU: return z > 0.f ? z : 0.f;
    case ACT_ELU: return z > 0.f ? z : expm1f(z);
    case ACT_TANH: return tanhf(z);
    case ACT_LRELU: return z > 0.f ? z : 0.2f * z;
    default: return z;                       // "I" = LeakyReLU(negative_slope=1)
  }
}
__device__ __forceinline__ float act_df(float z, float a, int act) {
  switch (act) {
    case ACT_RELU: return z > 0.f ? 1.f : 0.f;
    case ACT_ELU: return z > 0.f ? 1.f : a + 1.f;
    case ACT_TANH: return 1.f - a * a;
    case ACT_LRELU: return z > 0.f ? 1.f : 0.2f;
    default: return 1.f;
  }
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}
#define NORM_MAX_PER_LANE 32      // D <= 1024

// NPL = columns per lane (D <= 32*NPL): the row stays in registers
template <int NPL>
__global__ void __launch_bounds__(LAYER_BLOCK) act_norm_fwd_kernel(const float *__restrict__ Z, int ldz, const float *__restrict__ scale,
                                                                   const float *__restrict__ offset, float *__restrict__ out, int ldo,
                                                                   float *__restrict__ mean_out, float *__restrict__ rstd_out, int n, int D,
                                                                   int act, int do_norm, int accumulate) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < n; i += gridDim.x * wpb) {
    float a[NPL];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NPL; k++) { const int f = lane + 32 * k; a[k] = (f < D) ? act_f(Z[(size_t)i * ldz + f], act) : 0.f; s += a[k]; }
    float mean = 0.f, rstd = 1.f;
    if (do_norm) {
      mean = warp_sum(s) / (float)D;
      float v = 0.f;
#pragma unroll
      for (int k = 0; k < NPL; k++) { const float d = (lane + 32 * k < D) ? a[k] - mean : 0.f; v += d * d; }
      rstd = rsqrtf(warp_sum(v) / (float)D + 1e-9f);
      if (lane == 0) { mean_out[i] = mean; rstd_out[i] = rstd; }
    }
#pragma unroll
    for (int k = 0; k < NPL; k++) {
      const int f = lane + 32 * k;
      if (f < D) {
        float o = do_norm ? (a[k] - mean) * scale[f] * rstd + offset[f] : a[k];
        if (accumulate) o += out[(size_t)i * ldo + f];
        out[(size_t)i * ldo + f] = o;
      }
    }
  }
}
// dZ = d act . d norm ; dscale / doffset / dbias(=column sums of dZ) accumulated per warp in registers, then one shared-memory
// atomic per column per warp and one global atomic per column per CTA
template <int NPL>
__global__ void __launch_bounds__(LAYER_BLOCK) act_norm_bwd_kernel(const float *__restrict__ dOut, int ldo, const float *__restrict__ Z, int ldz,
                                                                   const float *__restrict__ scale, const float *__restrict__ mean_in,
                                                                   const float *__restrict__ rstd_in, float *__restrict__ dZ, int lddz,
                                                                   float *__restrict__ dscale, float *__restrict__ doffset, float *__restrict__ dbias,
                                                                   int n, int D, int act, int do_norm) {
  extern __shared__ float sh[];               // [warps][3*D] column partials of every warp: dscale, doffset, dbias (summed in a fixed order)
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  float ps[NPL], po[NPL], pb[NPL];
#pragma unroll
  for (int k = 0; k < NPL; k++) { ps[k] = 0.f; po[k] = 0.f; pb[k] = 0.f; }
  for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < n; i += gridDim.x * wpb) {
    float a[NPL], g[NPL], z[NPL];
#pragma unroll
    for (int k = 0; k < NPL; k++) {
      const int f = lane + 32 * k;
      z[k] = (f < D) ? Z[(size_t)i * ldz + f] : 0.f; a[k] = act_f(z[k], act); g[k] = (f < D) ? dOut[(size_t)i * ldo + f] : 0.f;
    }
    if (do_norm) {
      const float mean = mean_in[i], rstd = rstd_in[i];
      float s1 = 0.f, s2 = 0.f, xh[NPL], dxh[NPL];
#pragma unroll
      for (int k = 0; k < NPL; k++) {
        const int f = lane + 32 * k;
        xh[k] = (f < D) ? (a[k] - mean) * rstd : 0.f; dxh[k] = (f < D) ? g[k] * scale[f] : 0.f;
        s1 += dxh[k]; s2 += dxh[k] * xh[k];
        ps[k] += g[k] * xh[k]; po[k] += g[k];
      }
      s1 = warp_sum(s1) / (float)D; s2 = warp_sum(s2) / (float)D;
#pragma unroll
      for (int k = 0; k < NPL; k++) {
        const int f = lane + 32 * k;
        if (f < D) { const float dz = rstd * (dxh[k] - s1 - xh[k] * s2) * act_df(z[k], a[k], act); dZ[(size_t)i * lddz + f] = dz; pb[k] += dz; }
      }
    } else {
#pragma unroll
      for (int k = 0; k < NPL; k++) { const int f = lane + 32 * k; if (f < D) { const float dz = g[k] * act_df(z[k], a[k], act); dZ[(size_t)i * lddz + f] = dz; pb[k] += dz; } }
    }
  }
  float *mine = sh + (size_t)(threadIdx.x >> 5) * 3 * D;
#pragma unroll
  for (int k = 0; k < NPL; k++) {
    const int f = lane + 32 * k;
    if (f < D) { mine[f] = ps[k]; mine[D + f] = po[k]; mine[2 * D + f] = pb[k]; }
  }
  __syncthreads();
  for (int f = threadIdx.x; f < 3 * D; f += blockDim.x) {
    float v = 0.f;
    for (int w = 0; w < wpb; w++) v += sh[(size_t)w * 3 * D + f];
    if (f < D) { if (do_norm) atomicAdd(&dscale[f], v); }
    else if (f < 2 * D) { if (do_norm) atomicAdd(&doffset[f - D], v); }
    else if (dbias) atomicAdd(&dbias[f - 2 * D], v);
  }
}

// The same backward for up to TWO branches that share dOut (the self and the neighbour branch of a GraphSAGE layer: out = n0(act(Z0)) + n1(act(Z1)),
// layers.py:474-483): dOut is read once, and the column sums (dscale, doffset, dbias per branch) are reduced in two deterministic stages --
// every CTA writes its partial sums to `partials[cta][branch][3][D]`, colsum_finish_kernel adds them up in CTA order.  No atomics.
struct AnbPair {
  const float *Z[2], *scale[2], *mean[2], *rstd[2];
  float *dZ[2], *dscale[2], *doffset[2], *dbias[2];
};
template <int ACT>
__device__ __forceinline__ float act_ft(float z) {
  if (ACT == ACT_RELU) return z > 0.f ? z : 0.f;
  if (ACT == ACT_ELU) return z > 0.f ? z : expm1f(z);
  if (ACT == ACT_TANH) return tanhf(z);
  if (ACT == ACT_LRELU) return z > 0.f ? z : 0.2f * z;
  return z;
}
template <int ACT>
__device__ __forceinline__ float act_dft(float z, float a) {
  if (ACT == ACT_RELU) return z > 0.f ? 1.f : 0.f;
  if (ACT == ACT_ELU) return z > 0.f ? 1.f : a + 1.f;
  if (ACT == ACT_TANH) return 1.f - a * a;
  if (ACT == ACT_LRELU) return z > 0.f ? 1.f : 0.2f;
  return 1.f;
}
// NV float4 per lane (D <= 128 * NV, D % 4 == 0): lane l owns features 4 * (l + 32 v) .. + 3
template <int NV, int NB, int ACT, bool NORM>
__global__ void __launch_bounds__(LAYER_BLOCK) act_norm_bwd_pair_kernel(const float *__restrict__ dOut, int ldo, const AnbPair A, int ldz, int lddz, int n, int D,
                                                                        float *__restrict__ partials) {
  extern __shared__ float sh[];               // [warps][NB][3][D]
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5, warp = threadIdx.x >> 5;
  float ps[NB][NV][4], po[NB][NV][4], pb[NB][NV][4], sc[NB][NV][4];
  bool on[NV];
#pragma unroll
  for (int v = 0; v < NV; v++) on[v] = 4 * (lane + 32 * v) < D;
#pragma unroll
  for (int b = 0; b < NB; b++)
#pragma unroll
    for (int v = 0; v < NV; v++) {
      float4 s4 = make_float4(1.f, 1.f, 1.f, 1.f);
      if (NORM && on[v]) s4 = reinterpret_cast<const float4 *>(A.scale[b])[lane + 32 * v];
      sc[b][v][0] = s4.x; sc[b][v][1] = s4.y; sc[b][v][2] = s4.z; sc[b][v][3] = s4.w;
#pragma unroll
      for (int e = 0; e < 4; e++) { ps[b][v][e] = 0.f; po[b][v][e] = 0.f; pb[b][v][e] = 0.f; }
    }
  for (int i = blockIdx.x * wpb + warp; i < n; i += gridDim.x * wpb) {
    float g[NV][4];
#pragma unroll
    for (int v = 0; v < NV; v++) {
      float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (on[v]) g4 = reinterpret_cast<const float4 *>(dOut + (size_t)i * ldo)[lane + 32 * v];
      g[v][0] = g4.x; g[v][1] = g4.y; g[v][2] = g4.z; g[v][3] = g4.w;
    }
    float4 z4[NB][NV];
#pragma unroll
    for (int b = 0; b < NB; b++)
#pragma unroll
      for (int v = 0; v < NV; v++) z4[b][v] = on[v] ? reinterpret_cast<const float4 *>(A.Z[b] + (size_t)i * ldz)[lane + 32 * v] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int b = 0; b < NB; b++) {
      float z[NV][4], a[NV][4], dz[NV][4];
#pragma unroll
      for (int v = 0; v < NV; v++) {
        z[v][0] = z4[b][v].x; z[v][1] = z4[b][v].y; z[v][2] = z4[b][v].z; z[v][3] = z4[b][v].w;
#pragma unroll
        for (int e = 0; e < 4; e++) a[v][e] = act_ft<ACT>(z[v][e]);
      }
      if (NORM) {
        const float mean = A.mean[b][i], rstd = A.rstd[b][i];
        float s1 = 0.f, s2 = 0.f, xh[NV][4], dxh[NV][4];
#pragma unroll
        for (int v = 0; v < NV; v++)
#pragma unroll
          for (int e = 0; e < 4; e++) {
            xh[v][e] = on[v] ? (a[v][e] - mean) * rstd : 0.f; dxh[v][e] = g[v][e] * sc[b][v][e];
            s1 += dxh[v][e]; s2 += dxh[v][e] * xh[v][e];
            ps[b][v][e] += g[v][e] * xh[v][e]; po[b][v][e] += g[v][e];
          }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
        s1 /= (float)D; s2 /= (float)D;
#pragma unroll
        for (int v = 0; v < NV; v++)
#pragma unroll
          for (int e = 0; e < 4; e++) dz[v][e] = rstd * (dxh[v][e] - s1 - xh[v][e] * s2) * act_dft<ACT>(z[v][e], a[v][e]);
      } else {
#pragma unroll
        for (int v = 0; v < NV; v++)
#pragma unroll
          for (int e = 0; e < 4; e++) dz[v][e] = g[v][e] * act_dft<ACT>(z[v][e], a[v][e]);
      }
#pragma unroll
      for (int v = 0; v < NV; v++) {
        if (on[v]) {
          reinterpret_cast<float4 *>(A.dZ[b] + (size_t)i * lddz)[lane + 32 * v] = make_float4(dz[v][0], dz[v][1], dz[v][2], dz[v][3]);
#pragma unroll
          for (int e = 0; e < 4; e++) pb[b][v][e] += dz[v][e];
        }
      }
    }
  }
  float *mine = sh + (size_t)warp * NB * 3 * D;
#pragma unroll
  for (int b = 0; b < NB; b++)
#pragma unroll
    for (int v = 0; v < NV; v++)
      if (on[v]) {
        const int f4 = lane + 32 * v;
        reinterpret_cast<float4 *>(mine + (b * 3 + 0) * D)[f4] = make_float4(ps[b][v][0], ps[b][v][1], ps[b][v][2], ps[b][v][3]);
        reinterpret_cast<float4 *>(mine + (b * 3 + 1) * D)[f4] = make_float4(po[b][v][0], po[b][v][1], po[b][v][2], po[b][v][3]);
        reinterpret_cast<float4 *>(mine + (b * 3 + 2) * D)[f4] = make_float4(pb[b][v][0], pb[b][v][1], pb[b][v][2], pb[b][v][3]);
      }
  __syncthreads();
  for (int f = threadIdx.x; f < NB * 3 * D; f += blockDim.x) {
    float v = 0.f;
    for (int w = 0; w < wpb; w++) v += sh[(size_t)w * NB * 3 * D + f];
    partials[(size_t)blockIdx.x * NB * 3 * D + f] = v;
  }
}
// 32 columns of [NB][3][D] per CTA: 8 groups of threads each add every 8th CTA's partial, the groups are then summed in a fixed order, and the
// total is ADDED to its destination (each destination has exactly one writer): deterministic
__global__ void __launch_bounds__(256) colsum_finish_kernel(const float *__restrict__ partials, int nparts, int D, int nb, const AnbPair A, int do_norm) {
  __shared__ float red[8][32];
  const int cols = nb * 3 * D, c = blockIdx.x * 32 + (threadIdx.x & 31), grp = threadIdx.x >> 5;
  float v = 0.f;
  if (c < cols) {
#pragma unroll 4
    for (int p = grp; p < nparts; p += 8) v += partials[(size_t)p * cols + c];
  }
  red[grp][threadIdx.x & 31] = v;
  __syncthreads();
  if (grp == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; k++) t += red[k][threadIdx.x];
    const int b = c / (3 * D), kind = (c / D) % 3, f = c % D;
    float *dst = kind == 0 ? (do_norm ? A.dscale[b] : nullptr) : (kind == 1 ? (do_norm ? A.doffset[b] : nullptr) : A.dbias[b]);
    if (dst) dst[f] += t;
  }
}

// ------------------------------------------------------------------------------------------------
// GAT attention aggregation (layers.py:560-582), all heads in one launch, warp per (row, head):
//   e_ij = a_self[i,k] + a_neigh[j,k];  u_ij = exp(e_ij - max_j e_ij) * A_ij;  out[i,k,:] = sum_j u_ij h[j,k,:] / clamp(sum_j u_ij, 1e-10)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LAYER_BLOCK) gat_fwd_kernel(const int2 *__restrict__ row_span, const int *__restrict__ col, int col_off,
                                                              const float *__restrict__ val, const float *__restrict__ a_self,
                                                              const float *__restrict__ a_neigh, const float *__restrict__ H, float *__restrict__ out,
                                                              float *__restrict__ rowmax, float *__restrict__ denom, int n, int heads, int d) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int D = heads * d;
  for (int w = blockIdx.x * wpb + (threadIdx.x >> 5); w < n * heads; w += gridDim.x * wpb) {
    const int i = w / heads, k = w - i * heads;
    const int2 sp = row_span[i];
    const float as = a_self[(size_t)i * heads + k];
    float mx = -INFINITY;
    for (int p = sp.x + lane; p < sp.y; p += 32) mx = fmaxf(mx, as + a_neigh[(size_t)(col[p] - col_off) * heads + k]);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s));
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};       // d <= 256
    float den = 0.f;
    for (int p = sp.x; p < sp.y; p++) {
      const int j = col[p] - col_off;
      const float u = expf(as + a_neigh[(size_t)j * heads + k] - mx) * (val ? val[p] : 1.f);
      den += u;
      int c = 0;
      for (int f = lane; f < d; f += 32, c++) acc[c] += u * H[(size_t)j * D + k * d + f];
    }
    const float S = fmaxf(den, 1e-10f);
    int c = 0;
    for (int f = lane; f < d; f += 32, c++) out[(size_t)i * D + k * d + f] = acc[c] / S;
    if (lane == 0) { rowmax[w] = mx; denom[w] = den; }
  }
}
__global__ void __launch_bounds__(LAYER_BLOCK) gat_bwd_kernel(const int2 *__restrict__ row_span, const int *__restrict__ col, int col_off,
                                                              const float *__restrict__ val, const float *__restrict__ a_self,
                                                              const float *__restrict__ a_neigh, const float *__restrict__ H,
                                                              const float *__restrict__ out, const float *__restrict__ rowmax,
                                                              const float *__restrict__ denom, const float *__restrict__ dOut,
                                                              float *__restrict__ dH, float *__restrict__ da_self, float *__restrict__ da_neigh,
                                                              int n, int heads, int d) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int D = heads * d;
  for (int w = blockIdx.x * wpb + (threadIdx.x >> 5); w < n * heads; w += gridDim.x * wpb) {
    const int i = w / heads, k = w - i * heads;
    const int2 sp = row_span[i];
    const float as = a_self[(size_t)i * heads + k], mx = rowmax[w], den = denom[w];
    const bool clamped = den < 1e-10f;
    const float S = fmaxf(den, 1e-10f);
    float g[8], go = 0.f;
    int c = 0;
    for (int f = lane; f < d; f += 32, c++) { g[c] = dOut[(size_t)i * D + k * d + f]; go += g[c] * out[(size_t)i * D + k * d + f]; }
    go = warp_sum(go);
    float das = 0.f;
    for (int p = sp.x; p < sp.y; p++) {
      const int j = col[p] - col_off;
      const float u = expf(as + a_neigh[(size_t)j * heads + k] - mx) * (val ? val[p] : 1.f);
      if (u == 0.f) continue;
      const float alpha = u / S;
      float gh = 0.f;
      c = 0;
      for (int f = lane; f < d; f += 32, c++) {
        gh += g[c] * H[(size_t)j * D + k * d + f];
        atomicAdd(dH + (size_t)j * D + k * d + f, alpha * g[c]);
      }
      gh = warp_sum(gh);
      const float de = alpha * (gh - (clamped ? 0.f : go));      // the row max is shift-invariant: no gradient through it
      das += de;
      if (lane == 0) atomicAdd(da_neigh + (size_t)j * heads + k, de);
    }
    if (lane == 0) da_self[(size_t)i * heads + k] = das;
  }
}

// ------------------------------------------------------------------------------------------------
// per-subgraph pooling (F.embedding_bag in ResPool, layers.py:168-184): out[s,:] = reduce over rows [seg[s], seg[s+1])
// mode 0 sum, 1 mean, 2 max (argmax row kept for the backward)
// ------------------------------------------------------------------------------------------------
__global__ void segment_pool_fwd_kernel(const float *__restrict__ X, const int *__restrict__ seg, int seg_off, int S, int F, int mode,
                                        float *__restrict__ out, int *__restrict__ argmax) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < (long long)S * F; t += (long long)gridDim.x * blockDim.x) {
    const int s = (int)(t / F), f = (int)(t - (long long)s * F);
    const int r0 = seg[s] - seg_off, r1 = seg[s + 1] - seg_off;
    if (mode == 2) {
      float m = -INFINITY; int am = r0;
      for (int r = r0; r < r1; r++) { const float x = X[(size_t)r * F + f]; if (x > m) { m = x; am = r; } }
      out[t] = (r1 > r0) ? m : 0.f; argmax[t] = am;
    } else {
      float a = 0.f;
      for (int r = r0; r < r1; r++) a += X[(size_t)r * F + f];
      out[t] = (mode == 1 && r1 > r0) ? a / (float)(r1 - r0) : a;
    }
  }
}
__global__ void segment_pool_bwd_kernel(const float *__restrict__ dOut, const int *__restrict__ seg, int seg_off, int S, int F, int mode,
                                        const int *__restrict__ argmax, float *__restrict__ dX) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < (long long)S * F; t += (long long)gridDim.x * blockDim.x) {
    const int s = (int)(t / F), f = (int)(t - (long long)s * F);
    const int r0 = seg[s] - seg_off, r1 = seg[s + 1] - seg_off;
    const float g = dOut[t];
    if (mode == 2) { if (r1 > r0) dX[(size_t)argmax[t] * F + f] = g; }
    else { const float v = (mode == 1 && r1 > r0) ? g / (float)(r1 - r0) : g; for (int r = r0; r < r1; r++) dX[(size_t)r * F + f] = v; }
  }
}

// ------------------------------------------------------------------------------------------------
// optimizer step of DeepGNN.step (models.py:219-224): clip_grad_norm_(params, 5) + Adam, over ONE flat fp32 buffer
// (the same buffer NCCL all-reduces in the data-parallel run).  Two launches: squared norm, then the update.
// ------------------------------------------------------------------------------------------------
__global__ void sqnorm_kernel(const float *__restrict__ g, long long n, float scale, float *__restrict__ out) {
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) { const float x = g[i] * scale; s += x * x; }
  s = warp_sum(s);
  __shared__ float sh[32];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
    s = warp_sum(s);
    if (threadIdx.x == 0) atomicAdd(out, s);
  }
}
__global__ void adam_clip_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v, long long n,
                                 const float *__restrict__ sqnorm, float gscale, float max_norm, float lr, float b1, float b2, float eps,
                                 const int *__restrict__ step_dev) {
  const float total = sqrtf(*sqnorm);
  const float clip = fminf(max_norm / (total + 1e-6f), 1.f) * gscale;               // torch.nn.utils.clip_grad_norm_
  const int t = *step_dev;
  const float bc1 = 1.f - powf(b1, (float)t), bc2 = 1.f - powf(b2, (float)t);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * clip;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    p[i] -= lr / bc1 * mi / (sqrtf(vi) / sqrtf(bc2) + eps);                          // torch.optim.Adam (no amsgrad, no weight decay)
  }
}
__global__ void bump_step_kernel(int *step_dev) { *step_dev += 1; }

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
static inline int grid_for(long long work_items, int per_block, int cap) { return (int)std::max<long long>(1, std::min<long long>((work_items + per_block - 1) / per_block, cap)); }
#define ST(s) ((cudaStream_t)(s))
#define WPB (LAYER_BLOCK / 32)

extern "C" int shadow_edge_vals_fill(const int32_t *row_span, int32_t n, float *val, float x, void *stream) {
  if (n <= 0) return 0;
  fill_edge_vals_kernel<<<grid_for(n, WPB, 4096), LAYER_BLOCK, 0, ST(stream)>>>((const int2 *)row_span, n, val, x);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
extern "C" int shadow_edge_vals_dropedge(const int32_t *row_span, const int32_t *row_ord, int32_t n, float p, uint32_t seed,
                                         const uint32_t *step_dev, float *val, void *stream) {
  if (n <= 0 || p <= 0.f) return 0;
  // the number of draws, int(e*p), and the stream position live on the device so the launch can sit in a CUDA graph
  dropedge_kernel<<<64, 256, 0, ST(stream)>>>((const int2 *)row_span, row_ord, n, p, seed, step_dev, val);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
extern "C" int shadow_edge_vals_row_normalize(const int32_t *row_span, int32_t n, int32_t mode, float *val, void *stream) {
  if (n <= 0) return 0;
  row_normalize_kernel<<<grid_for(n, WPB, 4096), LAYER_BLOCK, 0, ST(stream)>>>((const int2 *)row_span, n, mode, val);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
extern "C" int shadow_edge_vals_sym_normalize(const int32_t *row_span, const int32_t *col, int32_t col_off, int32_t n, int32_t symmetric_survival,
                                              const float *mask, float *val, float *deg_scratch, void *stream) {
  if (n <= 0) return 0;
  if (symmetric_survival) {
    if (mask == val) FAIL(SHADOW_EINVAL, "sym_normalize with symmetric survival needs mask != val");
    sym_survive_kernel<<<grid_for(n, WPB, 4096), LAYER_BLOCK, 0, ST(stream)>>>((const int2 *)row_span, col, col_off, n, mask, val, deg_scratch);
  } else {
    // no dropedge: `mask` (all ones) is copied through; degree = row sum clipped at 1
    if (mask != val) FAIL(SHADOW_EINVAL, "sym_normalize without symmetric survival works in place (mask == val)");
    row_degree_kernel<<<grid_for(n, WPB, 4096), LAYER_BLOCK, 0, ST(stream)>>>((const int2 *)row_span, n, val, deg_scratch);
  }
  sym_scale_kernel<<<grid_for(n, WPB, 4096), LAYER_BLOCK, 0, ST(stream)>>>((const int2 *)row_span, col, col_off, n, deg_scratch, val);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

extern "C" int shadow_spmm_csr_fwd_f32(const int32_t *row_span, const int32_t *col, int32_t col_off, const float *val, const float *X, float *Y,
                                       int32_t n, int32_t F, float beta, void *stream) {
  if (n <= 0 || F <= 0) return 0;
  const bool vec = (F % 4 == 0) && (F <= 1024) && (((uintptr_t)X | (uintptr_t)Y) % 16 == 0);
  const int grid = grid_for(n, WPB, 8192);
  const int2 *rs = (const int2 *)row_span;
  if (!vec) spmm_fwd_scalar_kernel<<<grid, LAYER_BLOCK, 0, ST(stream)>>>(rs, col, col_off, val, X, Y, n, F, beta);
  else if (F <= 128) spmm_fwd_vec_kernel<1><<<grid, LAYER_BLOCK, 0, ST(stream)>>>(rs, col, col_off, val, X, Y, n, F, beta);
  else if (F <= 256) spmm_fwd_vec_kernel<2><<<grid, LAYER_BLOCK, 0, ST(stream)>>>(rs, col, col_off, val, X, Y, n, F, beta);
  else if (F <= 512) spmm_fwd_vec_kernel<4><<<grid, LAYER_BLOCK, 0, ST(stream)>>>(rs, col, col_off, val, X, Y, n, F, beta);
  else spmm_fwd_vec_kernel<8><<<grid, LAYER_BLOCK, 0, ST(stream)>>>(rs, col, col_off, val, X, Y, n, F, beta);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
extern "C" int shadow_spmm_csr_bwd_f32(const int32_t *row_span, const int32_t *col, int32_t col_off, const float *val, const float *dY, float *dX,
                                       int32_t n, int32_t F, void *stream) {
  if (n <= 0 || F <= 0) return 0;
  const bool vec = (F % 4 == 0) && (((uintptr_t)dY | (uintptr_t)dX) % 16 == 0);
  if (vec) spmm_bwd_kernel<true><<<grid_for(n, WPB, 8192), LAYER_BLOCK, 0, ST(stream)>>>((const int2 *)row_span, col, col_off, val, dY, dX, n, F);
  else spmm_bwd_kernel<false><<<grid_for(n, WPB, 8192), LAYER_BLOCK, 0, ST(stream)>>>((const int2 *)row_span, col, col_off, val, dY, dX, n, F);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
extern "C" int shadow_act_norm_fwd_f32(const float *Z, int32_t ldz, const float *scale, const float *offset, float *out, int32_t ldo,
                                       float *mean, float *rstd, int32_t n, int32_t D, int32_t act, int32_t do_norm, int32_t accumulate, void *stream) {
  if (n <= 0) return 0;
  if (D > 32 * NORM_MAX_PER_LANE) FAIL(SHADOW_EINVAL, "norm_feat: D=%d exceeds %d", D, 32 * NORM_MAX_PER_LANE);
#define LAUNCH_FWD(NPL) act_norm_fwd_kernel<NPL><<<grid_for(n, WPB, 8192), LAYER_BLOCK, 0, ST(stream)>>>(Z, ldz, scale, offset, out, ldo, mean, rstd, n, D, act, do_norm, accumulate)
  if (D <= 32) LAUNCH_FWD(1); else if (D <= 64) LAUNCH_FWD(2); else if (D <= 128) LAUNCH_FWD(4); else if (D <= 256) LAUNCH_FWD(8);
  else if (D <= 512) LAUNCH_FWD(16); else LAUNCH_FWD(32);
#undef LAUNCH_FWD
  CUDA_TRY(cudaGetLastError());
  return 0;
}
extern "C" int shadow_act_norm_bwd_f32(const float *dOut, int32_t ldo, const float *Z, int32_t ldz, const float *scale, const float *mean,
                                       const float *rstd, float *dZ, int32_t lddz, float *dscale, float *doffset, float *dbias, int32_t n, int32_t D,
                                       int32_t act, int32_t do_norm, void *stream) {
  if (n <= 0) return 0;
  if (D > 32 * NORM_MAX_PER_LANE) FAIL(SHADOW_EINVAL, "norm_feat: D=%d exceeds %d", D, 32 * NORM_MAX_PER_LANE);
  // every CTA ends with 3*D global atomics onto the same 3*D addresses; 148 / 296 / 592 / 1184 CTAs measured within noise of each other
  static const int cap = getenv("SHADOW_ANB_CAP") ? std::max(1, atoi(getenv("SHADOW_ANB_CAP"))) : 592;
  const size_t anb_smem = (size_t)WPB * 3 * D * sizeof(float);
#define LAUNCH_BWD(NPL)                                                                                                                       \
  do {                                                                                                                                        \
    if (anb_smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(act_norm_bwd_kernel<NPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)anb_smem)); \
    act_norm_bwd_kernel<NPL><<<grid_for(n, WPB, cap), LAYER_BLOCK, anb_smem, ST(stream)>>>(dOut, ldo, Z, ldz, scale, mean, rstd, dZ, lddz, dscale, doffset, dbias, n, D, act, do_norm); \
  } while (0)
  if (D <= 32) LAUNCH_BWD(1); else if (D <= 64) LAUNCH_BWD(2); else if (D <= 128) LAUNCH_BWD(4); else if (D <= 256) LAUNCH_BWD(8);
  else if (D <= 512) LAUNCH_BWD(16); else LAUNCH_BWD(32);
#undef LAUNCH_BWD
  CUDA_TRY(cudaGetLastError());
  return 0;
}
static int anb_pair_parts(int n) {
  int sms = 148;
  { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  // CTAs (= partial column sums) per SM: the sums are finished off the critical path (side stream), so what counts is the row loop's latency
  static const int mult = getenv("SHADOW_ANB_PAIR_MULT") ? std::min(8, std::max(1, atoi(getenv("SHADOW_ANB_PAIR_MULT")))) : 2;
  return grid_for(n, WPB, mult * sms);
}
extern "C" int32_t shadow_act_norm_bwd_pair_nparts(int32_t n) { return n <= 0 ? 0 : anb_pair_parts(n); }
static int act_norm_bwd_pair_impl(const float *dOut, int32_t ldo, const float *Z0, const float *Z1, int32_t ldz, const float *scale0,
                                  const float *scale1, const float *mean0, const float *rstd0, const float *mean1, const float *rstd1,
                                  float *dZ0, float *dZ1, int32_t lddz, float *dscale0, float *doffset0, float *dbias0, float *dscale1,
                                  float *doffset1, float *dbias1, int32_t n, int32_t D, int32_t act, int32_t do_norm, float *scratch,
                                  int64_t scratch_floats, void *stream, bool finish) {
  if (n <= 0) return 0;
  if (D > 256) FAIL(SHADOW_EINVAL, "act_norm_bwd_pair: D=%d exceeds 256", D);
  const int nb = Z1 ? 2 : 1;
  AnbPair A;
  A.Z[0] = Z0; A.Z[1] = Z1; A.scale[0] = scale0; A.scale[1] = scale1; A.mean[0] = mean0; A.mean[1] = mean1; A.rstd[0] = rstd0; A.rstd[1] = rstd1;
  A.dZ[0] = dZ0; A.dZ[1] = dZ1; A.dscale[0] = dscale0; A.dscale[1] = dscale1; A.doffset[0] = doffset0; A.doffset[1] = doffset1; A.dbias[0] = dbias0; A.dbias[1] = dbias1;
  const int grid = anb_pair_parts(n);
  if (!scratch || scratch_floats < (int64_t)grid * nb * 3 * D) FAIL(SHADOW_EINVAL, "act_norm_bwd_pair: scratch needs %lld floats", (long long)grid * nb * 3 * D);
  const size_t smem = (size_t)WPB * nb * 3 * D * sizeof(float);
  if ((D & 3) || (ldo & 3) || (ldz & 3) || (lddz & 3)) FAIL(SHADOW_EINVAL, "act_norm_bwd_pair: D and the leading dimensions must be multiples of 4");
#define LAUNCH_PAIR4(NV, NB, ACT, NORM)                                                                                                              \
  do {                                                                                                                                               \
    if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(act_norm_bwd_pair_kernel<NV, NB, ACT, NORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    act_norm_bwd_pair_kernel<NV, NB, ACT, NORM><<<grid, LAYER_BLOCK, smem, ST(stream)>>>(dOut, ldo, A, ldz, lddz, n, D, scratch);                    \
  } while (0)
#define LAUNCH_PAIR3(NV, NB, ACT) do { if (do_norm) LAUNCH_PAIR4(NV, NB, ACT, true); else LAUNCH_PAIR4(NV, NB, ACT, false); } while (0)
#define LAUNCH_PAIR2(NV, NB)                                                                                                                          \
  do {                                                                                                                                               \
    switch (act) {                                                                                                                                   \
      case ACT_RELU: LAUNCH_PAIR3(NV, NB, ACT_RELU); break;                                                                                          \
      case ACT_ELU: LAUNCH_PAIR3(NV, NB, ACT_ELU); break;                                                                                            \
      case ACT_TANH: LAUNCH_PAIR3(NV, NB, ACT_TANH); break;                                                                                          \
      case ACT_LRELU: LAUNCH_PAIR3(NV, NB, ACT_LRELU); break;                                                                                        \
      default: LAUNCH_PAIR3(NV, NB, ACT_I); break;                                                                                                   \
    }                                                                                                                                                \
  } while (0)
  if (nb == 2) { if (D <= 128) LAUNCH_PAIR2(1, 2); else LAUNCH_PAIR2(2, 2); }
  else { if (D <= 128) LAUNCH_PAIR2(1, 1); else LAUNCH_PAIR2(2, 1); }
#undef LAUNCH_PAIR2
#undef LAUNCH_PAIR3
#undef LAUNCH_PAIR4
  CUDA_TRY(cudaGetLastError());
  if (!finish) return 0;
  const int cols = nb * 3 * D;
  colsum_finish_kernel<<<(cols + 31) / 32, 256, 0, ST(stream)>>>(scratch, grid, D, nb, A, do_norm);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
extern "C" int shadow_act_norm_bwd_pair_f32(const float *dOut, int32_t ldo, const float *Z0, const float *Z1, int32_t ldz, const float *scale0,
                                            const float *scale1, const float *mean0, const float *rstd0, const float *mean1, const float *rstd1,
                                            float *dZ0, float *dZ1, int32_t lddz, float *dscale0, float *doffset0, float *dbias0, float *dscale1,
                                            float *doffset1, float *dbias1, int32_t n, int32_t D, int32_t act, int32_t do_norm, float *scratch,
                                            int64_t scratch_floats, void *stream) {
  return act_norm_bwd_pair_impl(dOut, ldo, Z0, Z1, ldz, scale0, scale1, mean0, rstd0, mean1, rstd1, dZ0, dZ1, lddz, dscale0, doffset0, dbias0, dscale1,
                                doffset1, dbias1, n, D, act, do_norm, scratch, scratch_floats, stream, true);
}
// the same without the final column-sum launch: the caller runs shadow_colsum_finish_f32 over `scratch` (shadow_act_norm_bwd_pair_nparts(n)
// parts; two branches with norm_feat only) on a stream of its own choice -- the parameter gradients are off the backward pass's critical path
extern "C" int shadow_act_norm_bwd_pair_nofinish_f32(const float *dOut, int32_t ldo, const float *Z0, const float *Z1, int32_t ldz, const float *scale0,
                                                     const float *scale1, const float *mean0, const float *rstd0, const float *mean1, const float *rstd1,
                                                     float *dZ0, float *dZ1, int32_t lddz, int32_t n, int32_t D, int32_t act, int32_t do_norm, float *scratch,
                                                     int64_t scratch_floats, void *stream) {
  if (!Z1 || !do_norm) FAIL(SHADOW_EINVAL, "act_norm_bwd_pair_nofinish: two branches with norm_feat only");
  return act_norm_bwd_pair_impl(dOut, ldo, Z0, Z1, ldz, scale0, scale1, mean0, rstd0, mean1, rstd1, dZ0, dZ1, lddz, nullptr, nullptr, nullptr, nullptr,
                                nullptr, nullptr, n, D, act, do_norm, scratch, scratch_floats, stream, false);
}
// [parts][2][3][D] partial column sums -> dst[b][plane] += total (fixed order; NULL destinations are skipped).  Used by csrc/gat.cu.
extern "C" int shadow_colsum_finish_f32(const float *partials, int32_t nparts, int32_t D, float *d00, float *d01, float *d02, float *d10, float *d11, float *d12,
                                        void *stream) {
  if (nparts <= 0 || D <= 0) return 0;
  AnbPair A;
  memset(&A, 0, sizeof(A));
  A.dscale[0] = d00; A.doffset[0] = d01; A.dbias[0] = d02; A.dscale[1] = d10; A.doffset[1] = d11; A.dbias[1] = d12;
  const int cols = 2 * 3 * D;
  colsum_finish_kernel<<<(cols + 31) / 32, 256, 0, ST(stream)>>>(partials, nparts, D, 2, A, 1);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
extern "C" int shadow_gat_fwd_f32(const int32_t *row_span, const int32_t *col, int32_t col_off, const float *val, const float *a_self,
                                  const float *a_neigh, const float *H, float *out, float *rowmax, float *denom, int32_t n, int32_t heads,
                                  int32_t d, void *stream) {
  if (n <= 0) return 0;
  if (d > 256) FAIL(SHADOW_EINVAL, "gat: head dim %d exceeds 256", d);
  gat_fwd_kernel<<<grid_for((long long)n * heads, WPB, 8192), LAYER_BLOCK, 0, ST(stream)>>>((const int2 *)row_span, col, col_off, val, a_self, a_neigh, H, out,
                                                                                             rowmax, denom, n, heads, d);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
extern "C" int shadow_gat_bwd_f32(const int32_t *row_span, const int32_t *col, int32_t col_off, const float *val, const float *a_self,
                                  const float *a_neigh, const float *H, const float *out, const float *rowmax, const float *denom,
                                  const float *dOut, float *dH, float *da_self, float *da_neigh, int32_t n, int32_t heads, int32_t d, void *stream) {
  if (n <= 0) return 0;
  if (d > 256) FAIL(SHADOW_EINVAL, "gat: head dim %d exceeds 256", d);
  gat_bwd_kernel<<<grid_for((long long)n * heads, WPB, 8192), LAYER_BLOCK, 0, ST(stream)>>>((const int2 *)row_span, col, col_off, val, a_self, a_neigh, H, out,
                                                                                             rowmax, denom, dOut, dH, da_self, da_neigh, n, heads, d);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
extern "C" int shadow_segment_pool_fwd_f32(const float *X, const int32_t *seg, int32_t seg_off, int32_t S, int32_t F, int32_t mode, float *out,
                                           int32_t *argmax, void *stream) {
  if (S <= 0) return 0;
  segment_pool_fwd_kernel<<<grid_for((long long)S * F, 256, 4096), 256, 0, ST(stream)>>>(X, seg, seg_off, S, F, mode, out, argmax);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
extern "C" int shadow_segment_pool_bwd_f32(const float *dOut, const int32_t *seg, int32_t seg_off, int32_t S, int32_t F, int32_t mode,
                                           const int32_t *argmax, float *dX, void *stream) {
  if (S <= 0) return 0;
  segment_pool_bwd_kernel<<<grid_for((long long)S * F, 256, 4096), 256, 0, ST(stream)>>>(dOut, seg, seg_off, S, F, mode, argmax, dX);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
extern "C" int shadow_adam_clip_step_f32(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, float grad_scale,
                                         float max_norm, float lr, float beta1, float beta2, float eps, int32_t *step_dev, float *sqnorm_scratch,
                                         void *stream) {
  if (n <= 0) return 0;
  CUDA_TRY(cudaMemsetAsync(sqnorm_scratch, 0, 4, ST(stream)));
  bump_step_kernel<<<1, 1, 0, ST(stream)>>>(step_dev);
  sqnorm_kernel<<<grid_for(n, 256, 592), 256, 0, ST(stream)>>>(grad, n, grad_scale, sqnorm_scratch);
  adam_clip_kernel<<<grid_for(n, 256, 1184), 256, 0, ST(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, sqnorm_scratch, grad_scale, max_norm, lr, beta1, beta2,
                                                                   eps, step_dev);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Data-parallel gradient exchange over NVLink peer memory, fused with the optimizer's norm pass (the step's one exchange).
// Every rank's flat gradient buffer lives in cudaMalloc'd memory opened by all peers through CUDA IPC.  One kernel per step and rank:
//   1. tell every peer "my gradients of step s are complete" (system-scope release store into the peer's flag array) and wait for the same
//      word from every peer (each CTA polls LOCAL memory on its own: no inter-CTA dependency, no co-residency requirement);
//   2. one-shot all-reduce: out[i] = sum over ranks r = 0..W-1 of grad_r[i], read straight from the peers' HBM (fixed order: every rank
//      forms bit-identical sums), accumulating the squared norm of the scaled sum for the clip on the way;
//   3. the last CTA tells every peer "I am done reading your buffer" -- the peer's next zero_grad waits for that word.
// No NCCL call, no staging copy, ~2.4 MB x (W-1) of NVLink reads per rank; the Adam kernel then runs on the reduced copy.
// A rank that never shows up does not hang the GPU: the polls give up after ~10 s and raise an error flag the host can read.
// ------------------------------------------------------------------------------------------------
#define P2P_MAX_WORLD 16
struct P2PArgs {
  const float *grads[P2P_MAX_WORLD];       // grads[r]: rank r's gradient buffer (own one included), device pointers valid on this GPU
  unsigned int *flags[P2P_MAX_WORLD];      // flags[r]: rank r's flag array uint32[2 * P2P_MAX_WORLD]: [q] = "rank q's gradients ready", [MAX + q] = "rank q done reading"
  int world, rank;
};
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ bool p2p_wait(const unsigned int *flag, unsigned int want, int *err) {
  if ((int)(ld_acquire_sys(flag) - want) >= 0) return true;
  const unsigned long long t0 = global_ns();
  while ((int)(ld_acquire_sys(flag) - want) < 0) {
    __nanosleep(100);
    if (global_ns() - t0 > 10000000000ull) { *err = 1; return false; }      // 10 s: a peer is gone; do not hang the GPU
  }
  return true;
}
__global__ void __launch_bounds__(256) p2p_reduce_sqnorm_kernel(const P2PArgs a, float *__restrict__ out, long long n, float scale, float *__restrict__ sqnorm,
                                                                unsigned int *step_ctr, unsigned int *done_ticket, int *err, float *partials) {
  const unsigned int step = *step_ctr + 1u;            // bumped by the last CTA below; every CTA reads it before that can happen (ticket)
  if (blockIdx.x == 0 && threadIdx.x < a.world) { __threadfence_system(); st_release_sys(a.flags[threadIdx.x] + a.rank, step); }
  if (threadIdx.x < a.world) p2p_wait(a.flags[a.rank] + threadIdx.x, step, err);
  __syncthreads();
  float s = 0.f;
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < a.world; r++) {
      const float4 v = __ldcv(reinterpret_cast<const float4 *>(a.grads[r]) + i);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    reinterpret_cast<float4 *>(out)[i] = acc;
    s += (acc.x * scale) * (acc.x * scale) + (acc.y * scale) * (acc.y * scale) + (acc.z * scale) * (acc.z * scale) + (acc.w * scale) * (acc.w * scale);
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int r = 0; r < a.world; r++) acc += __ldcv(a.grads[r] + i);
    out[i] = acc;
    s += (acc * scale) * (acc * scale);
  }
  s = warp_sum(s);
  __shared__ float sh[32];
  __shared__ unsigned int last;
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
    s = warp_sum(s);
    if (threadIdx.x == 0) {
      partials[blockIdx.x] = s;                        // per-CTA partial; summed below in CTA order: the same bits on every rank (same grid,
      __threadfence();                                 // same sums), so the clip factor -- and with it the replicas' weights -- cannot drift
      last = atomicAdd(done_ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
    }
  }
  __syncthreads();
  if (last) {                                          // every CTA of this rank has read all it needs
    if (threadIdx.x < a.world) st_release_sys(a.flags[threadIdx.x] + P2P_MAX_WORLD + a.rank, step);
    if (threadIdx.x == 0) {
      __threadfence();
      float tot = 0.f;
      for (unsigned int b = 0; b < gridDim.x; b++) tot += __ldcg(partials + b);
      *sqnorm = tot;
      *done_ticket = 0u; *step_ctr = step;
    }
  }
}
// zero_grad of a shared gradient buffer: wait until every peer has finished reading the previous step's gradients, then clear
__global__ void __launch_bounds__(256) p2p_wait_zero_kernel(const P2PArgs a, float *__restrict__ grad, long long n, const unsigned int *step_ctr, int *err) {
  const unsigned int step = *step_ctr;
  if (threadIdx.x < a.world) p2p_wait(a.flags[a.rank] + P2P_MAX_WORLD + threadIdx.x, step, err);
  __syncthreads();
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
    reinterpret_cast<float4 *>(grad)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) grad[i] = 0.f;
}

extern "C" int shadow_p2p_alloc(int64_t bytes, void **ptr, unsigned char *handle64) {
  if (!ptr || !handle64 || bytes <= 0) FAIL(SHADOW_EINVAL, "bad argument to shadow_p2p_alloc");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void *p = nullptr;
  CUDA_TRY(cudaMalloc(&p, (size_t)bytes));
  CUDA_TRY(cudaMemset(p, 0, (size_t)bytes));
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, p) != cudaSuccess) { cudaFree(p); FAIL(SHADOW_ECUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(cudaGetLastError())); }
  memcpy(handle64, &h, 64);
  *ptr = p;
  return 0;
}
extern "C" int shadow_p2p_open(const unsigned char *handle64, void **ptr) {
  if (!ptr || !handle64) FAIL(SHADOW_EINVAL, "bad argument to shadow_p2p_open");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  CUDA_TRY(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}
extern "C" int shadow_p2p_close(void *ptr) { if (ptr) CUDA_TRY(cudaIpcCloseMemHandle(ptr)); return 0; }
extern "C" int shadow_p2p_free(void *ptr) { if (ptr) CUDA_TRY(cudaFree(ptr)); return 0; }

static int p2p_args(P2PArgs *a, const uint64_t *grad_ptrs, const uint64_t *flag_ptrs, int world, int rank) {
  if (!grad_ptrs || !flag_ptrs || world < 1 || world > P2P_MAX_WORLD || rank < 0 || rank >= world) FAIL(SHADOW_EINVAL, "p2p: world must be in [1,%d]", P2P_MAX_WORLD);
  memset(a, 0, sizeof(*a));
  for (int r = 0; r < world; r++) { a->grads[r] = (const float *)(uintptr_t)grad_ptrs[r]; a->flags[r] = (unsigned int *)(uintptr_t)flag_ptrs[r]; }
  a->world = world; a->rank = rank;
  return 0;
}
// state: uint32[4 + 512] in LOCAL device memory = {step counter, done ticket, error flag, -, per-CTA partial squared norms (float)}
extern "C" int shadow_p2p_zero_grad_f32(const uint64_t *grad_ptrs, const uint64_t *flag_ptrs, int32_t world, int32_t rank, int64_t n, uint32_t *state, void *stream) {
  P2PArgs a;
  int rc = p2p_args(&a, grad_ptrs, flag_ptrs, world, rank); if (rc) return rc;
  p2p_wait_zero_kernel<<<grid_for(n / 4 + 1, 256, 296), 256, 0, ST(stream)>>>(a, (float *)(uintptr_t)grad_ptrs[rank], n, state, (int *)(state + 2));
  CUDA_TRY(cudaGetLastError());
  return 0;
}
// the optimizer step of shadow_adam_clip_step_f32 on the SUM of all ranks' gradients (grad_scale = 1 / world for the mean): p2p_reduce_sqnorm_kernel
// writes the sum to `gsum` (local, n floats), adam_clip_kernel consumes it
extern "C" int shadow_p2p_adam_clip_step_f32(const uint64_t *grad_ptrs, const uint64_t *flag_ptrs, int32_t world, int32_t rank, float *param, float *gsum,
                                             float *exp_avg, float *exp_avg_sq, int64_t n, float grad_scale, float max_norm, float lr, float beta1, float beta2,
                                             float eps, int32_t *step_dev, float *sqnorm_scratch, uint32_t *state, void *stream) {
  if (n <= 0) return 0;
  P2PArgs a;
  int rc = p2p_args(&a, grad_ptrs, flag_ptrs, world, rank); if (rc) return rc;
  CUDA_TRY(cudaMemsetAsync(sqnorm_scratch, 0, 4, ST(stream)));
  bump_step_kernel<<<1, 1, 0, ST(stream)>>>(step_dev);
  p2p_reduce_sqnorm_kernel<<<grid_for(n / 4 + 1, 256, 296), 256, 0, ST(stream)>>>(a, gsum, n, grad_scale, sqnorm_scratch, state, state + 1, (int *)(state + 2),
                                                                                   (float *)(state + 4));
  adam_clip_kernel<<<grid_for(n, 256, 1184), 256, 0, ST(stream)>>>(param, gsum, exp_avg, exp_avg_sq, n, sqnorm_scratch, grad_scale, max_norm, lr, beta1, beta2,
                                                                   eps, step_dev);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
