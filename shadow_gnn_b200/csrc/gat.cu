// gat.cu -- the GAT layer of shaDow (shaDow/layers.py:539-645) between its two Linear products, as six row-parallel kernels.
//
//   h_self = act(X W0^T + b0), h_neigh = act(X W1^T + b1)                      csrc/linear_tc.cu (one launch, both branches)
//   s_b[i,k] = sum_f att[b,k,f] h_b[i,k,f];  a_b = LeakyReLU_0.2(s_b)           gat_logits_fwd_kernel          (layers.py:567-568)
//   agg[i,k,:] = sum_j u_ij h_neigh[j,k,:] / clamp(sum_j u_ij, 1e-10),
//        u_ij = exp(a_self[i,k] + a_neigh[j,k] - max_j(..)) * A_ij             gat_agg_fwd_kernel             (layers.py:569-582)
//   out = (norm_feat_{1,k}(h_self[:,k]) + norm_feat_{0,k}(agg[:,k])) / 2       gat_headnorm_fwd_kernel        (layers.py:621-627,329-338)
// and the three backward kernels in reverse order (gat_headnorm_bwd, gat_agg_bwd, gat_pre_bwd = attention-logit backward + activation
// backward + the column sums for d att and d bias).
//
// Mapping: ONE WARP PER ROW, all heads at once: a lane owns the float4 "slots" lane and lane + 32 of the row (D = heads * d <= 256 features),
// a head is a run of d / 4 consecutive slots, so per-head reductions are xor-butterflies inside lane segments.  The aggregation walks a row's
// edges 32 at a time: lane e prepares edge e (neighbour id, softmax weight of every head), the warp then takes the 32 neighbours one after the
// other with their feature rows loaded as two coalesced 128-bit loads per lane, four edges in flight.  Column sums (d scale, d offset, d att,
// d bias) leave each CTA as partial sums and are added up in CTA order by colsum_finish_kernel (csrc/layers.cu): no atomics there; the
// scatter into d h_neigh uses vector reductions (red.global.add.v4.f32) like the SpMM backward.
// Supported shapes: d in {16, 32, 64, 128, 256}, heads <= 8, D <= 256; the Python layer keeps its composed path for anything else.
#include <algorithm>

#include "common.cuh"

#define GAT_BLOCK 256
#define GAT_MAXH 8

namespace {

enum { ACT_RELU = 0, ACT_I = 1, ACT_ELU = 2, ACT_TANH = 3, ACT_LRELU = 4 };
__device__ __forceinline__ float act_grad_from_output(float a, int act) {      // act'(z) written in terms of a = act(z)
  switch (act) {
    case ACT_RELU: return a > 0.f ? 1.f : 0.f;
    case ACT_ELU: return a > 0.f ? 1.f : a + 1.f;
    case ACT_TANH: return 1.f - a * a;
    case ACT_LRELU: return a > 0.f ? 1.f : 0.2f;
    default: return 1.f;
  }
}
__device__ __forceinline__ float warp_sum_all(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
struct RowMap {            // where this lane's two slots sit
  int lph;                 // slots (lanes) per head: d / 4
  int head[2];             // head of slot v (slot index = lane + 32 v)
  bool on[2];              // slot exists (4 * slot < D)
};
__device__ __forceinline__ RowMap row_map(int lane, int D, int d) {
  RowMap m;
  m.lph = d >> 2;
#pragma unroll
  for (int v = 0; v < 2; v++) { const int s = lane + 32 * v; m.on[v] = 4 * s < D; m.head[v] = m.on[v] ? (4 * s) / d : 0; }
  return m;
}
// per-head sums of x[v] (one value per slot): afterwards every lane holds the sum of ITS head in x[v]
__device__ __forceinline__ void head_sum(float (&x)[2], const RowMap &m) {
  if (m.lph >= 64) {                                   // one head spans both slots
    const float t = warp_sum_all(x[0] + x[1]);
    x[0] = t; x[1] = t;
  } else {
    for (int o = m.lph >> 1; o > 0; o >>= 1) { x[0] += __shfl_xor_sync(0xffffffffu, x[0], o); x[1] += __shfl_xor_sync(0xffffffffu, x[1], o); }
  }
}
__device__ __forceinline__ float4 ld4(const float *p, bool on) { return on ? *reinterpret_cast<const float4 *>(p) : make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float dot4(const float4 a, const float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
__device__ __forceinline__ float sum4(const float4 a) { return a.x + a.y + a.z + a.w; }

// ---------------- attention logits ----------------
__global__ void __launch_bounds__(GAT_BLOCK) gat_logits_fwd_kernel(const float *__restrict__ h_self, const float *__restrict__ h_neigh, const float *__restrict__ att,
                                                                   float *__restrict__ s_self, float *__restrict__ s_neigh, float *__restrict__ a_self,
                                                                   float *__restrict__ a_neigh, int n, int heads, int d) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5, D = heads * d;
  const RowMap m = row_map(lane, D, d);
  float4 at[2][2];
#pragma unroll
  for (int b = 0; b < 2; b++)
#pragma unroll
    for (int v = 0; v < 2; v++) at[b][v] = ld4(att + (size_t)b * D + 4 * (lane + 32 * v), m.on[v]);
  for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < n; i += gridDim.x * wpb) {
#pragma unroll
    for (int b = 0; b < 2; b++) {
      const float *h = (b ? h_neigh : h_self) + (size_t)i * D;
      float x[2];
#pragma unroll
      for (int v = 0; v < 2; v++) x[v] = dot4(ld4(h + 4 * (lane + 32 * v), m.on[v]), at[b][v]);
      head_sum(x, m);
#pragma unroll
      for (int v = 0; v < 2; v++) {
        const int s = lane + 32 * v;
        if (m.on[v] && (4 * s) % d == 0 && !(m.lph >= 64 && v == 1)) {
          const float sv = x[v];
          (b ? s_neigh : s_self)[(size_t)i * heads + m.head[v]] = sv;
          (b ? a_neigh : a_self)[(size_t)i * heads + m.head[v]] = sv > 0.f ? sv : 0.2f * sv;
        }
      }
    }
  }
}

// ---------------- attention aggregation, forward ----------------
// Rows of up to 32 edges are handled by one warp.  A longer row (the root's ~140 in-scope neighbours) is queued and then taken by ALL warps
// of the CTA: every warp finds the logit maxima of its 32-edge pieces, the maxima are combined through shared memory, every warp
// accumulates its pieces, and the partial sums are added in warp order (deterministic).
template <int NH>
__global__ void __launch_bounds__(GAT_BLOCK) gat_agg_fwd_kernel(const int2 *__restrict__ row_span, const int *__restrict__ col, int col_off, const float *__restrict__ val,
                                                                const float *__restrict__ a_self, const float *__restrict__ a_neigh, const float *__restrict__ Hn,
                                                                float *__restrict__ out, float *__restrict__ rowmax, float *__restrict__ denom, int n, int d) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5, warp = threadIdx.x >> 5, D = NH * d;
  const RowMap m = row_map(lane, D, d);
  __shared__ int long_rows[GAT_BLOCK / 32];
  __shared__ int n_long;
  __shared__ float sh_mx[GAT_BLOCK / 32][GAT_MAXH], sh_den[GAT_BLOCK / 32][GAT_MAXH];
  __shared__ float4 sh_acc[GAT_BLOCK / 32][64];

  // logit maxima of edges [p_lo, p_hi) (any length) for every head, warp-uniform
  auto piece_max = [&](const float (&as)[NH], const int p_lo, const int p_hi, float (&mx)[NH]) {
#pragma unroll
    for (int k = 0; k < NH; k++) mx[k] = -INFINITY;
    for (int p = p_lo + lane; p < p_hi; p += 32) {
      const int j = col[p] - col_off;
#pragma unroll
      for (int k = 0; k < NH; k++) mx[k] = fmaxf(mx[k], as[k] + a_neigh[(size_t)j * NH + k]);
    }
#pragma unroll
    for (int k = 0; k < NH; k++)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
  };
  // acc += sum over edges [p_lo, p_hi) (at most 32) of u_ij h_j; den += sum of u_ij (lane-local partials)
  auto piece_acc = [&](const float (&as)[NH], const float (&mx)[NH], const int p_lo, const int p_hi, float4 (&acc)[2], float (&den)[NH]) {
    const int pl = p_lo + lane;
    int j_l = 0;
    float u_l[NH];
#pragma unroll
    for (int k = 0; k < NH; k++) u_l[k] = 0.f;
    if (pl < p_hi) {
      j_l = col[pl] - col_off;
      const float w = val ? val[pl] : 1.f;
#pragma unroll
      for (int k = 0; k < NH; k++) { u_l[k] = expf(as[k] + a_neigh[(size_t)j_l * NH + k] - mx[k]) * w; den[k] += u_l[k]; }
    }
    const int cnt = p_hi - p_lo;
#pragma unroll 4
    for (int t = 0; t < cnt; t++) {
      const int j = __shfl_sync(0xffffffffu, j_l, t);
      float u[2] = {0.f, 0.f};
#pragma unroll
      for (int k = 0; k < NH; k++) {
        const float uk = __shfl_sync(0xffffffffu, u_l[k], t);
        if (m.head[0] == k) u[0] = uk;
        if (m.head[1] == k) u[1] = uk;
      }
#pragma unroll
      for (int v = 0; v < 2; v++) {
        const float4 h = ld4(Hn + (size_t)j * D + 4 * (lane + 32 * v), m.on[v]);
        acc[v].x += u[v] * h.x; acc[v].y += u[v] * h.y; acc[v].z += u[v] * h.z; acc[v].w += u[v] * h.w;
      }
    }
  };
  auto finish = [&](const int i, const float (&mx)[NH], const float4 (&acc)[2], const float (&den)[NH]) {      // den: warp-uniform totals
#pragma unroll
    for (int v = 0; v < 2; v++) {
      if (m.on[v]) {
        float S = 1.f;
#pragma unroll
        for (int k = 0; k < NH; k++) if (m.head[v] == k) S = fmaxf(den[k], 1e-10f);
        *reinterpret_cast<float4 *>(out + (size_t)i * D + 4 * (lane + 32 * v)) = make_float4(acc[v].x / S, acc[v].y / S, acc[v].z / S, acc[v].w / S);
      }
    }
    if (lane < NH) {
      float mk = 0.f, dk = 0.f;
#pragma unroll
      for (int k = 0; k < NH; k++) if (lane == k) { mk = mx[k]; dk = den[k]; }
      rowmax[(size_t)i * NH + lane] = mk; denom[(size_t)i * NH + lane] = dk;
    }
  };

  for (int base = blockIdx.x * wpb; base < n; base += gridDim.x * wpb) {
    if (threadIdx.x == 0) n_long = 0;
    __syncthreads();
    const int i = base + warp;
    if (i < n) {
      const int2 sp = row_span[i];
      if (sp.y - sp.x > 32) { if (lane == 0) long_rows[atomicAdd(&n_long, 1)] = i; }
      else {
        float as[NH], mx[NH], den[NH];
#pragma unroll
        for (int k = 0; k < NH; k++) { as[k] = a_self[(size_t)i * NH + k]; den[k] = 0.f; }
        piece_max(as, sp.x, sp.y, mx);
        float4 acc[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
        if (sp.y > sp.x) piece_acc(as, mx, sp.x, sp.y, acc, den);
#pragma unroll
        for (int k = 0; k < NH; k++) den[k] = warp_sum_all(den[k]);
        finish(i, mx, acc, den);
      }
    }
    __syncthreads();
    const int nl = n_long;
    for (int r = 0; r < nl; r++) {
      const int il = long_rows[r];
      const int2 sp = row_span[il];
      float as[NH], mx[NH], den[NH];
#pragma unroll
      for (int k = 0; k < NH; k++) { as[k] = a_self[(size_t)il * NH + k]; den[k] = 0.f; }
      // this warp's edges: pieces warp, warp + wpb, ... of 32 edges
      float pm[NH];
#pragma unroll
      for (int k = 0; k < NH; k++) pm[k] = -INFINITY;
      for (int p = sp.x + 32 * warp; p < sp.y; p += 32 * wpb) {
        float t[NH];
        piece_max(as, p, min(p + 32, sp.y), t);
#pragma unroll
        for (int k = 0; k < NH; k++) pm[k] = fmaxf(pm[k], t[k]);
      }
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NH; k++) sh_mx[warp][k] = pm[k];
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < NH; k++) {
        mx[k] = -INFINITY;
        for (int w = 0; w < wpb; w++) mx[k] = fmaxf(mx[k], sh_mx[w][k]);
      }
      float4 acc[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
      for (int p = sp.x + 32 * warp; p < sp.y; p += 32 * wpb) piece_acc(as, mx, p, min(p + 32, sp.y), acc, den);
#pragma unroll
      for (int k = 0; k < NH; k++) den[k] = warp_sum_all(den[k]);
      sh_acc[warp][lane] = acc[0]; sh_acc[warp][lane + 32] = acc[1];
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NH; k++) sh_den[warp][k] = den[k];
      }
      __syncthreads();
      if (warp == 0) {
        float4 tot[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
#pragma unroll
        for (int k = 0; k < NH; k++) den[k] = 0.f;
        for (int w = 0; w < wpb; w++) {
#pragma unroll
          for (int v = 0; v < 2; v++) { const float4 a = sh_acc[w][lane + 32 * v]; tot[v].x += a.x; tot[v].y += a.y; tot[v].z += a.z; tot[v].w += a.w; }
#pragma unroll
          for (int k = 0; k < NH; k++) den[k] += sh_den[w][k];
        }
        finish(il, mx, tot, den);
      }
      __syncthreads();
    }
    __syncthreads();
  }
}

// ---------------- attention aggregation, backward: d agg -> d h_neigh (scatter), d a_self, d a_neigh (scatter) ----------------
// Every output is a scatter / sum, so a row can be cut freely: rows of up to 32 edges are walked by their warp, longer ones (the root's ~140
// in-scope neighbours) are queued and all warps of the CTA take 32-edge pieces of them.  da_self is accumulated with atomics (pre-zeroed).
template <int NH>
__global__ void __launch_bounds__(GAT_BLOCK) gat_agg_bwd_kernel(const int2 *__restrict__ row_span, const int *__restrict__ col, int col_off, const float *__restrict__ val,
                                                                const float *__restrict__ a_self, const float *__restrict__ a_neigh, const float *__restrict__ Hn,
                                                                const float *__restrict__ out, const float *__restrict__ rowmax, const float *__restrict__ denom,
                                                                const float *__restrict__ dOut, float *__restrict__ dHn, float *__restrict__ da_self,
                                                                float *__restrict__ da_neigh, int n, int d) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5, warp = threadIdx.x >> 5, D = NH * d;
  const RowMap m = row_map(lane, D, d);
  __shared__ int long_rows[GAT_BLOCK / 32];
  __shared__ int n_long;
  const bool writer[2] = {m.on[0] && (4 * lane) % d == 0, m.on[1] && (4 * (lane + 32)) % d == 0 && m.lph < 64};      // lanes that own a head's first slot

  auto walk = [&](const int i, const int p_lo, const int p_hi) {        // edges [p_lo, p_hi) of row i, at most 32
    float as[NH], mx[NH], S[NH];
    bool clamped[NH];
#pragma unroll
    for (int k = 0; k < NH; k++) {
      as[k] = a_self[(size_t)i * NH + k]; mx[k] = rowmax[(size_t)i * NH + k];
      const float dn = denom[(size_t)i * NH + k];
      clamped[k] = dn < 1e-10f; S[k] = fmaxf(dn, 1e-10f);
    }
    float4 g[2];
    float go[2];
#pragma unroll
    for (int v = 0; v < 2; v++) {
      g[v] = ld4(dOut + (size_t)i * D + 4 * (lane + 32 * v), m.on[v]);
      go[v] = dot4(g[v], ld4(out + (size_t)i * D + 4 * (lane + 32 * v), m.on[v]));
    }
    head_sum(go, m);
#pragma unroll
    for (int v = 0; v < 2; v++)
#pragma unroll
      for (int k = 0; k < NH; k++) if (m.head[v] == k && clamped[k]) go[v] = 0.f;
    const int pl = p_lo + lane;
    int j_l = 0;
    float al_l[NH];                                     // alpha_ij of every head for my edge
#pragma unroll
    for (int k = 0; k < NH; k++) al_l[k] = 0.f;
    if (pl < p_hi) {
      j_l = col[pl] - col_off;
      const float w = val ? val[pl] : 1.f;
#pragma unroll
      for (int k = 0; k < NH; k++) al_l[k] = expf(as[k] + a_neigh[(size_t)j_l * NH + k] - mx[k]) * w / S[k];
    }
    const int cnt = p_hi - p_lo;
    float das[2] = {0.f, 0.f};
    for (int t0 = 0; t0 < cnt; t0 += 2) {               // two edges per round: their loads and butterflies overlap
      int j[2];
      float al[2][2], gh[2][2];
      float4 hrow[2][2];
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int t = min(t0 + e, cnt - 1);
        j[e] = __shfl_sync(0xffffffffu, j_l, t);
        al[e][0] = 0.f; al[e][1] = 0.f;
#pragma unroll
        for (int k = 0; k < NH; k++) {
          const float x = __shfl_sync(0xffffffffu, al_l[k], t);
          if (m.head[0] == k) al[e][0] = x;
          if (m.head[1] == k) al[e][1] = x;
        }
        if (t0 + e >= cnt) { al[e][0] = 0.f; al[e][1] = 0.f; }
#pragma unroll
        for (int v = 0; v < 2; v++) hrow[e][v] = ld4(Hn + (size_t)j[e] * D + 4 * (lane + 32 * v), m.on[v]);
      }
#pragma unroll
      for (int e = 0; e < 2; e++)
#pragma unroll
        for (int v = 0; v < 2; v++) {
          gh[e][v] = dot4(g[v], hrow[e][v]);
          if (m.on[v] && al[e][v] != 0.f)
            atomicAdd(reinterpret_cast<float4 *>(dHn + (size_t)j[e] * D) + lane + 32 * v,
                      make_float4(al[e][v] * g[v].x, al[e][v] * g[v].y, al[e][v] * g[v].z, al[e][v] * g[v].w));
        }
      if (m.lph >= 64) {
#pragma unroll
        for (int e = 0; e < 2; e++) { const float x = warp_sum_all(gh[e][0] + gh[e][1]); gh[e][0] = x; gh[e][1] = x; }
      } else {
        for (int o = m.lph >> 1; o > 0; o >>= 1) {
#pragma unroll
          for (int e = 0; e < 2; e++) { gh[e][0] += __shfl_xor_sync(0xffffffffu, gh[e][0], o); gh[e][1] += __shfl_xor_sync(0xffffffffu, gh[e][1], o); }
        }
      }
      // de_k = alpha_k (g.h_j - g.out_i) (the row max is shift-invariant: no gradient through it); one lane per head writes
#pragma unroll
      for (int e = 0; e < 2; e++)
#pragma unroll
        for (int v = 0; v < 2; v++) {
          if (writer[v] && al[e][v] != 0.f) {
            const float de = al[e][v] * (gh[e][v] - go[v]);
            atomicAdd(da_neigh + (size_t)j[e] * NH + m.head[v], de);
            das[v] += de;
          }
        }
    }
#pragma unroll
    for (int v = 0; v < 2; v++) if (writer[v] && das[v] != 0.f) atomicAdd(da_self + (size_t)i * NH + m.head[v], das[v]);
  };

  for (int base = blockIdx.x * wpb; base < n; base += gridDim.x * wpb) {
    if (threadIdx.x == 0) n_long = 0;
    __syncthreads();
    const int i = base + warp;
    if (i < n) {
      const int2 sp = row_span[i];
      if (sp.y - sp.x > 32) { if (lane == 0) long_rows[atomicAdd(&n_long, 1)] = i; }
      else if (sp.y > sp.x) walk(i, sp.x, sp.y);
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
}

// ---------------- per-head norm_feat of both branches, averaged ----------------
// scale / offset: [2][D] (index 0 = aggregated neighbours, 1 = self: the order of f_norm's argument list, layers.py:621-623)
__global__ void __launch_bounds__(GAT_BLOCK) gat_headnorm_fwd_kernel(const float *__restrict__ h_self, const float *__restrict__ agg, const float *__restrict__ scale,
                                                                     const float *__restrict__ offset, float *__restrict__ out, float *__restrict__ mean,
                                                                     float *__restrict__ rstd, int n, int heads, int d, int do_norm) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5, D = heads * d;
  const RowMap m = row_map(lane, D, d);
  float4 sc[2][2], of[2][2];
#pragma unroll
  for (int b = 0; b < 2; b++)
#pragma unroll
    for (int v = 0; v < 2; v++) {
      sc[b][v] = do_norm ? ld4(scale + (size_t)b * D + 4 * (lane + 32 * v), m.on[v]) : make_float4(1.f, 1.f, 1.f, 1.f);
      of[b][v] = do_norm ? ld4(offset + (size_t)b * D + 4 * (lane + 32 * v), m.on[v]) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < n; i += gridDim.x * wpb) {
    float4 o[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
#pragma unroll
    for (int b = 0; b < 2; b++) {                       // b = 0: agg, b = 1: self
      const float *src = (b ? h_self : agg) + (size_t)i * D;
      float4 x[2];
      float s1[2];
#pragma unroll
      for (int v = 0; v < 2; v++) { x[v] = ld4(src + 4 * (lane + 32 * v), m.on[v]); s1[v] = sum4(x[v]); }
      float mu[2] = {0.f, 0.f}, rs[2] = {1.f, 1.f};
      if (do_norm) {
        head_sum(s1, m);
        float s2[2];
#pragma unroll
        for (int v = 0; v < 2; v++) {
          mu[v] = s1[v] / (float)d;
          const float4 c = make_float4(x[v].x - mu[v], x[v].y - mu[v], x[v].z - mu[v], x[v].w - mu[v]);
          s2[v] = m.on[v] ? dot4(c, c) : 0.f;
        }
        head_sum(s2, m);
#pragma unroll
        for (int v = 0; v < 2; v++) {
          rs[v] = rsqrtf(s2[v] / (float)d + 1e-9f);
          const int s = lane + 32 * v;
          if (m.on[v] && (4 * s) % d == 0 && !(m.lph >= 64 && v == 1)) {
            mean[((size_t)b * n + i) * heads + m.head[v]] = mu[v];
            rstd[((size_t)b * n + i) * heads + m.head[v]] = rs[v];
          }
        }
      }
#pragma unroll
      for (int v = 0; v < 2; v++) {
        o[v].x += (x[v].x - mu[v]) * rs[v] * sc[b][v].x + of[b][v].x; o[v].y += (x[v].y - mu[v]) * rs[v] * sc[b][v].y + of[b][v].y;
        o[v].z += (x[v].z - mu[v]) * rs[v] * sc[b][v].z + of[b][v].z; o[v].w += (x[v].w - mu[v]) * rs[v] * sc[b][v].w + of[b][v].w;
      }
    }
#pragma unroll
    for (int v = 0; v < 2; v++)
      if (m.on[v]) *reinterpret_cast<float4 *>(out + (size_t)i * D + 4 * (lane + 32 * v)) = make_float4(0.5f * o[v].x, 0.5f * o[v].y, 0.5f * o[v].z, 0.5f * o[v].w);
  }
}

// backward: d h_self, d agg; partial column sums [cta][2 branches][3 planes: d scale, d offset, unused][D] (summed by colsum_finish_kernel)
__global__ void __launch_bounds__(GAT_BLOCK) gat_headnorm_bwd_kernel(const float *__restrict__ dOut, const float *__restrict__ h_self, const float *__restrict__ agg,
                                                                     const float *__restrict__ scale, const float *__restrict__ mean, const float *__restrict__ rstd,
                                                                     float *__restrict__ dh_self, float *__restrict__ dagg, float *__restrict__ partials, int n, int heads,
                                                                     int d, int do_norm) {
  extern __shared__ float sh[];               // [warps][2][2][D]
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5, warp = threadIdx.x >> 5, D = heads * d;
  const RowMap m = row_map(lane, D, d);
  float4 sc[2][2], ps[2][2], po[2][2];
#pragma unroll
  for (int b = 0; b < 2; b++)
#pragma unroll
    for (int v = 0; v < 2; v++) {
      sc[b][v] = do_norm ? ld4(scale + (size_t)b * D + 4 * (lane + 32 * v), m.on[v]) : make_float4(1.f, 1.f, 1.f, 1.f);
      ps[b][v] = make_float4(0.f, 0.f, 0.f, 0.f); po[b][v] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  for (int i = blockIdx.x * wpb + warp; i < n; i += gridDim.x * wpb) {
    float4 g[2];
#pragma unroll
    for (int v = 0; v < 2; v++) { g[v] = ld4(dOut + (size_t)i * D + 4 * (lane + 32 * v), m.on[v]); g[v].x *= 0.5f; g[v].y *= 0.5f; g[v].z *= 0.5f; g[v].w *= 0.5f; }
#pragma unroll
    for (int b = 0; b < 2; b++) {
      float *dst = (b ? dh_self : dagg) + (size_t)i * D;
      if (!do_norm) {
#pragma unroll
        for (int v = 0; v < 2; v++) if (m.on[v]) *reinterpret_cast<float4 *>(dst + 4 * (lane + 32 * v)) = g[v];
        continue;
      }
      const float *src = (b ? h_self : agg) + (size_t)i * D;
      float4 xh[2], dxh[2];
      float s1[2], s2[2];
#pragma unroll
      for (int v = 0; v < 2; v++) {
        const float mu = mean[((size_t)b * n + i) * heads + m.head[v]], rs = rstd[((size_t)b * n + i) * heads + m.head[v]];
        const float4 x = ld4(src + 4 * (lane + 32 * v), m.on[v]);
        xh[v] = m.on[v] ? make_float4((x.x - mu) * rs, (x.y - mu) * rs, (x.z - mu) * rs, (x.w - mu) * rs) : make_float4(0.f, 0.f, 0.f, 0.f);
        dxh[v] = make_float4(g[v].x * sc[b][v].x, g[v].y * sc[b][v].y, g[v].z * sc[b][v].z, g[v].w * sc[b][v].w);
        s1[v] = sum4(dxh[v]); s2[v] = dot4(dxh[v], xh[v]);
        ps[b][v].x += g[v].x * xh[v].x; ps[b][v].y += g[v].y * xh[v].y; ps[b][v].z += g[v].z * xh[v].z; ps[b][v].w += g[v].w * xh[v].w;
        po[b][v].x += g[v].x; po[b][v].y += g[v].y; po[b][v].z += g[v].z; po[b][v].w += g[v].w;
      }
      head_sum(s1, m); head_sum(s2, m);
#pragma unroll
      for (int v = 0; v < 2; v++) {
        if (m.on[v]) {
          const float rs = rstd[((size_t)b * n + i) * heads + m.head[v]], a = s1[v] / (float)d, c = s2[v] / (float)d;
          *reinterpret_cast<float4 *>(dst + 4 * (lane + 32 * v)) =
              make_float4(rs * (dxh[v].x - a - xh[v].x * c), rs * (dxh[v].y - a - xh[v].y * c), rs * (dxh[v].z - a - xh[v].z * c), rs * (dxh[v].w - a - xh[v].w * c));
        }
      }
    }
  }
  float *mine = sh + (size_t)warp * 4 * D;
#pragma unroll
  for (int b = 0; b < 2; b++)
#pragma unroll
    for (int v = 0; v < 2; v++)
      if (m.on[v]) {
        reinterpret_cast<float4 *>(mine + (b * 2 + 0) * D)[lane + 32 * v] = ps[b][v];
        reinterpret_cast<float4 *>(mine + (b * 2 + 1) * D)[lane + 32 * v] = po[b][v];
      }
  __syncthreads();
  for (int f = threadIdx.x; f < 4 * D; f += blockDim.x) {
    float v = 0.f;
    for (int w = 0; w < wpb; w++) v += sh[(size_t)w * 4 * D + f];
    const int b = f / (2 * D), plane = (f / D) & 1, c = f % D;
    partials[(size_t)blockIdx.x * 6 * D + (b * 3 + plane) * D + c] = v;
  }
  for (int f = threadIdx.x; f < 2 * D; f += blockDim.x) partials[(size_t)blockIdx.x * 6 * D + ((f / D) * 3 + 2) * D + f % D] = 0.f;
}

// ---------------- logits backward + activation backward, both branches ----------------
//   ds_b[i,k] = da_b[i,k] * LeakyReLU'(s_b[i,k]);  dh_b += ds_b[i,k] att[b,k,:];  dZ_b = dh_b * act'(z_b)  (from a = h_b)
//   partial column sums [cta][2 branches][3 planes: d att, unused, d bias][D]
__global__ void __launch_bounds__(GAT_BLOCK) gat_pre_bwd_kernel(const float *__restrict__ dh_self, const float *__restrict__ dh_neigh, const float *__restrict__ h_self,
                                                                const float *__restrict__ h_neigh, const float *__restrict__ s_self, const float *__restrict__ s_neigh,
                                                                const float *__restrict__ da_self, const float *__restrict__ da_neigh, const float *__restrict__ att,
                                                                float *__restrict__ dZ_self, float *__restrict__ dZ_neigh, float *__restrict__ partials, int n, int heads,
                                                                int d, int act) {
  extern __shared__ float sh[];               // [warps][2][2][D]
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5, warp = threadIdx.x >> 5, D = heads * d;
  const RowMap m = row_map(lane, D, d);
  float4 at[2][2], pa[2][2], pb[2][2];
#pragma unroll
  for (int b = 0; b < 2; b++)
#pragma unroll
    for (int v = 0; v < 2; v++) {
      at[b][v] = ld4(att + (size_t)b * D + 4 * (lane + 32 * v), m.on[v]);
      pa[b][v] = make_float4(0.f, 0.f, 0.f, 0.f); pb[b][v] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  for (int i = blockIdx.x * wpb + warp; i < n; i += gridDim.x * wpb) {
#pragma unroll
    for (int b = 0; b < 2; b++) {                       // b = 0: self, b = 1: neigh (the order of `attention` and f_lin, layers.py:553-556)
      const float *h = (b ? h_neigh : h_self) + (size_t)i * D, *dh = (b ? dh_neigh : dh_self) + (size_t)i * D;
      const float *sv = (b ? s_neigh : s_self) + (size_t)i * heads, *da = (b ? da_neigh : da_self) + (size_t)i * heads;
      float *dz = (b ? dZ_neigh : dZ_self) + (size_t)i * D;
#pragma unroll
      for (int v = 0; v < 2; v++) {
        if (!m.on[v]) continue;
        const float s = sv[m.head[v]], ds = da[m.head[v]] * (s > 0.f ? 1.f : 0.2f);
        const float4 a = ld4(h + 4 * (lane + 32 * v), true), gin = ld4(dh + 4 * (lane + 32 * v), true);
        float4 o;
        o.x = (gin.x + ds * at[b][v].x) * act_grad_from_output(a.x, act); o.y = (gin.y + ds * at[b][v].y) * act_grad_from_output(a.y, act);
        o.z = (gin.z + ds * at[b][v].z) * act_grad_from_output(a.z, act); o.w = (gin.w + ds * at[b][v].w) * act_grad_from_output(a.w, act);
        *reinterpret_cast<float4 *>(dz + 4 * (lane + 32 * v)) = o;
        pa[b][v].x += ds * a.x; pa[b][v].y += ds * a.y; pa[b][v].z += ds * a.z; pa[b][v].w += ds * a.w;
        pb[b][v].x += o.x; pb[b][v].y += o.y; pb[b][v].z += o.z; pb[b][v].w += o.w;
      }
    }
  }
  float *mine = sh + (size_t)warp * 4 * D;
#pragma unroll
  for (int b = 0; b < 2; b++)
#pragma unroll
    for (int v = 0; v < 2; v++)
      if (m.on[v]) {
        reinterpret_cast<float4 *>(mine + (b * 2 + 0) * D)[lane + 32 * v] = pa[b][v];
        reinterpret_cast<float4 *>(mine + (b * 2 + 1) * D)[lane + 32 * v] = pb[b][v];
      }
  __syncthreads();
  for (int f = threadIdx.x; f < 4 * D; f += blockDim.x) {
    float v = 0.f;
    for (int w = 0; w < wpb; w++) v += sh[(size_t)w * 4 * D + f];
    const int b = f / (2 * D), plane = (f / D) & 1, c = f % D;
    partials[(size_t)blockIdx.x * 6 * D + (b * 3 + (plane ? 2 : 0)) * D + c] = v;
  }
  for (int f = threadIdx.x; f < 2 * D; f += blockDim.x) partials[(size_t)blockIdx.x * 6 * D + ((f / D) * 3 + 1) * D + f % D] = 0.f;
}

inline int grid_rows(int n, int wpb, int cap) { return std::max(1, std::min((n + wpb - 1) / wpb, cap)); }
inline bool shape_ok(int heads, int d) {
  return heads >= 1 && heads <= GAT_MAXH && heads * d <= 256 && (d == 16 || d == 32 || d == 64 || d == 128 || d == 256);
}
int num_sms() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

}  // namespace

// colsum_finish_kernel lives in layers.cu: [parts][nb][3][D] -> dst[b][plane] += sum over parts (NULL destinations are skipped)
extern "C" int shadow_colsum_finish_f32(const float *partials, int32_t nparts, int32_t D, float *d00, float *d01, float *d02, float *d10, float *d11, float *d12,
                                        void *cuda_stream);

#define ST(s) ((cudaStream_t)(s))
#define WPB (GAT_BLOCK / 32)

extern "C" int shadow_gat_supported(int32_t heads, int32_t d) { return shape_ok(heads, d) ? 1 : 0; }
extern "C" int64_t shadow_gat_scratch_floats(int32_t heads, int32_t d) { return (int64_t)2 * num_sms() * 6 * heads * d; }

extern "C" int shadow_gat_logits_fwd_f32(const float *h_self, const float *h_neigh, const float *att, float *s_self, float *s_neigh, float *a_self, float *a_neigh,
                                         int32_t n, int32_t heads, int32_t d, void *stream) {
  if (n <= 0) return 0;
  if (!shape_ok(heads, d)) FAIL(SHADOW_EINVAL, "gat: unsupported shape heads=%d d=%d", heads, d);
  gat_logits_fwd_kernel<<<grid_rows(n, WPB, 8192), GAT_BLOCK, 0, ST(stream)>>>(h_self, h_neigh, att, s_self, s_neigh, a_self, a_neigh, n, heads, d);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

#define GAT_DISPATCH_NH(CALL)                 \
  switch (heads) {                            \
    case 1: CALL(1); break;                   \
    case 2: CALL(2); break;                   \
    case 4: CALL(4); break;                   \
    case 8: CALL(8); break;                   \
    default: FAIL(SHADOW_EINVAL, "gat: heads must be 1, 2, 4 or 8 (got %d)", heads); \
  }

extern "C" int shadow_gat_agg_fwd_f32(const int32_t *row_span, const int32_t *col, int32_t col_off, const float *val, const float *a_self, const float *a_neigh,
                                      const float *Hn, float *out, float *rowmax, float *denom, int32_t n, int32_t heads, int32_t d, void *stream) {
  if (n <= 0) return 0;
  if (!shape_ok(heads, d)) FAIL(SHADOW_EINVAL, "gat: unsupported shape heads=%d d=%d", heads, d);
#define CALL(NH) gat_agg_fwd_kernel<NH><<<grid_rows(n, WPB, 8192), GAT_BLOCK, 0, ST(stream)>>>((const int2 *)row_span, col, col_off, val, a_self, a_neigh, Hn, out, rowmax, denom, n, d)
  GAT_DISPATCH_NH(CALL)
#undef CALL
  CUDA_TRY(cudaGetLastError());
  return 0;
}
extern "C" int shadow_gat_agg_bwd_f32(const int32_t *row_span, const int32_t *col, int32_t col_off, const float *val, const float *a_self, const float *a_neigh,
                                      const float *Hn, const float *out, const float *rowmax, const float *denom, const float *dOut, float *dHn, float *da_self,
                                      float *da_neigh, int32_t n, int32_t heads, int32_t d, void *stream) {
  if (n <= 0) return 0;
  if (!shape_ok(heads, d)) FAIL(SHADOW_EINVAL, "gat: unsupported shape heads=%d d=%d", heads, d);
#define CALL(NH) gat_agg_bwd_kernel<NH><<<grid_rows(n, WPB, 8192), GAT_BLOCK, 0, ST(stream)>>>((const int2 *)row_span, col, col_off, val, a_self, a_neigh, Hn, out, rowmax, denom, dOut, dHn, da_self, da_neigh, n, d)
  GAT_DISPATCH_NH(CALL)
#undef CALL
  CUDA_TRY(cudaGetLastError());
  return 0;
}
extern "C" int shadow_gat_headnorm_fwd_f32(const float *h_self, const float *agg, const float *scale, const float *offset, float *out, float *mean, float *rstd,
                                           int32_t n, int32_t heads, int32_t d, int32_t do_norm, void *stream) {
  if (n <= 0) return 0;
  if (!shape_ok(heads, d)) FAIL(SHADOW_EINVAL, "gat: unsupported shape heads=%d d=%d", heads, d);
  gat_headnorm_fwd_kernel<<<grid_rows(n, WPB, 8192), GAT_BLOCK, 0, ST(stream)>>>(h_self, agg, scale, offset, out, mean, rstd, n, heads, d, do_norm);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
extern "C" int shadow_gat_headnorm_bwd_f32(const float *dOut, const float *h_self, const float *agg, const float *scale, const float *mean, const float *rstd,
                                           float *dh_self, float *dagg, float *dscale, float *doffset, int32_t n, int32_t heads, int32_t d, int32_t do_norm,
                                           float *scratch, int64_t scratch_floats, void *stream) {
  if (n <= 0) return 0;
  if (!shape_ok(heads, d)) FAIL(SHADOW_EINVAL, "gat: unsupported shape heads=%d d=%d", heads, d);
  const int D = heads * d, grid = grid_rows(n, WPB, 2 * num_sms());
  if (!scratch || scratch_floats < (int64_t)grid * 6 * D) FAIL(SHADOW_EINVAL, "gat_headnorm_bwd: scratch too small");
  const size_t smem = (size_t)WPB * 4 * D * sizeof(float);
  gat_headnorm_bwd_kernel<<<grid, GAT_BLOCK, smem, ST(stream)>>>(dOut, h_self, agg, scale, mean, rstd, dh_self, dagg, scratch, n, heads, d, do_norm);
  CUDA_TRY(cudaGetLastError());
  if (do_norm) return shadow_colsum_finish_f32(scratch, grid, D, dscale, doffset, nullptr, dscale + D, doffset + D, nullptr, stream);
  return 0;
}
extern "C" int shadow_gat_pre_bwd_f32(const float *dh_self, const float *dh_neigh, const float *h_self, const float *h_neigh, const float *s_self, const float *s_neigh,
                                      const float *da_self, const float *da_neigh, const float *att, float *dZ_self, float *dZ_neigh, float *datt, float *dbias_self,
                                      float *dbias_neigh, int32_t n, int32_t heads, int32_t d, int32_t act, float *scratch, int64_t scratch_floats, void *stream) {
  if (n <= 0) return 0;
  if (!shape_ok(heads, d)) FAIL(SHADOW_EINVAL, "gat: unsupported shape heads=%d d=%d", heads, d);
  const int D = heads * d, grid = grid_rows(n, WPB, 2 * num_sms());
  if (!scratch || scratch_floats < (int64_t)grid * 6 * D) FAIL(SHADOW_EINVAL, "gat_pre_bwd: scratch too small");
  const size_t smem = (size_t)WPB * 4 * D * sizeof(float);
  gat_pre_bwd_kernel<<<grid, GAT_BLOCK, smem, ST(stream)>>>(dh_self, dh_neigh, h_self, h_neigh, s_self, s_neigh, da_self, da_neigh, att, dZ_self, dZ_neigh, scratch, n,
                                                           heads, d, act);
  CUDA_TRY(cudaGetLastError());
  return shadow_colsum_finish_f32(scratch, grid, D, datt, nullptr, dbias_self, datt ? datt + D : nullptr, nullptr, dbias_neigh, stream);
}
