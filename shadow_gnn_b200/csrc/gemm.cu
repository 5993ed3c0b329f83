// gemm.cu -- the dense Linear of the message-passing layers (nn.Linear in shaDow/layers.py:421,451-452,501-505,556,670 and
// models.py:81) on the tensor cores, fp32 in / fp32 out.
//
// C[M,N] (+)= op(A)[M,K] * op(B)[K,N] (+ bias[N]) with every product formed as an error-compensated 3xTF32 sum
//     a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi,   x_hi = tf32(x), x_lo = tf32(x - x_hi)
// accumulated in fp32 by mma.sync.m16n8k8.tf32: the result is fp32-accurate (the parity budget is 1e-3 relative through 5 layers,
// which one plain TF32 pass per layer does not keep), and a 4,832 x 256 x 256 product takes a few microseconds instead of the ~18 us of
// the fp32 SIMT library GEMM.  The three operand layouts are exactly the three products of a Linear:
//   forward   Z  = X  W^T + b      A = X  [M,K] row-major,          B = W  [N,K] row-major  (b_kn = 0)
//   dgrad     dX = dZ W            A = dZ [M,N2] row-major,         B = W  [N2,K2] = [K][N] (b_kn = 1)
//   wgrad     dW += dZ^T X         A = dZ read transposed (ta = 1), B = X  [Mred,K2] = [K][N] (b_kn = 1), split along Mred, fp32 atomics
// One CTA = 4 warps computes a 64 x 64 tile (warp tile 32 x 32); operand tiles go global -> registers (128-bit words, one k-tile ahead of
// their use) -> shared memory with conflict-free fragment reads.  These GEMMs are tiny (0.6 GFLOP) and L2-resident: they are bound by launch latency and the tile pipeline, not by
// tensor-core peak, which is why this is the legacy warp-level MMA and not a tcgen05/TMEM pipeline ("what comes next").
#include <algorithm>

#include "common.cuh"

#define GB_M 64
#define GB_N 64
#define GB_K 32

__device__ __forceinline__ void split_tf32(const float x, uint32_t &hi, uint32_t &lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  const float r = x - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// operand tile loaders: global -> registers (issued one k-tile ahead of its use) -> shared memory.
// VEC: every operand row is a multiple of 4 floats and 16-byte aligned, so a tile moves as 128-bit words; otherwise scalar words.
template <bool TA, bool VEC>
struct TileA {
  float4 v[VEC ? 4 : 16];
  __device__ __forceinline__ void load(const float *__restrict__ A, const int lda, const int m0, const int k0, const int M, const int ke, const int tid) {
    if (VEC) {
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int e = tid + i * 128;
        if (!TA) { const int m = e >> 3, k = (e & 7) * 4; v[i] = (m0 + m < M && k0 + k < ke) ? *reinterpret_cast<const float4 *>(A + (size_t)(m0 + m) * lda + k0 + k) : make_float4(0.f, 0.f, 0.f, 0.f); }
        else { const int k = e >> 4, m = (e & 15) * 4; v[i] = (m0 + m < M && k0 + k < ke) ? *reinterpret_cast<const float4 *>(A + (size_t)(k0 + k) * lda + m0 + m) : make_float4(0.f, 0.f, 0.f, 0.f); }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; i++) {
        const int e = tid + i * 128;
        if (!TA) { const int m = e / GB_K, k = e % GB_K; v[i].x = (m0 + m < M && k0 + k < ke) ? A[(size_t)(m0 + m) * lda + k0 + k] : 0.f; }
        else { const int k = e / GB_M, m = e % GB_M; v[i].x = (m0 + m < M && k0 + k < ke) ? A[(size_t)(k0 + k) * lda + m0 + m] : 0.f; }
      }
    }
  }
  __device__ __forceinline__ void store(float (*As)[GB_K + 4], const int tid) const {
    if (VEC) {
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int e = tid + i * 128;
        if (!TA) { const int m = e >> 3, k = (e & 7) * 4; *reinterpret_cast<float4 *>(&As[m][k]) = v[i]; }
        else { const int k = e >> 4, m = (e & 15) * 4; As[m][k] = v[i].x; As[m + 1][k] = v[i].y; As[m + 2][k] = v[i].z; As[m + 3][k] = v[i].w; }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; i++) {
        const int e = tid + i * 128;
        if (!TA) As[e / GB_K][e % GB_K] = v[i].x;
        else As[e % GB_M][e / GB_M] = v[i].x;
      }
    }
  }
};
template <bool BKN, bool VEC>
struct TileB {
  float4 v[VEC ? 4 : 16];
  __device__ __forceinline__ void load(const float *__restrict__ B, const int ldb, const int n0, const int k0, const int N, const int ke, const int tid) {
    if (VEC) {
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int e = tid + i * 128;
        if (!BKN) { const int n = e >> 3, k = (e & 7) * 4; v[i] = (n0 + n < N && k0 + k < ke) ? *reinterpret_cast<const float4 *>(B + (size_t)(n0 + n) * ldb + k0 + k) : make_float4(0.f, 0.f, 0.f, 0.f); }
        else { const int k = e >> 4, n = (e & 15) * 4; v[i] = (n0 + n < N && k0 + k < ke) ? *reinterpret_cast<const float4 *>(B + (size_t)(k0 + k) * ldb + n0 + n) : make_float4(0.f, 0.f, 0.f, 0.f); }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; i++) {
        const int e = tid + i * 128;
        if (!BKN) { const int n = e / GB_K, k = e % GB_K; v[i].x = (n0 + n < N && k0 + k < ke) ? B[(size_t)(n0 + n) * ldb + k0 + k] : 0.f; }
        else { const int k = e / GB_N, n = e % GB_N; v[i].x = (n0 + n < N && k0 + k < ke) ? B[(size_t)(k0 + k) * ldb + n0 + n] : 0.f; }
      }
    }
  }
  template <class BsT>
  __device__ __forceinline__ void store(BsT &Bs, const int tid) const {
    if (VEC) {
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int e = tid + i * 128;
        if (!BKN) { const int n = e >> 3, k = (e & 7) * 4; *reinterpret_cast<float4 *>(&Bs[n][k]) = v[i]; }
        else { const int k = e >> 4, n = (e & 15) * 4; *reinterpret_cast<float4 *>(&Bs[k][n]) = v[i]; }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; i++) {
        const int e = tid + i * 128;
        if (!BKN) Bs[e / GB_K][e % GB_K] = v[i].x;
        else Bs[e / GB_N][e % GB_N] = v[i].x;
      }
    }
  }
};

// Two products of the same shape can share one launch (GraphSAGE's self / neighbour branch: twice the resident warps for the MMA pipe,
// half the launches): problem = blockIdx.z % nprob, k-slice = blockIdx.z / nprob.
struct GemmPair { const float *A[2]; const float *B[2]; float *C[2]; const float *bias[2]; int nprob; };

template <bool TA, bool BKN, bool VEC>
__global__ void __launch_bounds__(128) gemm_tf32x3_kernel(const GemmPair P, const int lda, const int ldb, const int ldc, const int M,
                                                          const int N, const int K, const int k_per_split, const int atomic_out) {
  const int prob = blockIdx.z % P.nprob, zsplit = blockIdx.z / P.nprob;
  const float *__restrict__ A = P.A[prob];
  const float *__restrict__ B = P.B[prob];
  float *__restrict__ C = P.C[prob];
  const float *__restrict__ bias = P.bias[prob];
  __shared__ __align__(16) float As[GB_M][GB_K + 4];                                       // [m][k]
  __shared__ __align__(16) float Bs[BKN ? GB_K : GB_N][BKN ? GB_N + 8 : GB_K + 4];          // [k][n] or [n][k]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
  const int m0 = blockIdx.x * GB_M, n0 = blockIdx.y * GB_N;
  const int kb = zsplit * k_per_split, ke = min(K, kb + k_per_split);
  float acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
      for (int q = 0; q < 4; q++) acc[i][j][q] = 0.f;

  TileA<TA, VEC> ta;
  TileB<BKN, VEC> tb;
  ta.load(A, lda, m0, kb, M, ke, tid);
  tb.load(B, ldb, n0, kb, N, ke, tid);
  for (int k0 = kb; k0 < ke; k0 += GB_K) {
    ta.store(As, tid);
    tb.store(Bs, tid);
    __syncthreads();
    if (k0 + GB_K < ke) {                              // the next tile's loads fly while this one is multiplied
      ta.load(A, lda, m0, k0 + GB_K, M, ke, tid);
      tb.load(B, ldb, n0, k0 + GB_K, N, ke, tid);
    }
#pragma unroll
    for (int kk = 0; kk < GB_K; kk += 8) {
      uint32_t ah[2][4], al[2][4], bh[4][2], bl[4][2];
#pragma unroll
      for (int mi = 0; mi < 2; mi++) {                 // A fragment (16 x 8, row): a0 (g, t)  a1 (g+8, t)  a2 (g, t+4)  a3 (g+8, t+4)
        const int r = wm + mi * 16 + g;
        split_tf32(As[r][kk + t], ah[mi][0], al[mi][0]);
        split_tf32(As[r + 8][kk + t], ah[mi][1], al[mi][1]);
        split_tf32(As[r][kk + t + 4], ah[mi][2], al[mi][2]);
        split_tf32(As[r + 8][kk + t + 4], ah[mi][3], al[mi][3]);
      }
#pragma unroll
      for (int ni = 0; ni < 4; ni++) {                 // B fragment (8 x 8, col): b0 (k = t, n = g)  b1 (k = t+4, n = g)
        const int c = wn + ni * 8 + g;
        const float b0 = BKN ? Bs[kk + t][c] : Bs[c][kk + t];
        const float b1 = BKN ? Bs[kk + t + 4][c] : Bs[c][kk + t + 4];
        split_tf32(b0, bh[ni][0], bl[ni][0]);
        split_tf32(b1, bh[ni][1], bl[ni][1]);
      }
#pragma unroll
      for (int mi = 0; mi < 2; mi++)
#pragma unroll
        for (int ni = 0; ni < 4; ni++) {               // small terms first
          mma_tf32(acc[mi][ni], al[mi], bh[ni]);
          mma_tf32(acc[mi][ni], ah[mi], bl[ni]);
          mma_tf32(acc[mi][ni], ah[mi], bh[ni]);
        }
    }
    __syncthreads();
  }
  // ---- epilogue: accumulator (16 x 8): c0 (g, 2t)  c1 (g, 2t+1)  c2 (g+8, 2t)  c3 (g+8, 2t+1) ----
  const bool add_bias = bias != nullptr && zsplit == 0;
#pragma unroll
  for (int mi = 0; mi < 2; mi++)
#pragma unroll
    for (int ni = 0; ni < 4; ni++)
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int m = m0 + wm + mi * 16 + g + h * 8, n = n0 + wn + ni * 8 + 2 * t;
        if (m < M) {
          float v0 = acc[mi][ni][2 * h], v1 = acc[mi][ni][2 * h + 1];
          if (add_bias) { if (n < N) v0 += bias[n]; if (n + 1 < N) v1 += bias[n + 1]; }
          float *dst = &C[(size_t)m * ldc + n];
          if (atomic_out) { if (n < N) atomicAdd(dst, v0); if (n + 1 < N) atomicAdd(dst + 1, v1); }
          else if (VEC && n + 1 < N) *reinterpret_cast<float2 *>(dst) = make_float2(v0, v1);
          else { if (n < N) dst[0] = v0; if (n + 1 < N) dst[1] = v1; }
        }
      }
}

template <bool TA, bool BKN>
static void launch_gemm(bool vec, dim3 grid, cudaStream_t st, const GemmPair &P, int lda, int ldb, int ldc, int M, int N, int K, int kps, int atomic_out) {
  if (vec) gemm_tf32x3_kernel<TA, BKN, true><<<grid, 128, 0, st>>>(P, lda, ldb, ldc, M, N, K, kps, atomic_out);
  else gemm_tf32x3_kernel<TA, BKN, false><<<grid, 128, 0, st>>>(P, lda, ldb, ldc, M, N, K, kps, atomic_out);
}

// Small x W^T (+ bias): a handful of rows against a weight matrix (the 47-class classifier on the 32 pooled rows of a batch).  One tile
// of the tensor-core kernel would walk K alone (13 us at K = 256: latency); here every output is one warp's dot product -- lanes stride
// along K (coalesced in both operands), fixed-order shuffle reduction, plain fp32 FMA.
__global__ void __launch_bounds__(256) small_linear_kernel(const float *__restrict__ A, int lda, const float *__restrict__ B, int ldb, float *__restrict__ C, int ldc,
                                                           const float *__restrict__ bias, int M, int N, int K, int accumulate) {
  const int lane = threadIdx.x & 31;
  const long long total = (long long)M * N;
  for (long long o = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); o < total; o += (long long)gridDim.x * (blockDim.x >> 5)) {
    const int m = (int)(o / N), n = (int)(o - (long long)m * N);
    const float *a = A + (long long)m * lda, *b = B + (long long)n * ldb;
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) acc = fmaf(a[k], b[k], acc);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if (lane == 0) {
      float v = acc + (bias ? bias[n] : 0.f);
      float *c = C + (long long)m * ldc + n;
      *c = accumulate ? *c + v : v;
    }
  }
}

static int gemm_dispatch(const GemmPair &P, int32_t lda, int32_t trans_a, int32_t ldb, int32_t b_kn, int32_t ldc, int32_t M, int32_t N, int32_t K,
                         int32_t accumulate, int32_t split_k, void *cuda_stream) {
  for (int i = 0; i < P.nprob; i++)
    if (!P.A[i] || !P.B[i] || !P.C[i]) FAIL(SHADOW_EINVAL, "shadow_gemm_tf32x3: NULL operand");
  if (M < 0 || N < 0 || K < 0) FAIL(SHADOW_EINVAL, "shadow_gemm_tf32x3: negative dimension");
  if (M == 0 || N == 0) return 0;
  if (P.nprob == 1 && !trans_a && !b_kn && split_k <= 1 && (long long)M * N <= 16384 && K > 0) {
    const long long warps = (long long)M * N;
    small_linear_kernel<<<(int)std::min<long long>((warps + 7) / 8, 4096), 256, 0, (cudaStream_t)cuda_stream>>>(P.A[0], lda, P.B[0], ldb, P.C[0], ldc, P.bias[0], M, N, K,
                                                                                                             accumulate ? 1 : 0);
    CUDA_TRY(cudaGetLastError());
    return 0;
  }
  int splits = std::max(1, split_k);
  int kps = ((K + splits - 1) / splits + GB_K - 1) / GB_K * GB_K;
  if (kps <= 0) kps = GB_K;
  splits = std::max(1, (K + kps - 1) / kps);
  if (splits > 1 && !accumulate) FAIL(SHADOW_EINVAL, "shadow_gemm_tf32x3: split_k > 1 needs accumulate = 1 (C += ..., C initialised by the caller)");
  const dim3 grid((M + GB_M - 1) / GB_M, (N + GB_N - 1) / GB_N, splits * P.nprob);
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const int atomic_out = accumulate ? 1 : 0;
  // 128-bit tile moves need: 16-byte aligned bases, leading dimensions multiples of 4, and the contiguous extent of each operand
  // (K for row-major A / weight-layout B, M for transposed A, N for [K][N] B) a multiple of 4; C rows even for the paired stores
  auto al16 = [](const void *p) { return ((uintptr_t)p & 15) == 0; };
  bool vec = (lda % 4 == 0) && (ldb % 4 == 0) && (ldc % 2 == 0) && ((trans_a ? M : K) % 4 == 0) && ((b_kn ? N : K) % 4 == 0) && (kps % 4 == 0);
  for (int i = 0; i < P.nprob; i++) vec = vec && al16(P.A[i]) && al16(P.B[i]) && al16(P.C[i]);
  if (!trans_a && !b_kn) launch_gemm<false, false>(vec, grid, st, P, lda, ldb, ldc, M, N, K, kps, atomic_out);
  else if (!trans_a && b_kn) launch_gemm<false, true>(vec, grid, st, P, lda, ldb, ldc, M, N, K, kps, atomic_out);
  else if (trans_a && b_kn) launch_gemm<true, true>(vec, grid, st, P, lda, ldb, ldc, M, N, K, kps, atomic_out);
  else launch_gemm<true, false>(vec, grid, st, P, lda, ldb, ldc, M, N, K, kps, atomic_out);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

extern "C" int shadow_gemm_tf32x3_f32(const float *A, int32_t lda, int32_t trans_a, const float *B, int32_t ldb, int32_t b_kn, float *C, int32_t ldc,
                                      const float *bias, int32_t M, int32_t N, int32_t K, int32_t accumulate, int32_t split_k, void *cuda_stream) {
  GemmPair P;
  P.A[0] = P.A[1] = A; P.B[0] = P.B[1] = B; P.C[0] = P.C[1] = C; P.bias[0] = P.bias[1] = bias; P.nprob = 1;
  return gemm_dispatch(P, lda, trans_a, ldb, b_kn, ldc, M, N, K, accumulate, split_k, cuda_stream);
}

extern "C" int shadow_gemm_tf32x3_pair_f32(const float *A0, const float *A1, int32_t lda, int32_t trans_a, const float *B0, const float *B1, int32_t ldb,
                                           int32_t b_kn, float *C0, float *C1, int32_t ldc, const float *bias0, const float *bias1, int32_t M, int32_t N,
                                           int32_t K, int32_t accumulate, int32_t split_k, void *cuda_stream) {
  GemmPair P;
  P.A[0] = A0; P.A[1] = A1; P.B[0] = B0; P.B[1] = B1; P.C[0] = C0; P.C[1] = C1; P.bias[0] = bias0; P.bias[1] = bias1; P.nprob = 2;
  return gemm_dispatch(P, lda, trans_a, ldb, b_kn, ldc, M, N, K, accumulate, split_k, cuda_stream);
}
