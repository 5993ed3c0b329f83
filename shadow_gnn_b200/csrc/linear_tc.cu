// linear_tc.cu -- the dense Linear of the shaDow layers (shaDow/layers.py:329-338,421,451-452,474-483) on the 5th-generation tensor
// cores, hand-written for sm_100a: TMA-staged operands, tcgen05.mma with the accumulator in TMEM, fused epilogue out of tcgen05.ld.
//
//     Z   = X W^T + b                                  X [M, K]  W [N, K]  (both K-major, fp32)         N <= 256
//     out = [out +] norm_feat(act(Z))                  norm_feat over the N features of a row, learned scale / offset
//
// One CTA owns 128 rows and ALL N (<= 256) columns: its fp32 accumulator is 128 lanes x N columns of TMEM (256 of the 512 columns), so the
// row statistics of norm_feat never leave the SM.  The reference is fp32 (1e-3 layer budget over 5 layers), so every product is an
// error-compensated 3xTF32 sum: each fp32 operand tile that TMA lands in shared memory (128B-swizzled, K-major) is split IN PLACE into
// hi = the value with its low 13 mantissa bits cleared (exactly what a TF32 tensor core reads) and lo = x - hi (exact in fp32) in a second
// tile at the same offsets, and three tcgen05.mma.kind::tf32 run per 8-wide k-step:  hi*lo + lo*hi + hi*hi  (small terms first).
//
// Warp roles (192 threads):  warp 0 = TMA producer (one elected lane), warp 1 = TMEM allocation + MMA issue (one elected lane),
// warps 2-5 = split the landed tiles (the "transform" stage), then the epilogue: each thread owns one accumulator row (TMEM lane),
// three passes over its N columns (sum -> mean, squared deviations -> rstd, normalise + store), bias / activation applied on the fly.
// Pipeline: 2 stages x (A hi/lo 2 x 16 KB + B hi/lo 2 x 32 KB) = 192 KB of shared memory; mbarriers full (TMA -> transform),
// ready (transform -> MMA), empty (tcgen05.commit -> producer), acc (last commit -> epilogue).
#include <algorithm>
#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/shadow_b200.h"
#include "common.cuh"

namespace {

constexpr int BLOCK_M = 128;          // rows per CTA = TMEM lanes
constexpr int BLOCK_N = 256;          // columns per CTA = TMEM columns (fp32)
#ifndef LTC_BLOCK_K
#define LTC_BLOCK_K 16
#endif
constexpr int BLOCK_K = LTC_BLOCK_K;  // fp32 elements per k-block = one swizzle row: 16 (64-byte swizzle, 4 stages of 48 KB) or 32 (128-byte swizzle, 2 stages of 96 KB)
constexpr int UMMA_K = 8;             // tf32: 32 bytes per MMA k-step
constexpr int NSTAGES = BLOCK_K == 16 ? 4 : 2;
static_assert(BLOCK_K == 16 || BLOCK_K == 32, "one swizzle row per k-block");
constexpr int A_TILE = BLOCK_M * BLOCK_K * 4;      // 16 KB
constexpr int B_TILE = BLOCK_N * BLOCK_K * 4;      // 32 KB
constexpr int STAGE_BYTES = 2 * A_TILE + 2 * B_TILE;
constexpr int SMEM_BYTES = NSTAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 128 /*barriers*/ + 3 * BLOCK_N * 4 /*bias, scale, offset*/ + 2 * 2 * BLOCK_M * 4 /*row statistics of the two column halves*/;
static_assert(NSTAGES * STAGE_BYTES >= 8 * 4 * 4096, "the epilogue stages its tiles in the pipeline buffers");
constexpr int NUM_WORKERS = 256;      // transform + epilogue threads (warps 2..9: two warps per TMEM lane quadrant)
constexpr int NUM_THREADS = 64 + NUM_WORKERS;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
               "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, const void *src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap *map, const void *src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
// shared-memory matrix descriptor, K-major, swizzled (cute::UMMA::SmemDescriptor): start >> 4 | LBO (unused: 1) << 16 | SBO (8 rows of one
// swizzle row each: 1024 B or 512 B) >> 4 << 32 | version 1 << 46 | layout (2 = SWIZZLE_128B, 4 = SWIZZLE_64B) << 61
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  constexpr uint64_t sbo = (8 * BLOCK_K * 4) >> 4, layout = BLOCK_K == 32 ? 2 : 4;
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor), kind::tf32: D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10), both K-major, N >> 3 << 17, M >> 4 << 24
__device__ __forceinline__ uint32_t umma_idesc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {      // arrives on `bar` once every MMA issued so far has completed (implies fence::before_thread_sync)
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
      "%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
        "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
        "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

enum { ACT_RELU = 0, ACT_I = 1, ACT_ELU = 2, ACT_TANH = 3, ACT_LRELU = 4 };
template <int ACT>
__device__ __forceinline__ float act_f(float z) {
  if (ACT == ACT_RELU) return z > 0.f ? z : 0.f;
  if (ACT == ACT_ELU) return z > 0.f ? z : __expf(z) - 1.f;          // absolute error ~1e-7 on a value in (-1, 0]
  if (ACT == ACT_TANH) return tanhf(z);
  if (ACT == ACT_LRELU) return z > 0.f ? z : 0.2f * z;
  return z;
}

// One launch = up to two products of the same shape (the self and the neighbour branch of a GraphSAGE layer, layers.py:474-483; blockIdx.y).
struct LinearBranch {
  const float *bias, *scale, *offset;      // [N] each (bias may be NULL; scale / offset only with norm_feat)
  float *Z, *out, *mean, *rstd;            // Z [M, ldz] pre-activation (saved for backward; may be NULL), out [M, ldo], mean / rstd [M] (NULL without norm)
};
struct LinearParams {
  LinearBranch br[2];
  int ldz, ldo, M, N, K, act, do_norm;
  int presplit;                            // the weights arrive as two planes (TF32 head, exact remainder): no in-kernel split of the B tiles
  int debug;                               // SHADOW_LTC_DEBUG bits (timing experiments only): 1 skip the main loop, 2 skip the epilogue
  int out_mode;                            // 0: out = o   1: out += o (one branch per launch)   2: red.global.add (both branches into one zeroed buffer)
};

template <int ACT, bool NORM>
__global__ void __launch_bounds__(NUM_THREADS, 1)
linear_tc_kernel(const __grid_constant__ CUtensorMap map_x0, const __grid_constant__ CUtensorMap map_w0, const __grid_constant__ CUtensorMap map_x1,
                 const __grid_constant__ CUtensorMap map_w1, const __grid_constant__ CUtensorMap map_z0, const __grid_constant__ CUtensorMap map_o0,
                 const __grid_constant__ CUtensorMap map_z1, const __grid_constant__ CUtensorMap map_o1, const __grid_constant__ CUtensorMap map_wl0,
                 const __grid_constant__ CUtensorMap map_wl1, const LinearParams P) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char *smem = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);      // SWIZZLE_128B atoms are 1024-byte aligned
  uint64_t *bars = (uint64_t *)(smem + NSTAGES * STAGE_BYTES);
  uint64_t *full = bars, *ready = bars + NSTAGES, *empty = bars + 2 * NSTAGES, *acc = bars + 3 * NSTAGES;
  uint32_t *tmem_slot = (uint32_t *)(bars + 3 * NSTAGES + 1);
  float *vecs = (float *)(bars + 3 * NSTAGES + 2);                 // bias / scale / offset of this branch (3 x BLOCK_N floats)
  float *part = vecs + 3 * BLOCK_N;                                // [2 passes][2 column halves][BLOCK_M] partial row sums
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BLOCK_M;
  const int num_kb = (P.debug & 1) ? 0 : (P.K + BLOCK_K - 1) / BLOCK_K;
  // column split (gridDim.z = 2): this CTA owns output columns [n0, n0 + bn) -- half the MMA work and half the epilogue per CTA, twice the CTAs
  // (a 4,832-row batch is only 38 row tiles).  Only without norm_feat: its row statistics need every column of the row.
  const int bn = BLOCK_N / (int)gridDim.z, n0 = (int)blockIdx.z * bn;
  const int b_tile = bn * BLOCK_K * 4;
  const CUtensorMap *map_x = blockIdx.y ? &map_x1 : &map_x0, *map_w = blockIdx.y ? &map_w1 : &map_w0;
  const CUtensorMap *map_z = blockIdx.y ? &map_z1 : &map_z0, *map_o = blockIdx.y ? &map_o1 : &map_o0, *map_wl = blockIdx.y ? &map_wl1 : &map_wl0;
  const LinearBranch B = P.br[blockIdx.y];

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(map_w) : "memory");
    for (int s = 0; s < NSTAGES; s++) { mbar_init(&full[s], 1); mbar_init(&ready[s], NUM_WORKERS); mbar_init(&empty[s], 1); }
    mbar_init(acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {                                   // TMEM: 256 columns x 128 lanes of fp32 accumulator
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(BLOCK_N) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < BLOCK_N; i += NUM_THREADS) {
    const bool in = n0 + i < P.N && i < bn;
    vecs[i] = (in && B.bias) ? B.bias[n0 + i] : 0.f;
    vecs[BLOCK_N + i] = (in && NORM) ? B.scale[n0 + i] : 1.f;
    vecs[2 * BLOCK_N + i] = (in && NORM) ? B.offset[n0 + i] : 0.f;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; kb++) {
        const int s = kb % NSTAGES;
        mbar_wait(&empty[s], ((kb / NSTAGES) & 1) ^ 1);
        unsigned char *st = smem + s * STAGE_BYTES;
        mbar_expect_tx(&full[s], A_TILE + (P.presplit ? 2 : 1) * b_tile);
        tma_load_2d(st, map_x, &full[s], kb * BLOCK_K, m0);                       // A     [128 rows][BLOCK_K]
        tma_load_2d(st + 2 * A_TILE, map_w, &full[s], kb * BLOCK_K, n0);          // B (hi) [bn rows][BLOCK_K]
        if (P.presplit) tma_load_2d(st + 2 * A_TILE + B_TILE, map_wl, &full[s], kb * BLOCK_K, n0);     // B lo
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(BLOCK_M, bn);
      for (int kb = 0; kb < num_kb; kb++) {
        const int s = kb % NSTAGES;
        mbar_wait(&ready[s], (kb / NSTAGES) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_hi = smem_u32(smem + s * STAGE_BYTES), a_lo = a_hi + A_TILE, b_hi = a_hi + 2 * A_TILE, b_lo = b_hi + B_TILE;
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; k++) {
          const uint32_t ko = k * UMMA_K * 4;          // byte offset of the k-step inside the 128-byte swizzle row
          umma_tf32(tmem_base, umma_desc_sw128(a_hi + ko), umma_desc_sw128(b_lo + ko), idesc, (kb | k) ? 1u : 0u);
          umma_tf32(tmem_base, umma_desc_sw128(a_lo + ko), umma_desc_sw128(b_hi + ko), idesc, 1u);
          umma_tf32(tmem_base, umma_desc_sw128(a_hi + ko), umma_desc_sw128(b_hi + ko), idesc, 1u);
        }
        umma_commit(&empty[s]);                        // the stage's shared memory is free once these MMAs have read it
      }
      if (num_kb) umma_commit(acc);                    // accumulator complete
    }
  } else {
    // ===== transform warps (2..5): split every landed tile into hi / lo in place =====
    const int t = threadIdx.x - 64;                    // 0..255
    for (int kb = 0; kb < num_kb; kb++) {
      const int s = kb % NSTAGES;
      mbar_wait(&full[s], (kb / NSTAGES) & 1);
      float4 *a_hi = (float4 *)(smem + s * STAGE_BYTES), *a_lo = a_hi + A_TILE / 16, *b_hi = a_hi + 2 * A_TILE / 16, *b_lo = b_hi + B_TILE / 16;
#pragma unroll 4
      for (int i = t; i < A_TILE / 16; i += NUM_WORKERS) {
        const float4 v = a_hi[i];
        float4 h, l;
        h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
        h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
        l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
        a_hi[i] = h; a_lo[i] = l;
      }
      if (!P.presplit)
#pragma unroll 4
      for (int i = t; i < b_tile / 16; i += NUM_WORKERS) {
        const float4 v = b_hi[i];
        float4 h, l;
        h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
        h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
        l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
        b_hi[i] = h; b_lo[i] = l;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the tensor core (async proxy)
      mbar_arrive(&ready[s]);
    }

    // ===== epilogue: thread <-> accumulator row (TMEM lane); a warp may only touch lanes 32 * (warp % 4) .. + 31.  Two warps share a lane
    // quadrant: the 32-column blocks alternate between them (row statistics of norm_feat are combined through shared memory).  Every
    // 32 x 32 block is written to a 128B-swizzled staging tile (the pipeline stages are free by now) and leaves through TMA: a plain store
    // for Z, a store or a reduce-add for the output, so global memory sees whole 128-byte lines and out-of-range rows / columns are clipped
    // by the tensor map. =====
    if (num_kb) mbar_wait(acc, 0);
    if (!(P.debug & 2)) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int quad = warp & 3, half = (warp - 2) >> 2;
    const int row_l = quad * 32 + lane;                // row inside the CTA tile
    const int row = m0 + row_l;
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int N = min(bn, P.N - n0);                     // this CTA's columns (local index c <-> global column n0 + c)
    const float *vb = vecs, *vs = vecs + BLOCK_N, *vo = vecs + 2 * BLOCK_N;
    float mean = 0.f, rstd = 1.f;
    uint32_t r[32];
    if (NORM) {
      float s1 = 0.f;
      for (int c0 = half * 32; c0 < N; c0 += 64) {
        tmem_ld32(taddr + c0, r);
#pragma unroll
        for (int j = 0; j < 32; j++)
          if (c0 + j < N) s1 += act_f<ACT>(__uint_as_float(r[j]) + vb[c0 + j]);
      }
      part[half * BLOCK_M + row_l] = s1;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
      mean = (part[row_l] + part[BLOCK_M + row_l]) / (float)N;
      float s2 = 0.f;
      for (int c0 = half * 32; c0 < N; c0 += 64) {
        tmem_ld32(taddr + c0, r);
#pragma unroll
        for (int j = 0; j < 32; j++)
          if (c0 + j < N) { const float d = act_f<ACT>(__uint_as_float(r[j]) + vb[c0 + j]) - mean; s2 += d * d; }
      }
      part[2 * BLOCK_M + half * BLOCK_M + row_l] = s2;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
      rstd = rsqrtf((part[2 * BLOCK_M + row_l] + part[3 * BLOCK_M + row_l]) / (float)N + 1e-9f);
      if (half == 0 && row < P.M && B.mean) { B.mean[row] = mean; B.rstd[row] = rstd; }
    }
    unsigned char *stage_w = smem + (size_t)(warp - 2) * (4 * 4096);       // this warp: {Z, out} x 2 buffers of 32 rows x 128 bytes
    const uint32_t sw = (uint32_t)(lane & 7);                               // 128-byte swizzle: 16-byte chunk q of row l lives at chunk q ^ (l % 8)
    int it = 0;
    for (int c0 = half * 32; c0 < N; c0 += 64, it++) {
      tmem_ld32(taddr + c0, r);
      unsigned char *zb = stage_w + (size_t)(it & 1) * 8192, *ob = zb + 4096;
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");     // the stores that last read this buffer pair are done
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 8; q++) {
        float z[4], o[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const int j = 4 * q + e;
          z[e] = __uint_as_float(r[j]) + vb[c0 + j];
          o[e] = act_f<ACT>(z[e]);
          if (NORM) o[e] = (o[e] - mean) * vs[c0 + j] * rstd + vo[c0 + j];
        }
        const uint32_t off = (uint32_t)lane * 128u + ((((uint32_t)q) ^ sw) << 4);
        if (B.Z) *reinterpret_cast<float4 *>(zb + off) = make_float4(z[0], z[1], z[2], z[3]);
        *reinterpret_cast<float4 *>(ob + off) = make_float4(o[0], o[1], o[2], o[3]);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // generic-proxy writes -> visible to TMA
      __syncwarp();
      if (lane == 0) {
        if (B.Z) tma_store_2d(map_z, zb, n0 + c0, m0 + quad * 32);
        if (P.out_mode == 0) tma_store_2d(map_o, ob, n0 + c0, m0 + quad * 32);
        else tma_reduce_add_2d(map_o, ob, n0 + c0, m0 + quad * 32);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // shared memory stays valid until TMA has read it; writes complete
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(BLOCK_N) : "memory");
  }
}

// ================================================================================================================================
// Weight gradient  dW[n, k] = sum_m dZ[m, n] X[m, k]  (shaDow/layers.py:421,451-452 backward): the reduction runs over the ROWS of both
// operands, so neither is K-major in memory.  TMA lands 16 rows of dZ and of X as they are ([16][256], no swizzle); the worker warps
// transpose them while they split them: thread t reads column t of the 16 rows and writes row t of the K-major, 64B-swizzled hi / lo tiles
// (the swizzle makes those 64-byte row writes conflict-free).  The batch is cut into slices of 128 rows, one CTA per slice and branch; a CTA
// holds the whole [N_out <= 256, K_in <= 256] partial gradient in TMEM (two 128-lane accumulators = all 512 columns) and writes it with TMA
// to partial[slice]; wgrad_finish_kernel adds the slices up in slice order into the gradient: deterministic, no atomics.
// ================================================================================================================================
constexpr int WG_ROWS = 128;                         // rows of the batch per CTA
constexpr int WG_BK = 16;                            // rows per k-block
constexpr int WG_RAW = WG_BK * 256 * 4;              // one raw tile: 16 KB
constexpr int WG_OP = 256 * WG_BK * 4;               // one operand plane ([256 rows][16] K-major): 16 KB
constexpr int WG_STAGES = 2;
constexpr int WG_STAGE_BYTES = 2 * WG_RAW + 4 * WG_OP;          // raw dZ, raw X, A hi, A lo, B hi, B lo = 96 KB
constexpr int WG_SMEM = WG_STAGES * WG_STAGE_BYTES + 1024 + 128;

struct WgradParams {
  int M, N_out, K_in;
};

__device__ __forceinline__ void tma_store_3d(const CUtensorMap *map, const void *src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t saddr) {      // K-major, 64-byte rows, 8-row groups of 512 bytes
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | (32ull << 32) | (1ull << 46) | (4ull << 61);
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_dz0, const __grid_constant__ CUtensorMap map_x0, const __grid_constant__ CUtensorMap map_dz1,
                const __grid_constant__ CUtensorMap map_x1, const __grid_constant__ CUtensorMap map_p0, const __grid_constant__ CUtensorMap map_p1,
                const WgradParams P) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char *smem = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t *bars = (uint64_t *)(smem + WG_STAGES * WG_STAGE_BYTES);
  uint64_t *full = bars, *ready = bars + WG_STAGES, *empty = bars + 2 * WG_STAGES, *acc = bars + 3 * WG_STAGES;
  uint32_t *tmem_slot = (uint32_t *)(bars + 3 * WG_STAGES + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slice = blockIdx.x, m0 = slice * WG_ROWS;
  const CUtensorMap *map_dz = blockIdx.y ? &map_dz1 : &map_dz0, *map_x = blockIdx.y ? &map_x1 : &map_x0, *map_p = blockIdx.y ? &map_p1 : &map_p0;
  const int num_kb = WG_ROWS / WG_BK;
  const int halves = P.N_out > 128 ? 2 : 1;
  // column split (gridDim.z = 2): this CTA owns the gradient's columns [k0, k0 + kw): only that half of X is transposed and multiplied
  const int kw_cap = 256 / (int)gridDim.z, k0 = (int)blockIdx.z * kw_cap, kw = min(kw_cap, P.K_in - k0);
  const int n_mma = (kw + 15) & ~15;                 // MMA N: the gradient's columns of this CTA

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map_dz) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(map_x) : "memory");
    for (int s = 0; s < WG_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&ready[s], NUM_WORKERS); mbar_init(&empty[s], 1); }
    mbar_init(acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {                                   // all of TMEM: two 128-lane x 256-column fp32 accumulators
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; kb++) {
        const int s = kb % WG_STAGES;
        mbar_wait(&empty[s], ((kb / WG_STAGES) & 1) ^ 1);
        unsigned char *st = smem + s * WG_STAGE_BYTES;
        mbar_expect_tx(&full[s], WG_RAW + WG_RAW / (int)gridDim.z);
        tma_load_2d(st, map_dz, &full[s], 0, m0 + kb * WG_BK);                   // [16 rows][256 columns of dZ], out-of-range = 0
        tma_load_2d(st + WG_RAW, map_x, &full[s], k0, m0 + kb * WG_BK);          // [16 rows][kw_cap columns of X from k0]
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_tf32(128, n_mma);
      for (int kb = 0; kb < num_kb; kb++) {
        const int s = kb % WG_STAGES;
        mbar_wait(&ready[s], (kb / WG_STAGES) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_hi = smem_u32(smem + s * WG_STAGE_BYTES + 2 * WG_RAW), a_lo = a_hi + WG_OP, b_hi = a_lo + WG_OP, b_lo = b_hi + WG_OP;
        for (int h = 0; h < halves; h++) {
          const uint32_t ah = a_hi + h * (WG_OP / 2), al = a_lo + h * (WG_OP / 2), d = tmem_base + h * kw_cap;
#pragma unroll
          for (int k = 0; k < WG_BK / UMMA_K; k++) {
            const uint32_t ko = k * UMMA_K * 4;
            umma_tf32(d, umma_desc_sw64(ah + ko), umma_desc_sw64(b_lo + ko), idesc, (kb | k) ? 1u : 0u);
            umma_tf32(d, umma_desc_sw64(al + ko), umma_desc_sw64(b_hi + ko), idesc, 1u);
            umma_tf32(d, umma_desc_sw64(ah + ko), umma_desc_sw64(b_hi + ko), idesc, 1u);
          }
        }
        umma_commit(&empty[s]);
      }
      umma_commit(acc);
    }
  } else {
    // ===== transposing transform: thread t owns column t of both raw tiles =====
    const int t = threadIdx.x - 64;                    // 0..255
    const uint32_t row_off = (uint32_t)(t >> 3) * 512u + (uint32_t)(t & 7) * 64u, sw = (uint32_t)(t >> 1) & 3u;
    for (int kb = 0; kb < num_kb; kb++) {
      const int s = kb % WG_STAGES;
      mbar_wait(&full[s], (kb / WG_STAGES) & 1);
      unsigned char *st = smem + s * WG_STAGE_BYTES;
#pragma unroll
      for (int op = 0; op < 2; op++) {
        if (op == 1 && t >= kw_cap) continue;          // X: only this CTA's columns (raw tile rows are kw_cap floats wide)
        const float *raw = reinterpret_cast<const float *>(st + op * WG_RAW);
        unsigned char *hi = st + 2 * WG_RAW + op * 2 * WG_OP, *lo = hi + WG_OP;
        const int rw = op ? kw_cap : 256;
        float v[WG_BK];
#pragma unroll
        for (int m = 0; m < WG_BK; m++) v[m] = raw[m * rw + t];
#pragma unroll
        for (int q = 0; q < WG_BK / 4; q++) {
          float4 h4, l4;
          h4.x = __uint_as_float(__float_as_uint(v[4 * q]) & 0xFFFFE000u); h4.y = __uint_as_float(__float_as_uint(v[4 * q + 1]) & 0xFFFFE000u);
          h4.z = __uint_as_float(__float_as_uint(v[4 * q + 2]) & 0xFFFFE000u); h4.w = __uint_as_float(__float_as_uint(v[4 * q + 3]) & 0xFFFFE000u);
          l4.x = v[4 * q] - h4.x; l4.y = v[4 * q + 1] - h4.y; l4.z = v[4 * q + 2] - h4.z; l4.w = v[4 * q + 3] - h4.w;
          const uint32_t off = row_off + ((((uint32_t)q) ^ sw) << 4);
          *reinterpret_cast<float4 *>(hi + off) = h4;
          *reinterpret_cast<float4 *>(lo + off) = l4;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(&ready[s]);
    }
    // ===== epilogue: partial[slice][n][k] out of TMEM through swizzled staging tiles and TMA =====
    mbar_wait(acc, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int quad = warp & 3, half = (warp - 2) >> 2;
    unsigned char *stage_w = smem + (size_t)(warp - 2) * (2 * 4096);
    const uint32_t sw7 = (uint32_t)(lane & 7);
    uint32_t r[32];
    int it = 0;
    for (int h = 0; h < halves; h++) {
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + h * kw_cap;
      for (int c0 = half * 32; c0 < kw; c0 += 64, it++) {
        tmem_ld32(taddr + c0, r);
        unsigned char *ob = stage_w + (size_t)(it & 1) * 4096;
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 8; q++)
          *reinterpret_cast<float4 *>(ob + (uint32_t)lane * 128u + ((((uint32_t)q) ^ sw7) << 4)) =
              make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(map_p, ob, k0 + c0, h * 128 + quad * 32, slice);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

// grad[i] += sum over the slices, in slice order (grad is W.grad: the optimizer's flat gradient bucket)
__global__ void wgrad_finish_kernel(const float4 *__restrict__ part0, const float4 *__restrict__ part1, float4 *__restrict__ g0, float4 *__restrict__ g1, int n4,
                                    int slices) {
  const float4 *part = blockIdx.y ? part1 : part0;
  float4 *g = blockIdx.y ? g1 : g0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
    float4 a = g[i];
#pragma unroll 8
    for (int s = 0; s < slices; s++) { const float4 v = part[(size_t)s * n4 + i]; a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; }
    g[i] = a;
  }
}

// ---- the weights' two TF32 planes, refreshed once per optimizer step instead of once per tile and CTA ----
__global__ void tf32_split_kernel(const float *__restrict__ src, long long n, float *__restrict__ hi, float *__restrict__ lo) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = src[i], h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    hi[i] = h; lo[i] = v - h;
  }
}
// table[e] = {src offset, rows, cols, dst offset}: dst[c][r] = src[r][c] for every 2-D weight, split on the way (the input-gradient product
// reduces over the Linear's output features, so its K-major B operand is W^T)
__global__ void tf32_split_transpose_kernel(const float *__restrict__ src, const long long *__restrict__ table, float *__restrict__ t_hi, float *__restrict__ t_lo) {
  __shared__ float tile[32][33];
  const long long so = table[4 * blockIdx.y], rows = table[4 * blockIdx.y + 1], cols = table[4 * blockIdx.y + 2], dof = table[4 * blockIdx.y + 3];
  const long long tr = (rows + 31) / 32, tc = (cols + 31) / 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
  for (long long tid = blockIdx.x; tid < tr * tc; tid += gridDim.x) {
    const long long r0 = (tid / tc) * 32, c0 = (tid % tc) * 32;
    for (int k = ty; k < 32; k += 8) tile[k][tx] = (r0 + k < rows && c0 + tx < cols) ? src[so + (r0 + k) * cols + c0 + tx] : 0.f;
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
      if (c0 + k < cols && r0 + tx < rows) {
        const float v = tile[tx][k], h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        t_hi[dof + (c0 + k) * rows + r0 + tx] = h; t_lo[dof + (c0 + k) * rows + r0 + tx] = v - h;
      }
    }
    __syncthreads();
  }
}

// ---- host: TMA descriptors through the driver entry point (no -lcuda link dependency) ----
typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                    const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
encode_tiled_fn get_encode() {
  static encode_tiled_fn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
    fn = (encode_tiled_fn)p;
  }
  return fn;
}
// 2-D fp32 tensor [rows, cols], row stride ld floats; box = [BLOCK_K cols, box_rows rows]; 128-byte swizzle; out-of-bounds reads give 0
int make_map(CUtensorMap *map, const float *base, long long rows, long long cols, long long ld, int box_rows, int box_cols = BLOCK_K) {
  encode_tiled_fn enc = get_encode();
  if (!enc) FAIL(SHADOW_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) FAIL(SHADOW_EINVAL, "cuTensorMapEncodeTiled failed (%d): base %p rows %lld cols %lld ld %lld", (int)r, (const void *)base, rows, cols, ld);
  return 0;
}

}  // namespace

extern "C" int shadow_linear_tc_f32(const shadow_linear_branch *br, int32_t nbranch, int64_t ldx, int64_t ldw, int64_t ldz, int64_t ldo, int32_t M,
                                    int32_t N, int32_t K, int32_t act, int32_t do_norm, int32_t out_mode, void *cuda_stream) {
  if (M <= 0) return 0;
  if (!br || nbranch < 1 || nbranch > 2) FAIL(SHADOW_EINVAL, "linear_tc: 1 or 2 branches");
  if (N < 8 || N > BLOCK_N || (N & 3) || K < 1) FAIL(SHADOW_EINVAL, "linear_tc: N must be a multiple of 4 in [8, %d] (got %d), K >= 1", BLOCK_N, N);
  if ((ldx & 3) || (ldw & 3) || (ldz & 3) || (ldo & 3)) FAIL(SHADOW_EINVAL, "linear_tc: leading dimensions must be multiples of 4 floats");
  if (out_mode < 0 || out_mode > 2) FAIL(SHADOW_EINVAL, "linear_tc: out_mode");
  CUtensorMap mx[2], mw[2], mz[2], mo[2], mwl[2];
  const bool presplit = br[0].W_lo != nullptr;
  // two column halves per row tile when no row statistics are needed and the doubled grid still fits ONE wave (one CTA per SM: a 4,832-row
  // pair launch is 76 CTAs, doubled it would be 152 > 148 and the four stragglers double the kernel's duration -- measured)
  const int row_tiles = (M + BLOCK_M - 1) / BLOCK_M;
  static const bool no_split = getenv("SHADOW_LTC_NOSPLIT") != nullptr;
  static int sms = 0;
  if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  const int nsplit = (!do_norm && N > 128 && !no_split && row_tiles * nbranch * 2 <= sms) ? 2 : 1;
  const int bn = BLOCK_N / nsplit;
  LinearParams P;
  for (int b = 0; b < 2; b++) {
    const shadow_linear_branch &s = br[b < nbranch ? b : 0];
    if (!s.X || !s.W || !s.out) FAIL(SHADOW_EINVAL, "linear_tc: X / W / out is NULL");
    if (((uintptr_t)s.X & 15) || ((uintptr_t)s.W & 15) || ((uintptr_t)s.out & 15) || ((uintptr_t)s.Z & 15)) FAIL(SHADOW_EINVAL, "linear_tc: X / W / Z / out must be 16-byte aligned");
    if (do_norm && (!s.scale || !s.offset)) FAIL(SHADOW_EINVAL, "linear_tc: norm_feat needs scale and offset");
    int rc = make_map(&mx[b], s.X, M, K, ldx, BLOCK_M);
    if (rc) return rc;
    rc = make_map(&mw[b], s.W, N, K, ldw, bn);
    if (rc) return rc;
    if ((s.W_lo != nullptr) != presplit || ((uintptr_t)s.W_lo & 15)) FAIL(SHADOW_EINVAL, "linear_tc: W_lo must be given for all branches or none, 16-byte aligned");
    rc = make_map(&mwl[b], presplit ? s.W_lo : s.W, N, K, ldw, bn);
    if (rc) return rc;
    rc = make_map(&mo[b], s.out, M, N, ldo, 32, 32);                           // epilogue tiles: 32 rows x 32 columns, 128-byte swizzle
    if (rc) return rc;
    rc = make_map(&mz[b], s.Z ? s.Z : s.out, M, N, s.Z ? ldz : ldo, 32, 32);
    if (rc) return rc;
    P.br[b].bias = s.bias; P.br[b].scale = s.scale; P.br[b].offset = s.offset; P.br[b].Z = s.Z; P.br[b].out = s.out;
    P.br[b].mean = do_norm ? s.mean : nullptr; P.br[b].rstd = do_norm ? s.rstd : nullptr;
  }
  P.ldz = (int)ldz; P.ldo = (int)ldo; P.M = M; P.N = N; P.K = K; P.act = act; P.do_norm = do_norm; P.out_mode = out_mode; P.presplit = presplit ? 1 : 0;
  { const char *e = getenv("SHADOW_LTC_DEBUG"); P.debug = e ? atoi(e) : 0; }
  typedef void (*kern_t)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap,
                         const CUtensorMap, const CUtensorMap, const CUtensorMap, const LinearParams);
  static const kern_t kerns[5][2] = {{linear_tc_kernel<0, false>, linear_tc_kernel<0, true>}, {linear_tc_kernel<1, false>, linear_tc_kernel<1, true>},
                                     {linear_tc_kernel<2, false>, linear_tc_kernel<2, true>}, {linear_tc_kernel<3, false>, linear_tc_kernel<3, true>},
                                     {linear_tc_kernel<4, false>, linear_tc_kernel<4, true>}};
  if (act < 0 || act > 4) FAIL(SHADOW_EINVAL, "linear_tc: unknown activation id %d", act);
  static bool attr_set[5][2] = {};
  const kern_t kern = kerns[act][do_norm ? 1 : 0];
  if (!attr_set[act][do_norm ? 1 : 0]) { CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES)); attr_set[act][do_norm ? 1 : 0] = true; }
  kern<<<dim3(row_tiles, nbranch, nsplit), NUM_THREADS, SMEM_BYTES, (cudaStream_t)cuda_stream>>>(mx[0], mw[0], mx[1], mw[1], mz[0], mo[0], mz[1], mo[1], mwl[0], mwl[1], P);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

extern "C" int shadow_tf32_split_f32(const float *src, int64_t n, float *hi, float *lo, void *cuda_stream) {
  if (n <= 0) return 0;
  if (!src || !hi || !lo) FAIL(SHADOW_EINVAL, "tf32_split: NULL argument");
  tf32_split_kernel<<<(unsigned)std::min<long long>((n + 255) / 256, 1184), 256, 0, (cudaStream_t)cuda_stream>>>(src, n, hi, lo);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
extern "C" int shadow_tf32_split_transpose_f32(const float *src, const int64_t *table_dev, int32_t num_entries, float *t_hi, float *t_lo, void *cuda_stream) {
  if (num_entries <= 0) return 0;
  if (!src || !table_dev || !t_hi || !t_lo) FAIL(SHADOW_EINVAL, "tf32_split_transpose: NULL argument");
  tf32_split_transpose_kernel<<<dim3(64, num_entries), 256, 0, (cudaStream_t)cuda_stream>>>(src, (const long long *)table_dev, t_hi, t_lo);
  CUDA_TRY(cudaGetLastError());
  return 0;
}

namespace {
int make_map_raw(CUtensorMap *map, const float *base, long long rows, long long cols, long long ld, int box_cols = 256) {      // box = [box_cols columns, 16 rows], no swizzle
  encode_tiled_fn enc = get_encode();
  if (!enc) FAIL(SHADOW_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)WG_BK};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) FAIL(SHADOW_EINVAL, "cuTensorMapEncodeTiled (raw) failed (%d)", (int)r);
  return 0;
}
int make_map_partial(CUtensorMap *map, const float *base, long long slices, long long n_out, long long k_in) {     // [slices][n_out][k_in], box 32 x 32 x 1, 128-byte swizzle
  encode_tiled_fn enc = get_encode();
  if (!enc) FAIL(SHADOW_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[3] = {(cuuint64_t)k_in, (cuuint64_t)n_out, (cuuint64_t)slices};
  cuuint64_t strides[2] = {(cuuint64_t)k_in * 4, (cuuint64_t)k_in * n_out * 4};
  cuuint32_t box[3] = {32, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) FAIL(SHADOW_EINVAL, "cuTensorMapEncodeTiled (partial) failed (%d)", (int)r);
  return 0;
}
}  // namespace

extern "C" int64_t shadow_wgrad_tc_scratch_floats(int32_t M, int32_t N_out, int32_t K_in) {
  return (int64_t)((M + WG_ROWS - 1) / WG_ROWS) * N_out * K_in;
}
extern "C" int shadow_wgrad_tc_f32(const float *dZ0, const float *X0, float *grad0, float *scratch0, const float *dZ1, const float *X1, float *grad1,
                                   float *scratch1, int32_t M, int32_t N_out, int32_t K_in, void *cuda_stream) {
  if (M <= 0) return 0;
  const int nb = dZ1 ? 2 : 1;
  if (N_out < 8 || N_out > 256 || K_in < 8 || K_in > 256 || (N_out & 3) || (K_in & 3)) FAIL(SHADOW_EINVAL, "wgrad_tc: N_out / K_in must be multiples of 4 in [8, 256]");
  const float *dZ[2] = {dZ0, dZ1 ? dZ1 : dZ0}, *X[2] = {X0, X1 ? X1 : X0};
  float *grad[2] = {grad0, grad1 ? grad1 : grad0}, *scr[2] = {scratch0, scratch1 ? scratch1 : scratch0};
  const int slices = (M + WG_ROWS - 1) / WG_ROWS;
  static const bool no_split = getenv("SHADOW_LTC_NOSPLIT") != nullptr;
  static int sms = 0;
  if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  const int ksplit = (K_in > 128 && !no_split && slices * nb * 2 <= sms) ? 2 : 1;      // one wave only (see shadow_linear_tc_f32)
  CUtensorMap mdz[2], mx[2], mp[2];
  for (int b = 0; b < 2; b++) {
    if (!dZ[b] || !X[b] || !grad[b] || !scr[b]) FAIL(SHADOW_EINVAL, "wgrad_tc: NULL argument");
    if (((uintptr_t)dZ[b] | (uintptr_t)X[b] | (uintptr_t)grad[b] | (uintptr_t)scr[b]) & 15) FAIL(SHADOW_EINVAL, "wgrad_tc: pointers must be 16-byte aligned");
    int rc = make_map_raw(&mdz[b], dZ[b], M, N_out, N_out);
    if (rc) return rc;
    rc = make_map_raw(&mx[b], X[b], M, K_in, K_in, 256 / ksplit);
    if (rc) return rc;
    rc = make_map_partial(&mp[b], scr[b], slices, N_out, K_in);
    if (rc) return rc;
  }
  WgradParams P;
  P.M = M; P.N_out = N_out; P.K_in = K_in;
  static bool attr_set = false;
  if (!attr_set) { CUDA_TRY(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM)); attr_set = true; }
  wgrad_tc_kernel<<<dim3(slices, nb, ksplit), NUM_THREADS, WG_SMEM, (cudaStream_t)cuda_stream>>>(mdz[0], mx[0], mdz[1], mx[1], mp[0], mp[1], P);
  CUDA_TRY(cudaGetLastError());
  const int n4 = N_out * K_in / 4;
  wgrad_finish_kernel<<<dim3((n4 + 255) / 256, nb), 256, 0, (cudaStream_t)cuda_stream>>>((const float4 *)scr[0], (const float4 *)scr[1], (float4 *)grad[0], (float4 *)grad[1],
                                                                                         n4, slices);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
