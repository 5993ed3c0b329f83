// Shared device/host helpers for libshadow_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/shadow_b200.h"

#define NONE32 0xFFFFFFFFu

// ------------------------------------------------------------------------------------------------
// error plumbing: C ABI returns codes, message kept per thread
// ------------------------------------------------------------------------------------------------
void shadow_set_error(const char *fmt, ...);

#define CUDA_TRY(expr)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      shadow_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return SHADOW_ECUDA;                                                                     \
    }                                                                                          \
  } while (0)

#define FAIL(code, ...)               \
  do {                                \
    shadow_set_error(__VA_ARGS__);    \
    return (code);                    \
  } while (0)

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }
__device__ __forceinline__ uint32_t lanemask_lt() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// streaming loads for data that is read once per kernel (graph rows, feature rows): bypass L1 allocation
__device__ __forceinline__ uint32_t ldg_stream_u32(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ldg_stream_f4(const float4 *p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg_stream_f4(float4 *p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// block-wide exclusive scan of a[0..n) in place (shared or global memory); returns the total in every thread.
// `carry` is a 2-word shared scratch; blockDim.x <= 1024 and a multiple of 32.
__device__ inline uint32_t block_exclusive_scan(uint32_t *a, int n, uint32_t *warp_sums /* [33] shared */) {
  const int lane = lane_id(), w = warp_id(), nw = blockDim.x >> 5;
  uint32_t carry = 0;
  for (int base = 0; base < n; base += blockDim.x) {
    int i = base + threadIdx.x;
    uint32_t v = (i < n) ? a[i] : 0u, x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
      if (lane >= d) x += y;
    }
    if (lane == 31) warp_sums[w] = x;
    __syncthreads();
    if (w == 0) {
      uint32_t s = (lane < nw) ? warp_sums[lane] : 0u, t = s;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, t, d);
        if (lane >= d) t += y;
      }
      if (lane < nw) warp_sums[lane] = t - s;   // exclusive warp offsets
      if (lane == 31) warp_sums[32] = t;        // chunk total
    }
    __syncthreads();
    if (i < n) a[i] = carry + warp_sums[w] + x - v;
    carry += warp_sums[32];
    __syncthreads();
  }
  return carry;
}

// in-place ascending bitonic sort of a[0..np2) (np2 a power of two, padded by the caller with max keys)
template <typename T>
__device__ inline void block_bitonic_sort(T *a, int np2) {
  for (int k = 2; k <= np2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < np2; i += blockDim.x) {
        int ixj = i ^ j;
        if (ixj > i) {
          T x = a[i], y = a[ixj];
          bool asc = (i & k) == 0;
          if ((x > y) == asc) { a[i] = y; a[ixj] = x; }
        }
      }
      __syncthreads();
    }
  }
}

__host__ __device__ inline int next_pow2(int x) {
  int p = 1;
  while (p < x) p <<= 1;
  return p;
}

// Philox4x32-10 (Salmon et al. 2011), the counter-based generator of the production RNG mode
__host__ __device__ inline uint32_t philox_mulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
__host__ __device__ inline uint32_t philox4x32_10_x(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t hi0 = philox_mulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = philox_mulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return c0;
}
