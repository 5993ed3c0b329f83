// ubench_bulk.cu -- how fast can one warp per CTA pull many SMALL rows (64 B .. 4 KB, 16-byte aligned, random places of a 512 MB array)
// into shared memory with cp.async.bulk (1-D TMA, one copy per lane per round)?  Decides the staging design of ppr_induce_warp_kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench_bulk scripts/ubench_bulk.cu && /tmp/ubench_bulk
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWL:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DN;\nbra WL;\nDN:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
               "r"(smem_u32(bar))
               : "memory");
}

// every round: `lanes` lanes copy `sz` bytes each; 2 stages in flight
__global__ void __launch_bounds__(32) bulk_kernel(const uint4 *src, size_t n16, int sz, int lanes, int rounds, unsigned long long *sink) {
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t *bar = (uint64_t *)sm;
  unsigned char *buf = sm + 128;
  const int lane = threadIdx.x;
  const int stage_bytes = sz * lanes;
  if (lane == 0) { mbar_init(bar, 1); mbar_init(bar + 1, 1); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  uint32_t rng = (blockIdx.x * 32 + lane) * 2654435761u + 12345u;
  auto issue = [&](int st) {
    if (lane == 0) mbar_expect_tx(bar + st, (uint32_t)stage_bytes);
    __syncwarp();
    if (lane < lanes) {
      rng = rng * 1664525u + 1013904223u;
      const size_t o = ((size_t)rng * 977u) % (n16 - 512);
      bulk_g2s(buf + st * stage_bytes + lane * sz, src + o, (uint32_t)sz, bar + st);
    }
  };
  unsigned long long acc = 0;
  issue(0);
  for (int r = 0; r < rounds; r++) {
    const int st = r & 1;
    if (r + 1 < rounds) issue(st ^ 1);
    mbar_wait(bar + st, (r >> 1) & 1);
    const uint4 *b4 = (const uint4 *)(buf + st * stage_bytes);
    for (int i = lane; i < stage_bytes / 16; i += 32) { const uint4 v = b4[i]; acc += v.x ^ v.y ^ v.z ^ v.w; }
    __syncwarp();
  }
  if (acc == 0x1234567ull) sink[0] = acc;
}

// the same traffic with plain 128-bit loads: lane l of round-chunk reads consecutive 16 B of row (i / (sz/16))
__global__ void __launch_bounds__(32) ldg_kernel(const uint4 *src, size_t n16, int sz, int lanes, int rounds, unsigned long long *sink) {
  const int lane = threadIdx.x;
  uint32_t rng0 = (blockIdx.x * 32) * 2654435761u + 12345u;
  unsigned long long acc = 0;
  const int per = sz / 16, total = per * lanes;
  for (int r = 0; r < rounds; r++) {
    uint4 v[8];
    int cnt = 0;
    for (int i = lane; i < total; i += 32) {
      const int row = i / per;
      uint32_t rr = (rng0 + row * 2654435761u) * 1664525u + 1013904223u + r * 97u;
      const size_t o = ((size_t)rr * 977u) % (n16 - 512);
      v[cnt & 7] = __ldg(src + o + (i - row * per));
      cnt++;
      if ((cnt & 7) == 0) for (int j = 0; j < 8; j++) acc += v[j].x ^ v[j].y ^ v[j].z ^ v[j].w;
    }
    for (int j = 0; j < (cnt & 7); j++) acc += v[j].x ^ v[j].y ^ v[j].z ^ v[j].w;
  }
  if (acc == 0x1234567ull) sink[0] = acc;
}

int main() {
  const size_t bytes = 512ull << 20, n16 = bytes / 16;
  uint4 *src; unsigned long long *sink;
  cudaMalloc(&src, bytes); cudaMalloc(&sink, 8);
  cudaMemset(src, 1, bytes);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int szs[] = {64, 128, 256, 512, 1024, 2048};
  const int wps[] = {8, 12, 16, 24};
  for (int use_ldg = 0; use_ldg < 2; use_ldg++)
    for (int w : wps)
      for (int sz : szs) {
        const int lanes = sz <= 512 ? 32 : (sz == 1024 ? 8 : 4);       // stage <= 16 KB
        const int stage = sz * lanes;
        const int smem = 128 + 2 * stage;
        if ((size_t)smem * w > 220 * 1024) continue;
        const int rounds = (int)((8u << 20) / stage / 8) + 8;
        cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        const int grid = 148 * w;
        float best = 1e9f;
        for (int it = 0; it < 4; it++) {
          cudaEventRecord(e0);
          if (use_ldg) ldg_kernel<<<grid, 32, smem>>>(src, n16, sz, lanes, rounds, sink);
          else bulk_kernel<<<grid, 32, smem>>>(src, n16, sz, lanes, rounds, sink);
          cudaEventRecord(e1); cudaEventSynchronize(e1);
          float ms; cudaEventElapsedTime(&ms, e0, e1);
          if (it && ms < best) best = ms;
        }
        const double gb = (double)grid * rounds * stage / 1e9;
        printf("%s warps/SM %2d  row %4d B x %2d lanes  stage %5d B: %7.1f GB/s  %6.2f M rows/s/SM  (%s)\n", use_ldg ? "ldg " : "bulk", w, sz, lanes, stage,
               gb / (best * 1e-3), (double)grid * rounds * lanes / (best * 1e-3) / 148 / 1e6, cudaGetErrorString(cudaGetLastError()));
      }
  return 0;
}
