// gather.cu -- feature-row gather: out[i,:] = feat[ids[i],:]   (feat_full[subgs.node], shaDow/minibatch.py:469)
// Pure HBM-bandwidth kernel: 128-bit streaming loads/stores when the row stride allows it.
#include <algorithm>
#include "common.cuh"

__global__ void __launch_bounds__(256) gather_rows_vec4_kernel(const float4 *__restrict__ feat, int dim4, const uint32_t *__restrict__ ids,
                                                               long long n, float4 *__restrict__ out) {
  const long long total = n * dim4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  // 4 independent 16-byte requests in flight per thread
  for (; e + 3 * stride < total; e += 4 * stride) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const long long x = e + u * stride;
      const long long row = x / dim4;
      const int col = (int)(x - row * dim4);
      v[u] = ldg_stream_f4(feat + (long long)__ldg(ids + row) * dim4 + col);
    }
#pragma unroll
    for (int u = 0; u < 4; u++) stg_stream_f4(out + e + u * stride, v[u]);
  }
  for (; e < total; e += stride) {
    const long long row = e / dim4;
    const int col = (int)(e - row * dim4);
    stg_stream_f4(out + e, ldg_stream_f4(feat + (long long)__ldg(ids + row) * dim4 + col));
  }
}

__global__ void __launch_bounds__(256) gather_rows_scalar_kernel(const float *__restrict__ feat, int dim, const uint32_t *__restrict__ ids,
                                                                 long long n, float *__restrict__ out) {
  const long long total = n * dim;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long row = e / dim;
    out[e] = feat[(long long)ids[row] * dim + (e - row * dim)];
  }
}

extern "C" int shadow_gather_rows_f32(const float *feat, int64_t num_rows, int32_t dim, const uint32_t *ids, int64_t n, float *out,
                                      void *stream) {
  (void)num_rows;
  if (!feat || !out || (!ids && n) || dim <= 0) FAIL(SHADOW_EINVAL, "bad argument to shadow_gather_rows_f32");
  if (n == 0) return 0;
  int dev = 0, sms = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const bool vec = (dim % 4 == 0) && (((uintptr_t)feat | (uintptr_t)out) % 16 == 0);
  const long long total = vec ? n * (dim / 4) : n * (long long)dim;
  const int grid = (int)std::min<long long>((total + 255) / 256, (long long)sms * 8);
  if (vec) gather_rows_vec4_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float4 *)feat, dim / 4, ids, n, (float4 *)out);
  else gather_rows_scalar_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(feat, dim, ids, n, out);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
