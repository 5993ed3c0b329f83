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

// ------------------------------------------------------------------------------------------------
// One training batch of a super-batch -> the static buffers of the captured step (train.GraphedTrainer._load_static): the slice of the
// canonical CSR is rebased to row 0 / edge 0, padding rows get empty adjacency rows, the gathered feature rows are copied.  One launch
// instead of eight small ones between two graph replays (they sat on the critical path of the 0.5 ms step).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) load_batch_kernel(const int *__restrict__ rowptr_src, int e0, int n, int e, int row_cap, int *__restrict__ rowptr_dst,
                                                         int2 *__restrict__ span_dst, const int *__restrict__ idx_src, int lo, int *__restrict__ col_dst,
                                                         const float *__restrict__ feat_src, int F, float *__restrict__ feat_dst, const int *__restrict__ tgt_src,
                                                         int B, long long *__restrict__ tgt_dst) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  for (long long i = tid; i <= row_cap; i += stride) {
    const int a = i <= n ? rowptr_src[i] - e0 : e;
    rowptr_dst[i] = a;
    if (i < row_cap) span_dst[i] = make_int2(a, i + 1 <= n ? rowptr_src[i + 1] - e0 : e);
  }
  for (long long j = tid; j < e; j += stride) col_dst[j] = idx_src[j] - lo;
  for (long long j = tid; j < B; j += stride) tgt_dst[j] = (long long)(tgt_src[j] - lo);
  const long long nf = (long long)n * F;
  if ((F & 3) == 0 && ((((uintptr_t)feat_src) | ((uintptr_t)feat_dst)) & 15) == 0) {
    const float4 *s4 = reinterpret_cast<const float4 *>(feat_src);
    float4 *d4 = reinterpret_cast<float4 *>(feat_dst);
    for (long long j = tid; j < nf / 4; j += stride) d4[j] = s4[j];
  } else {
    for (long long j = tid; j < nf; j += stride) feat_dst[j] = feat_src[j];
  }
}
extern "C" int shadow_load_batch(const int32_t *rowptr_src, int32_t e0, int32_t n, int32_t e, int32_t row_cap, int32_t *rowptr_dst, int32_t *span_dst,
                                 const int32_t *idx_src, int32_t lo, int32_t *col_dst, const float *feat_src, int32_t F, float *feat_dst,
                                 const int32_t *tgt_src, int32_t B, int64_t *tgt_dst, void *stream) {
  if (!rowptr_src || !rowptr_dst || !span_dst || (e && (!idx_src || !col_dst)) || !feat_src || !feat_dst || (B && (!tgt_src || !tgt_dst)) || n < 0 || n > row_cap || F <= 0)
    FAIL(SHADOW_EINVAL, "bad argument to shadow_load_batch");
  int dev = 0, sms = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long work = std::max<long long>((long long)n * F / 4, (long long)row_cap + 1);
  const int grid = (int)std::min<long long>((work + 255) / 256, (long long)sms * 4);
  load_batch_kernel<<<std::max(grid, 1), 256, 0, (cudaStream_t)stream>>>(rowptr_src, e0, n, e, row_cap, rowptr_dst, (int2 *)span_dst, idx_src, lo, col_dst, feat_src, F,
                                                                         feat_dst, tgt_src, B, (long long *)tgt_dst);
  CUDA_TRY(cudaGetLastError());
  return 0;
}
