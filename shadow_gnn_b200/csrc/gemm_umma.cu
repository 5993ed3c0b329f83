// gemm_umma.cu -- the dense Linear of the layers on the 5th-generation tensor cores: tcgen05.mma with TMEM accumulators, operands staged
// by TMA, fp32 in / fp32 out with fp32-level accuracy (every fp32 operand is split into three bf16 terms inside the kernel and the
// nine-term product is accumulated in fp32: CUTLASS' "FastFP32" / 9xBF16 mainloop for sm_100).  The kernel is assembled here from the
// CUTLASS/CuTe collective templates vendored in the image (flashinfer/data/cutlass/include, CUTLASS 4.5): TMA-load warp, input
// transform warps (fp32 -> 3 x bf16, A into TMEM), one MMA-issuing thread, epilogue warps reading the accumulator back with tcgen05.ld.
//
// Three operand layouts = the three products of nn.Linear (shaDow/layers.py:421,451-452,501-505,556,670; models.py:81):
//   variant 0  forward  Z  = X W^T (+ b)   A = X  [M,K] K-major,  B = W [N,K] K-major; the bias enters as the C operand with row stride 0
//   variant 1  dgrad    dX = dZ W          A = dZ [M,K] K-major,  B = W [K,N] N-major
//   variant 2  wgrad    dW_l = dZ_l^T X_l  A = dZ^T, M-major,     B = X [K,N] N-major, batched over L slices of the reduction (batch) dimension
// Shapes must satisfy the 16-byte TMA alignment (leading dimensions and the contiguous extent multiples of 4 floats); anything else is
// refused with SHADOW_EINVAL and the caller uses the warp-level MMA kernel of gemm.cu.
// One translation unit per variant (-DUMMA_VARIANT=0|1|2): the templates take about a minute each to compile.
#include <cuda_runtime.h>
#include <stdint.h>

#include "cutlass/cutlass.h"
#include "cute/tensor.hpp"
#include "cutlass/gemm/device/gemm_universal_adapter.h"
#include "cutlass/gemm/kernel/gemm_universal.hpp"
#include "cutlass/gemm/collective/collective_builder.hpp"
#include "cutlass/epilogue/collective/collective_builder.hpp"

#include "../../include/shadow_b200.h"

void shadow_set_error(const char *fmt, ...);

using namespace cute;

template <class LayoutA, class LayoutB>
struct FastF32Gemm {
  using ElementA = float;
  using ElementB = float;
  using ElementC = float;
  using ElementAcc = float;
  using LayoutC = cutlass::layout::RowMajor;
  static constexpr int AlignA = 4, AlignB = 4, AlignC = 4;
  using ArchTag = cutlass::arch::Sm100;
  using OpClass = cutlass::arch::OpClassTensorOp;
  using MmaTileShape = Shape<_128, _64, _16>;      // 4,832 x 256 output -> 38 x 4 = 152 tiles for 148 SMs
  using ClusterShape = Shape<_1, _1, _1>;
  using CollectiveEpilogue = typename cutlass::epilogue::collective::CollectiveBuilder<
      ArchTag, OpClass, MmaTileShape, ClusterShape, cutlass::epilogue::collective::EpilogueTileAuto, ElementAcc, ElementAcc, ElementC, LayoutC,
      AlignC, ElementC, LayoutC, AlignC, cutlass::epilogue::FastF32NoSmemWarpSpecialized1Sm>::CollectiveOp;
  using CollectiveMainloop = typename cutlass::gemm::collective::CollectiveBuilder<
      ArchTag, OpClass, ElementA, LayoutA, AlignA, ElementB, LayoutB, AlignB, ElementAcc, MmaTileShape, ClusterShape,
      cutlass::gemm::collective::StageCountAutoCarveout<static_cast<int>(sizeof(typename CollectiveEpilogue::SharedStorage))>,
      cutlass::gemm::KernelTmaWarpSpecialized1SmFastFP32Sm100>::CollectiveOp;
  using GemmKernel = cutlass::gemm::kernel::GemmUniversal<Shape<int, int, int, int>, CollectiveMainloop, CollectiveEpilogue, void>;
  using Gemm = cutlass::gemm::device::GemmUniversalAdapter<GemmKernel>;
};

// (extent0, extent1, batch) strides: exactly one of the first two modes is the static unit stride
template <class S>
static void set_stride(S &s, int64_t ld, int64_t batch) {
  if constexpr (cute::is_static<cute::remove_cvref_t<decltype(cute::get<0>(s))>>::value) cute::get<1>(s) = ld;
  else cute::get<0>(s) = ld;
  cute::get<2>(s) = batch;
}

static void *g_ws[64] = {nullptr};
static const size_t kWsBytes = 4u << 20;

template <class G>
static int run_umma(const float *A, int64_t lda, int64_t bsA, const float *B, int64_t ldb, int64_t bsB, const float *Csrc, int64_t ldcs, float *D,
                    int64_t ldd, int64_t bsD, int M, int N, int K, int L, cudaStream_t st) {
  using Gemm = typename G::Gemm;
  typename Gemm::GemmKernel::StrideA sA;
  typename Gemm::GemmKernel::StrideB sB;
  typename Gemm::GemmKernel::StrideC sC;
  typename Gemm::GemmKernel::StrideD sD;
  set_stride(sA, lda, bsA);
  set_stride(sB, ldb, bsB);
  set_stride(sC, ldcs, 0);
  set_stride(sD, ldd, bsD);
  const float beta = Csrc ? 1.f : 0.f;
  typename Gemm::Arguments args{cutlass::gemm::GemmUniversalMode::kGemm, {M, N, K, L}, {A, sA, B, sB}, {{1.0f, beta}, Csrc ? Csrc : D, sC, D, sD}};
  if (!Csrc) set_stride(args.epilogue.dC, ldd, bsD);
  Gemm gemm;
  if (gemm.can_implement(args) != cutlass::Status::kSuccess) { shadow_set_error("umma linear: shape / alignment not supported by the TMA path"); return SHADOW_EINVAL; }
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { shadow_set_error("umma linear: cudaGetDevice failed"); return SHADOW_ECUDA; }
  if (Gemm::get_workspace_size(args) > kWsBytes) { shadow_set_error("umma linear: workspace too small"); return SHADOW_ECAP; }
  if (!g_ws[dev] && cudaMalloc(&g_ws[dev], kWsBytes) != cudaSuccess) { shadow_set_error("umma linear: cudaMalloc(workspace) failed"); return SHADOW_ECUDA; }
  if (gemm.initialize(args, g_ws[dev], st) != cutlass::Status::kSuccess) { shadow_set_error("umma linear: initialize failed"); return SHADOW_ECUDA; }
  if (gemm.run(st) != cutlass::Status::kSuccess) { shadow_set_error("umma linear: launch failed: %s", cudaGetErrorString(cudaGetLastError())); return SHADOW_ECUDA; }
  return 0;
}

static bool al4(int64_t x) { return (x & 3) == 0; }
static bool al16(const void *p) { return ((uintptr_t)p & 15) == 0; }

#if UMMA_VARIANT == 0
extern "C" int shadow_linear_umma_fwd_f32(const float *X, int64_t ldx, const float *W, int64_t ldw, const float *bias, float *Z, int64_t ldz, int32_t M,
                                          int32_t N, int32_t K, void *cuda_stream) {
  if (!X || !W || !Z || M <= 0 || N <= 0 || K <= 0 || !al4(ldx) || !al4(ldw) || !al4(ldz) || !al4(K) || !al4(N) || !al16(X) || !al16(W) || !al16(Z) ||
      (bias && !al16(bias))) {
    shadow_set_error("umma linear forward: needs 16-byte aligned operands and K, N, leading dimensions multiples of 4");
    return SHADOW_EINVAL;
  }
  return run_umma<FastF32Gemm<cutlass::layout::RowMajor, cutlass::layout::ColumnMajor>>(X, ldx, 0, W, ldw, 0, bias, 0, Z, ldz, 0, M, N, K, 1,
                                                                                         (cudaStream_t)cuda_stream);
}
#elif UMMA_VARIANT == 1
extern "C" int shadow_linear_umma_dgrad_f32(const float *dZ, int64_t lddz, const float *W, int64_t ldw, float *dX, int64_t lddx, int32_t M, int32_t N,
                                            int32_t K, void *cuda_stream) {
  // dX[M,N] = dZ[M,K] W[K,N]
  if (!dZ || !W || !dX || M <= 0 || N <= 0 || K <= 0 || !al4(lddz) || !al4(ldw) || !al4(lddx) || !al4(K) || !al4(N) || !al16(dZ) || !al16(W) || !al16(dX)) {
    shadow_set_error("umma linear dgrad: needs 16-byte aligned operands and K, N, leading dimensions multiples of 4");
    return SHADOW_EINVAL;
  }
  return run_umma<FastF32Gemm<cutlass::layout::RowMajor, cutlass::layout::RowMajor>>(dZ, lddz, 0, W, ldw, 0, nullptr, 0, dX, lddx, 0, M, N, K, 1,
                                                                                      (cudaStream_t)cuda_stream);
}
#else
extern "C" int shadow_linear_umma_wgrad_f32(const float *dZ, int64_t lddz, const float *X, int64_t ldx, float *part, int32_t rows_per_slice,
                                            int32_t slices, int32_t N_out, int32_t K_in, void *cuda_stream) {
  // part[l][N_out,K_in] = dZ[l*r:(l+1)*r, :N_out]^T  X[l*r:(l+1)*r, :K_in]
  if (!dZ || !X || !part || rows_per_slice <= 0 || slices <= 0 || N_out <= 0 || K_in <= 0 || !al4(lddz) || !al4(ldx) || !al4(N_out) || !al4(K_in) ||
      !al16(dZ) || !al16(X) || !al16(part)) {
    shadow_set_error("umma linear wgrad: needs 16-byte aligned operands and N_out, K_in, leading dimensions multiples of 4");
    return SHADOW_EINVAL;
  }
  return run_umma<FastF32Gemm<cutlass::layout::ColumnMajor, cutlass::layout::RowMajor>>(
      dZ, lddz, (int64_t)rows_per_slice * lddz, X, ldx, (int64_t)rows_per_slice * ldx, nullptr, 0, part, K_in, (int64_t)N_out * K_in, N_out, K_in,
      rows_per_slice, slices, (cudaStream_t)cuda_stream);
}
#endif
