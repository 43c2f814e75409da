// tc_path.cu — bf16 tensor-core path (tcgen05 / TMEM / bulk TMA).  See DESIGN.md §"bf16 path".
#include "tc_path.cuh"
#include "umma.cuh"

namespace gt {
using namespace umma;

// =============================================================================================
// Stand-alone tile GEMM: D[M,N] = A[M,K] B[N,K]^T.  One CTA per 128 rows; the whole B and the CTA's
// A tile are staged into the canonical K-major layout by the threads themselves, one elected thread
// issues K/16 UMMAs, completion arrives on an mbarrier, every warp drains its 32 TMEM lanes.
// Used by tests/test_tc_engine.py to validate descriptors, TMEM addressing and the epilogue mapping.
// variant bit0: swap LBO/SBO ; bit1: use M=64 tiles (two halves)
// =============================================================================================
__global__ void __launch_bounds__(128) tc_debug_gemm_kernel(const uint16_t *__restrict__ A, const uint16_t *__restrict__ B,
                                                           float *__restrict__ D, int M, int N, int K, int variant) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  uint8_t *sA = smem;                            // 128 x K bf16
  uint8_t *sB = smem + (size_t)128 * K * 2;      // N   x K bf16
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  const int m0 = blockIdx.x * 128;
  uint32_t ncols = 32;
  while ((int)ncols < N) ncols <<= 1;

  if (warp == 0) tmem_alloc(&tmem_slot, ncols);
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }

  // stage operands: 16-byte chunks.  bit2 / bit3: the operand is given TRANSPOSED in global memory
  // ([K, M] / [K, N] row-major) and staged with k as the row index -> consumed as an MN-major operand.
  const bool a_t = variant & 4, b_t = variant & 8;
  if (!a_t) {
    for (int kb = 0; kb < K / 8; ++kb) {
      uint4 v = *reinterpret_cast<const uint4 *>(A + (size_t)(m0 + tid) * K + kb * 8);
      *reinterpret_cast<uint4 *>(sA + kmajor_off(tid, kb * 8, 128)) = v;
    }
  } else {
    for (int k = tid; k < K; k += 128)
      for (int mb = 0; mb < 16; ++mb) {
        uint4 v = *reinterpret_cast<const uint4 *>(A + (size_t)k * M + m0 + mb * 8);
        *reinterpret_cast<uint4 *>(sA + kmajor_off(k, mb * 8, K)) = v;
      }
  }
  if (!b_t) {
    for (int r = tid; r < N; r += 128)
      for (int kb = 0; kb < K / 8; ++kb) {
        uint4 v = *reinterpret_cast<const uint4 *>(B + (size_t)r * K + kb * 8);
        *reinterpret_cast<uint4 *>(sB + kmajor_off(r, kb * 8, N)) = v;
      }
  } else {
    for (int k = tid; k < K; k += 128)
      for (int nb = 0; nb < N / 8; ++nb) {
        uint4 v = *reinterpret_cast<const uint4 *>(B + (size_t)k * N + nb * 8);
        *reinterpret_cast<uint4 *>(sB + kmajor_off(k, nb * 8, K)) = v;
      }
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (tid == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N, a_t ? 1 : 0, b_t ? 1 : 0);
    // K-major image [R rows x K]: K-adjacent core matrices (R/8)*128 B apart (LBO), row-adjacent 128 B (SBO).
    // MN-major use of an image [K rows x R]: k-adjacent cores 128 B apart (LBO), mn-adjacent (K/8)*128 B (SBO).
    uint32_t a_lbo = a_t ? 128u : (128 / 8) * 128u, a_sbo = a_t ? (uint32_t)(K / 8) * 128u : 128u;
    uint32_t b_lbo = b_t ? 128u : (uint32_t)(N / 8) * 128u, b_sbo = b_t ? (uint32_t)(K / 8) * 128u : 128u;
    if (variant & 1) { uint32_t t = a_lbo; a_lbo = a_sbo; a_sbo = t; t = b_lbo; b_lbo = b_sbo; b_sbo = t; }
    for (int k16 = 0; k16 < K / 16; ++k16) {
      uint32_t a_off = a_t ? (uint32_t)(k16 * 2) * 128u : (uint32_t)(k16 * 2) * (128 / 8) * 128u;
      uint32_t b_off = b_t ? (uint32_t)(k16 * 2) * 128u : (uint32_t)(k16 * 2) * (uint32_t)(N / 8) * 128u;
      uint64_t ad = make_desc(smem_u32(sA) + a_off, a_lbo, a_sbo);
      uint64_t bd = make_desc(smem_u32(sB) + b_off, b_lbo, b_sbo);
      mma_bf16_ss(tmem, ad, bd, idesc, k16 > 0 ? 1u : 0u);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  fence_after_sync();

  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
    tmem_ld_wait();
    float *out = D + (size_t)(m0 + warp * 32 + lane) * N + c0;
#pragma unroll
    for (int j = 0; j < 16; ++j) out[j] = v[j];
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, ncols);
}

int tc_debug_gemm(const uint16_t *a, const uint16_t *b, float *d, int m, int n, int k, int variant, cudaStream_t st) {
  GT_CHECK(a && b && d, "tc_debug_gemm: null pointer");
  GT_CHECK(m > 0 && m % 128 == 0, "tc_debug_gemm: M must be a positive multiple of 128");
  GT_CHECK(n >= 16 && n <= 256 && n % 16 == 0, "tc_debug_gemm: N must be a multiple of 16 in [16,256]");
  GT_CHECK(k >= 16 && k % 16 == 0, "tc_debug_gemm: K must be a positive multiple of 16");
  size_t smem = (size_t)(128 + n) * k * 2;
  GT_CHECK(smem <= 200 * 1024, "tc_debug_gemm: operands do not fit in shared memory");
  GT_CUDA(cudaFuncSetAttribute(tc_debug_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  tc_debug_gemm_kernel<<<m / 128, 128, smem, st>>>(a, b, d, m, n, k, variant);
  GT_CUDA(cudaGetLastError());
  return 0;
}

// =============================================================================================
// model passes — filled in by tc_layers.cu
// =============================================================================================
static const char *kNoTc = "precision=bf16 is not available for this configuration yet; use precision=fp32";

int64_t tc_workspace_bytes(const gt_config &, int64_t, int) { set_error(kNoTc); return -1; }
int tc_forward(const gt_config &, const Layout &, const float *, const float *, const float *, const float *, int64_t, float *,
               void *, int64_t, bool, uint64_t, uint64_t, int64_t, cudaStream_t) { GT_FAIL(kNoTc); }
int tc_backward(const gt_config &, const Layout &, const float *, const float *, const float *, const float *, int64_t,
                const float *, const float *, float *, void *, int64_t, uint64_t, uint64_t, int64_t, cudaStream_t) { GT_FAIL(kNoTc); }
int tc_train_step(const gt_config &, const Layout &, const float *, const float *, const float *, const float *, int64_t, float,
                  float *, float *, float *, void *, int64_t, uint64_t, uint64_t, int64_t, cudaStream_t) { GT_FAIL(kNoTc); }
int tc_predict(const gt_config &, const Layout &, const float *, const float *, const float *, int64_t, float, float *, void *,
               int64_t, cudaStream_t) { GT_FAIL(kNoTc); }

}  // namespace gt
