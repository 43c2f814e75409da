// tc256_dev.cuh — device helpers shared by the d_model = 256 forward / backward / weight-gradient kernels.
#pragma once
#include "tc256.cuh"
#include "umma.cuh"

namespace gt {
using namespace umma;

// 64-bit UMMA descriptors are built once per operand and advanced by adding (bytes >> 4) to the address field
__device__ __forceinline__ float bf16lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }
__device__ __forceinline__ uint64_t desc_adv(uint64_t d, uint32_t bytes) { return d + (uint64_t)(bytes >> 4); }
__device__ __forceinline__ uint64_t descA128(uint32_t base) { return make_desc(base, 2048u, 128u); }                 // A image with 128 rows: k16 step = 4096 B
__device__ __forceinline__ uint64_t descB(uint32_t base, int N) { return make_desc(base, (uint32_t)(N >> 3) * 128u, 128u); }   // k16 step = N * 32 B


// per-lane vector of 32 values -> lane l receives the sum over the warp's lanes of v[l]  (31 shuffles)
__device__ __forceinline__ float t256_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = up ? v[i] : v[i + s];
      const float keep = up ? v[i + s] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

#define T256_STAMP() do { if (dbg_on && ndbg < 60) a.dbg[ndbg++] = clock64(); } while (0)

// optional clock64 timeline of CTA 0 (GT_T256_DBG=<number of launches to trace>).  Developer builds only (-DGT_T256_TIMELINE):
// it allocates its own 4 KB device buffer, which the release library must not do (the caller owns all device memory).
#ifndef GT_T256_TIMELINE
struct T256Dbg {
  bool arm(T256Args &, cudaStream_t) { return false; }
  void report(const char *, cudaStream_t) {}
};
#else
struct T256Dbg {
  unsigned long long *buf = nullptr;
  int left = getenv("GT_T256_DBG") ? atoi(getenv("GT_T256_DBG")) : 0;
  bool arm(T256Args &a, cudaStream_t st) {
    if (left <= 0) return false;
    if (!buf) cudaMalloc(&buf, 512 * sizeof(unsigned long long));
    cudaMemsetAsync(buf, 0, 512 * sizeof(unsigned long long), st);
    a.dbg = buf;
    return true;
  }
  void report(const char *what, cudaStream_t st) {
    --left;
    unsigned long long h[512];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, buf, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "%s timeline (clk since tile start):", what);
    for (int i = 1; i < 60 && h[i]; ++i) fprintf(stderr, " %llu", h[i] - h[0]);
    fprintf(stderr, "\n  mma stage-available:");
    for (int i = 64; i < 184 && h[i]; ++i) fprintf(stderr, " %lld", (long long)(h[i] - h[0]));
    fprintf(stderr, "\n  producer slot-free:");
    for (int i = 192; i < 312 && h[i]; ++i) fprintf(stderr, " %lld", (long long)(h[i] - h[0]));
    fprintf(stderr, "\n");
  }
};
#endif
int t256_num_sms();

}  // namespace gt
