// peer_opt.cu — the data-parallel exchange of the path as ONE kernel over NVLink peer memory.
//
// The only coupling between the ranks of a data-parallel step is the gradient sum in front of the optimizer
// (BGT/models/train.py:138-141 on the concatenated batch).  The NCCL form is all-reduce(SUM) of the flat gradient, then the
// optimizer kernel.  Here every rank publishes its flat gradient in an exchange buffer that its peers have mapped
// (cudaIpcOpenMemHandle: one process per GPU, NVSwitch gives every GPU full bandwidth to every peer), and the optimizer kernel
// itself reads element i from all `world` buffers, adds them in RANK ORDER — every rank computes bit-identical sums, so the
// replicas cannot drift — and applies SGD / Adam: gradient exchange + optimizer in one pass, no ring, no staging copies, nothing
// left to overlap.  Bytes per rank and step: (world - 1) x 4 n over NVLink (C4: 7 x 13.1 MB = 92 MB, ~0.15 ms at 8 GPUs), against
// ~1.2 ms of exposed all-reduce time of the bucketed NCCL form at the end of a d_model = 256 backward.
//
// Synchronisation is stream-ordered and left to the caller (dp.py): publish (a copy into the local exchange buffer) -> a
// barrier every rank enqueues after its own publish (a tiny NCCL all-reduce that also carries the step's metrics) -> this
// kernel.  The exchange buffer is double buffered by step parity, so no second barrier is needed: when a rank passes the barrier of
// step s + 1, every rank has finished the kernel of step s (stream order), and step s + 2 may overwrite that half.
#include "common.cuh"

namespace gt {

constexpr int PEER_MAX = 16;
struct PeerBufs {
  const float *g[PEER_MAX];
  int world;
};

// peer memory is written by other GPUs between two launches: read it past the L1 (ld.global.cv)
__device__ __forceinline__ float4 peer_ld4(const float *p) { return __ldcv(reinterpret_cast<const float4 *>(p)); }

template <bool ADAM>
__global__ void __launch_bounds__(256) peer_opt_kernel(float *__restrict__ p, PeerBufs pb, float *__restrict__ m, float *__restrict__ v,
                                                       float *__restrict__ gsum, int64_t n, float lr, float b1, float b2, float eps,
                                                       float bc1, float bc2_sqrt, float gs) {
  const int64_t n4 = n >> 2;
  for (int64_t i4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i4 < n4 + (n & 3); i4 += (int64_t)gridDim.x * blockDim.x) {
    float g[4] = {0.f, 0.f, 0.f, 0.f};
    const bool vec = i4 < n4;
    const int64_t i0 = vec ? i4 * 4 : n4 * 4 + (i4 - n4);
    const int cnt = vec ? 4 : 1;
    if (vec) {
      float4 acc = peer_ld4(pb.g[0] + i0);
      for (int r = 1; r < pb.world; ++r) {                // rank order: the same sum on every rank
        const float4 t = peer_ld4(pb.g[r] + i0);
        acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
      }
      g[0] = acc.x; g[1] = acc.y; g[2] = acc.z; g[3] = acc.w;
      if (gsum != nullptr) *reinterpret_cast<float4 *>(gsum + i0) = acc;
    } else {
      float acc = __ldcv(pb.g[0] + i0);
      for (int r = 1; r < pb.world; ++r) acc += __ldcv(pb.g[r] + i0);
      g[0] = acc;
      if (gsum != nullptr) gsum[i0] = acc;
    }
    for (int j = 0; j < cnt; ++j) {
      const int64_t i = i0 + j;
      const float gi = g[j] * gs;
      if (ADAM) {                                          // kernels_simt.cu:adam_kernel, same arithmetic
        const float mi = m[i] + (gi - m[i]) * (1.f - b1);
        const float vi = v[i] * b2 + (1.f - b2) * gi * gi;
        m[i] = mi; v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = p[i] - (lr / bc1) * (mi / denom);
      } else {
        p[i] = p[i] - lr * gi;
      }
    }
  }
}

static int peer_bufs(PeerBufs &pb, const void *const *bufs, int world, int64_t offset_floats) {
  GT_CHECK(bufs != nullptr && world >= 1 && world <= PEER_MAX, "peer optimizer: world size must be in [1, 16]");
  pb.world = world;
  for (int r = 0; r < PEER_MAX; ++r) pb.g[r] = nullptr;
  for (int r = 0; r < world; ++r) {
    GT_CHECK(bufs[r] != nullptr && ((uintptr_t)bufs[r] & 15) == 0, "peer optimizer: null / unaligned exchange buffer");
    pb.g[r] = static_cast<const float *>(bufs[r]) + offset_floats;
  }
  GT_CHECK((offset_floats & 3) == 0, "peer optimizer: the exchange offset must be a multiple of 4 floats");
  return 0;
}

static unsigned peer_grid(int64_t n) {
  const int64_t want = (n / 4 + 255) / 256 + 1;
  return (unsigned)(want < 1 ? 1 : (want > 148 * 8 ? 148 * 8 : want));
}

}  // namespace gt

using namespace gt;

extern "C" {

int gt_peer_alloc(int64_t bytes, void **ptr, uint8_t *handle64) {
  GT_CHECK(bytes > 0 && ptr != nullptr && handle64 != nullptr, "gt_peer_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  void *p = nullptr;
  GT_CUDA(cudaMalloc(&p, (size_t)bytes));
  GT_CUDA(cudaMemset(p, 0, (size_t)bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    GT_FAIL(std::string("gt_peer_alloc: cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
  }
  memcpy(handle64, &h, 64);
  *ptr = p;
  return 0;
}

int gt_peer_open(const uint8_t *handle64, void **ptr) {
  GT_CHECK(handle64 != nullptr && ptr != nullptr, "gt_peer_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void *p = nullptr;
  GT_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *ptr = p;
  return 0;
}

int gt_peer_close(void *ptr) {
  if (ptr != nullptr) GT_CUDA(cudaIpcCloseMemHandle(ptr));
  return 0;
}

int gt_peer_free(void *ptr) {
  if (ptr != nullptr) GT_CUDA(cudaFree(ptr));
  return 0;
}

int gt_peer_publish(void *exchange, int64_t offset_floats, const float *grads, int64_t n, void *stream) {
  GT_CHECK(exchange != nullptr && grads != nullptr && n >= 0 && offset_floats >= 0, "gt_peer_publish: bad arguments");
  GT_CUDA(cudaMemcpyAsync(static_cast<float *>(exchange) + offset_floats, grads, (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice,
                          (cudaStream_t)stream));
  return 0;
}

int gt_sgd_step_peers(float *p, const void *const *exchange_bufs, int world, int64_t offset_floats, float *gsum, int64_t n, float lr,
                      float grad_scale, void *stream) {
  GT_CHECK(p != nullptr && n >= 0, "gt_sgd_step_peers: bad arguments");
  if (n == 0) return 0;
  PeerBufs pb;
  GT_TRY(peer_bufs(pb, exchange_bufs, world, offset_floats));
  GT_NVTX("groove.optimizer");
  { LaunchScope _ls(KC_OPT, (cudaStream_t)stream);
    peer_opt_kernel<false><<<peer_grid(n), 256, 0, (cudaStream_t)stream>>>(p, pb, nullptr, nullptr, gsum, n, lr, 0.f, 0.f, 0.f, 1.f, 1.f, grad_scale); }
  GT_CUDA(cudaGetLastError());
  return 0;
}

int gt_adam_step_peers(float *p, const void *const *exchange_bufs, int world, int64_t offset_floats, float *m, float *v, float *gsum,
                       int64_t n, float lr, float beta1, float beta2, float eps, int64_t step, float grad_scale, void *stream) {
  GT_CHECK(p != nullptr && m != nullptr && v != nullptr && n >= 0, "gt_adam_step_peers: bad arguments");
  GT_CHECK(step >= 1, "Adam step is 1-based");
  if (n == 0) return 0;
  PeerBufs pb;
  GT_TRY(peer_bufs(pb, exchange_bufs, world, offset_floats));
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  GT_NVTX("groove.optimizer");
  { LaunchScope _ls(KC_OPT, (cudaStream_t)stream);
    peer_opt_kernel<true><<<peer_grid(n), 256, 0, (cudaStream_t)stream>>>(p, pb, m, v, gsum, n, lr, beta1, beta2, eps, (float)bc1,
                                                                         (float)sqrt(bc2), grad_scale); }
  GT_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
