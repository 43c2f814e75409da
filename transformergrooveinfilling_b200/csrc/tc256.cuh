// tc256.cuh — fused encoder-layer kernels for d_model = 256 (C3 / C4 class shapes) on tcgen05 / TMEM.
//
// The layer's weights (581 KB in bf16 at the InfillingRandom_test_large shape) do not fit in shared
// memory, so they are STREAMED: the weight-prep kernel lays every B operand out as a sequence of
// <= 16 KB "stages" (canonical K-major UMMA images, in exactly the order the layer consumes them), a
// producer warp walks that sequence with 1-D bulk-TMA copies into a 4-slot ring, and one MMA-issuer
// thread consumes the ring.  16 compute warps own the epilogues and the per-(sequence, head)
// attention (mma.sync on register fragments), and talk to the issuer through mbarriers, so the
// q|k|v projection of head group g+1 and the partial out-projection of group g-1 run on the tensor
// pipe while the compute warps are inside the softmax of group g.
#pragma once
#include "tc_layers.cuh"

namespace gt {

constexpr uint32_t T256_STAGE = 16384;      // ring slot size / image stride of one stage
constexpr int T256_NS = 4;                  // ring slots
constexpr int T256_G = 8;                   // head groups: 32 feature columns of q, k and v each
constexpr int T256_CTHREADS = 512;          // 16 compute warps
constexpr int T256_THREADS = 576;           // + warp 16 (TMA producer) + warp 17 (MMA issuer)

// ---- forward stage stream ---------------------------------------------------------------------
// type 0: Wqkv rows of group a, K chunk b (of 4)   B image [ 96 x  64]   12288 B
// type 1: Wo   columns of group a                  B image [256 x  32]   16384 B
// type 2: W1   rows of FFN chunk a, K half b       B image [ 64 x 128]   16384 B
// type 3: W2   columns of FFN chunk a, K half b    B image [256 x  32]   16384 B
struct T256Stage {
  int type, a, b, N, K;
  uint32_t bytes;
};
__host__ __device__ inline int t256_fwd_stages(int F) { return 5 * T256_G + 4 * (F / 64); }
__host__ __device__ inline T256Stage t256_fwd_stage(int st) {
  T256Stage s;
  const int nq = 5 * T256_G;
  if (st < nq) {
    int type, a, b = 0;
    if (st < 4) { type = 0; a = 0; b = st; }
    else if (st == nq - 1) { type = 1; a = T256_G - 1; }
    else {
      const int r = st - 4, g = r / 5 + 1, k = r % 5;
      if (k < 4) { type = 0; a = g; b = k; } else { type = 1; a = g - 1; }
    }
    s.type = type; s.a = a; s.b = b;
    if (type == 0) { s.N = 96; s.K = 64; } else { s.N = 256; s.K = 32; }
  } else {
    const int r = st - nq, c = r / 4, k = r % 4;
    if (k < 2) { s.type = 2; s.a = c; s.b = k; s.N = 64; s.K = 128; }
    else { s.type = 3; s.a = c; s.b = k - 2; s.N = 256; s.K = 32; }
  }
  s.bytes = (uint32_t)(s.N * s.K * 2);
  return s;
}

// ---- backward (dgrad) stage stream: transposed weights, again as K-major B images ------------------
// type 0: W1  rows of FFN chunk a, K half b  (recompute H)         [ 64 x 128]
// type 1: W2^T rows (= F index) of chunk a, K quarter b (dH)       [ 64 x  64]  ( 8192 B)
// type 2: W1^T chunk a, K half b  (dx1 += dH W1)                   [256 x  32]
// type 3: Wo^T output-column group a, K chunk b (dctx group)       [ 32 x 256]  (16384 B) -> N = 32, K = 256
// type 4: Wqkv rows of group a, K chunk b (recompute q|k|v)        [ 96 x  64]
// type 5: Wqkv^T for group a, K half b (dx += dqkv_g Wqkv_g)       [256 x  48]  hmm see tc256.cu
__host__ __device__ inline uint32_t t256_img_bytes(int F) {
  // forward stream, then the backward stream (tc256.cu: t256_bwd_stages)
  return (uint32_t)(t256_fwd_stages(F) + (4 * (F / 64) * 2 + T256_G * 8)) * T256_STAGE;
}

bool t256_shape_supported(const gt_config &c, std::string *why);
int t256_prep_weights(const TcPrepArgs &a, cudaStream_t st);
int t256_layer_fwd(const TcLayerArgs &a, cudaStream_t st);
int t256_layer_bwd(const TcLayerArgs &a, cudaStream_t st);

}  // namespace gt
