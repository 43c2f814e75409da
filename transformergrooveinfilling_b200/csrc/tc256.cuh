// tc256.cuh — fused encoder-layer kernels for d_model = 256 (InfillingRandom_test_large class shapes)
// on tcgen05 / TMEM.
//
// The layer's weights (581 KB in bf16 at the C4 shape) do not fit in shared memory, so they are
// STREAMED: the weight-prep kernel lays every B operand out as a sequence of <= 16 KB "stages"
// (canonical K-major UMMA images, in exactly the order the layer consumes them), a producer warp walks
// that sequence with 1-D bulk-TMA copies into a 4-slot ring, and one MMA-issuer thread consumes the
// ring.  16 compute warps own the epilogues and the per-(sequence, head) attention (mma.sync on
// register fragments) and talk to the issuer through mbarriers, so the q|k|v projection of head group
// g+1 and the partial out-projection of group g-1 run on the tensor pipe while the compute warps are
// inside the softmax of group g.
//
// Activation layouts between the layers (all private to this path; tc256_path.cu converts at the
// stack boundaries):
//   (activations use the bf16 image; gradients flowing between layers use the fp32 tiled layout)
//   fp32 "tiled":  tile = 128 tokens (4 sequences); element (r, c) of a tile at float offset
//                  ((c >> 2) * 128 + r) * 4 + (c & 3)  -> a warp whose lanes are 32 consecutive token
//                  rows reads / writes 512 contiguous bytes per float4 (the TMEM lane = row mapping).
//   bf16 "image":  the canonical K-major UMMA A-operand image of the tile (umma.cuh: kmajor_off with
//                  R = 128), 64 KB per [128 x 256] tile -> staged with one bulk-TMA copy, written by
//                  lane = row threads as coalesced 16-byte chunks.
#pragma once
#include "tc_layers.cuh"

namespace gt {

constexpr uint32_t T256_STAGE = 16384;      // ring slot size / image stride of one stage
constexpr int T256_NS = 4;                  // ring slots
constexpr int T256_REP = 16;                // replicas of every stage stream: CTA b reads replica b % REP, so one L2 line serves 148/REP SMs instead of 148
constexpr int T256_G = 4;                   // head groups: 64 feature columns of q, k and v each
constexpr int T256_CTHREADS = 512;          // 16 compute warps
constexpr int T256_THREADS = 576;           // + warp 16 (TMA producer) + warp 17 (MMA issuer)
constexpr int64_t T256_TILE_F32 = 128 * 256;          // floats per fp32 tile
constexpr int64_t T256_TILE_IMG = 128 * 256 * 2;      // bytes per bf16 image tile

__host__ __device__ inline int64_t t256_tiled_off(int r, int c) { return (int64_t)(((c >> 2) * 128 + r) * 4 + (c & 3)); }

// ---- stage streams -----------------------------------------------------------------------------
// forward:
//   type 0: Wqkv rows of group a, K chunk b (of 8)       B image [192 x  32]   12288 B
//   type 1: Wo   columns of group a, K half b            B image [256 x  32]   16384 B
//   type 2: W1   rows of FFN chunk a, K half b           B image [ 64 x 128]   16384 B
//   type 3: W2   columns of FFN chunk a, K half b        B image [256 x  32]   16384 B
// backward (dgrad):
//   type 4: W2^T of FFN chunk a, K half b   (dH = da2 W2)            B image [ 64 x 128]
//   type 5: W1^T of FFN chunk a, K half b   (dx1 += dH W1)           B image [256 x  32]
//   type 6: Wo^T, K chunk b (of 8)          (dctx = da1 Wo)          B image [256 x  32]
//   type 0: as forward (recompute q|k|v of group a)
//   type 7: Wqkv^T of group a, K chunk b (of 6)  (dx += dqkv_g Wqkv_g)   B image [256 x  32]
struct T256Stage {
  int type, a, b, N, K;
  uint32_t bytes;
};
__host__ __device__ inline void t256_stage_dims(T256Stage &s) {
  switch (s.type) {
    case 0: s.N = 192; s.K = 32; break;
    case 2: case 4: s.N = 64; s.K = 128; break;
    default: s.N = 256; s.K = 32; break;
  }
  s.bytes = (uint32_t)(s.N * s.K * 2);
}
__host__ __device__ inline int t256_fwd_stages(int F) { return 10 * T256_G + 4 * (F / 64); }
__host__ __device__ inline T256Stage t256_fwd_stage(int st) {
  // QKV(0) x8, then for g = 1..G-1: QKV(g) x8, Wo(g-1) x2 ; then Wo(G-1) x2 ; then per FFN chunk W1 x2, W2 x2
  T256Stage s;
  const int nq = 10 * T256_G;
  s.b = 0;
  if (st < nq) {
    if (st < 8) { s.type = 0; s.a = 0; s.b = st; }
    else if (st >= nq - 2) { s.type = 1; s.a = T256_G - 1; s.b = st - (nq - 2); }
    else {
      const int r = st - 8, g = r / 10 + 1, k = r % 10;
      if (k < 8) { s.type = 0; s.a = g; s.b = k; } else { s.type = 1; s.a = g - 1; s.b = k - 8; }
    }
  } else {
    const int r = st - nq, c = r / 4, k = r % 4;
    if (k < 2) { s.type = 2; s.a = c; s.b = k; } else { s.type = 3; s.a = c; s.b = k - 2; }
  }
  t256_stage_dims(s);
  return s;
}
// head_dim 128 (C3: a head spans two 64-column groups): the backward does not recompute q | k | v — the forward saves their bf16
// group images — so its stream has no type-0 stages: per FFN chunk W2^T x2, W1^T x2 ; Wo^T x8 ; WqkvT(g) x6 for g = 0..G-1
__host__ __device__ inline int t256_bwd_stages(int F, int dh = 16) { return 4 * (F / 64) + 8 + (dh == 128 ? 6 : 14) * T256_G; }
__host__ __device__ inline T256Stage t256_bwd_stage(int st, int F, int dh = 16) {
  // per FFN chunk: W2^T x2, W1^T x2 ; Wo^T x8 ; then QKV(0) x8, for g = 1..G-1: QKV(g) x8, WqkvT(g-1) x6 ; WqkvT(G-1) x6
  T256Stage s;
  const int nf = 4 * (F / 64);
  s.b = 0;
  if (dh == 128 && st >= nf + 8) {
    const int r0 = st - nf - 8;
    s.type = 7; s.a = r0 / 6; s.b = r0 % 6;
    t256_stage_dims(s);
    return s;
  }
  if (st < nf) {
    const int c = st / 4, k = st % 4;
    s.a = c;
    if (k < 2) { s.type = 4; s.b = k; } else { s.type = 5; s.b = k - 2; }
  } else if (st < nf + 8) {
    s.type = 6; s.a = 0; s.b = st - nf;
  } else {
    const int r0 = st - nf - 8, na = 14 * T256_G;
    if (r0 < 8) { s.type = 0; s.a = 0; s.b = r0; }
    else if (r0 >= na - 6) { s.type = 7; s.a = T256_G - 1; s.b = r0 - (na - 6); }
    else {
      const int r = r0 - 8, g = r / 14 + 1, k = r % 14;
      if (k < 8) { s.type = 0; s.a = g; s.b = k; } else { s.type = 7; s.a = g - 1; s.b = k - 8; }
    }
  }
  t256_stage_dims(s);
  return s;
}
__host__ __device__ inline uint32_t t256_img_bytes(int F, int dh = 16) { return (uint32_t)(t256_fwd_stages(F) + t256_bwd_stages(F, dh)) * T256_STAGE; }
constexpr int64_t T256_QKV_GROUP_IMG = 128 * 192 * 2;    // bytes of one saved q | k | v group image [128 x 192] (head_dim 128 only)

struct T256Args {
  // activations: the residual stream between layers lives in HBM as bf16 images (tile-padded); gradients as tiled fp32
  const uint8_t *x_img_in;               // layer input: A operand of the q|k|v projection and the first residual
  uint8_t *x_img_out;                    // forward: layer output
  uint8_t *u1_img, *u2_img;              // pre-LayerNorm sums (written by forward in train mode, read by backward)
  uint8_t *x1_img, *ctx_img, *h_img;     // forward (train): images saved for backward / the weight-gradient kernel
  float2 *ln1_stat, *ln2_stat;           // forward (train) writes, backward reads: (mean, rstd) of LayerNorm1 / LayerNorm2 per token row (tile-padded)
  uint8_t *qkv_img;                      // head_dim 128 only: the four [128 x 192] q | k | v group images of every tile (q pre-scaled), saved
                                         // by the forward (train) and read by the backward instead of recomputing them
  // backward
  const float *dy;
  float *dx;
  uint8_t *da2_img, *da1_img, *dh_img, *dqkv_img;   // bf16 images written for the weight-gradient kernel
  uint8_t *dctx_scratch;                 // per-CTA 64 KB scratch (L2 resident)
  const uint8_t *img;                    // this layer's stage streams (forward, then backward), T256_REP replicas
  uint32_t img_rep_stride;
  const float *bqkv, *bo, *b1, *b2, *g1, *be1, *g2, *be2;
  float *gbqkv, *gbo, *gb1, *gb2, *gg1, *gbe1, *gg2, *gbe2;      // bias / LayerNorm gradients (accumulated with atomics)
  int64_t M;                             // valid token rows
  int n_tiles, F, H, dh;
  Drop d_attn, d1, d_ffn, d2;
  int64_t seq0;
  uint32_t stagger;                      // backward: CTA b starts (b % 4) * stagger clocks late, so that the HBM-heavy phases of the persistent CTAs do not coincide
  unsigned long long *dbg;               // optional clock64 timeline of CTA 0 (GT_T256_DBG=n)
};

struct T256WgradArgs {                   // dW accumulation over all tiles: see tc256.cu t256_wgrad_kernel
  const uint8_t *dqkv_img, *x_img, *da1_img, *ctx_img, *dh_img, *x1_img, *da2_img, *h_img;
  float *gwqkv, *gwo, *gw1, *gw2;
  int n_tiles, F;
};

bool t256_shape_supported(const gt_config &c, std::string *why);
int t256_prep_weights(const TcPrepArgs &a, cudaStream_t st);
int t256_layer_fwd(const T256Args &a, cudaStream_t st);
int t256_layer_bwd(const T256Args &a, cudaStream_t st);
constexpr int64_t T256_WG_JOBBUF = 16384;   // device bytes for the weight-gradient job list
int t256_wgrad(const T256WgradArgs &a, void *job_buf, cudaStream_t st);
int t256_wgrad_pair(const uint8_t *a_img, const uint8_t *b_img, int N, float *out, int n_tiles, void *job_buf, cudaStream_t st);
// fused ends of the stack on the tile-native layouts (edge256.cu)
int64_t edge256_loss_partials();
int edge256_stem_fwd(const float *src, int E, const float *W, const float *b, const float *pe, uint8_t *x_img, int64_t M, int n_tiles,
                     const Drop &drop, int64_t row0, cudaStream_t st);
int edge256_stem_bwd(const float *dx_tiled, const float *src, int E, const float *W, const float *b, uint8_t *g_img, uint8_t *src_img,
                     float *scratch, void *job_buf, float *gW, float *gb, int64_t M, int n_tiles, const Drop &drop, int64_t row0,
                     cudaStream_t st);
int edge256_tail_fwd(const uint8_t *x_img, const float *gamma, const float *beta, const float *Wout, const float *bout, float *hvo, float *mean,
                     float *rstd, int64_t M, int n_tiles, float thres, const float *y, float penalty, float *dlog, float *partials,
                     float *metrics6, cudaStream_t st);
int edge256_tail_bwd(const float *d_in, const float *hvo, const uint8_t *x_img, const float *mean, const float *rstd, const float *gamma,
                     const float *beta, const float *Wout, float *dx_tiled, uint8_t *dl_img, float *scratch, void *job_buf, float *gW,
                     float *gb, float *gg, float *gbe, int64_t M, int n_tiles, cudaStream_t st);
int t256_debug_umma_rate(int N, int n_mma, int ksteps, float *out, cudaStream_t st);
// layout conversion at the stack boundaries
int t256_to_image(const float *rowmajor, uint8_t *img, int64_t M, int n_tiles, cudaStream_t st);
int t256_from_image(const uint8_t *img, float *rowmajor, int64_t M, cudaStream_t st);
int t256_to_tiled(const float *rowmajor, float *tiled, int64_t M, int n_tiles, cudaStream_t st);
int t256_from_tiled(const float *tiled, float *rowmajor, int64_t M, cudaStream_t st);

// model passes (tc256_path.cu)
int64_t t256_workspace_bytes(const gt_config &c, int64_t n_seq, int mode);
int t256_forward(const gt_config &c, const Layout &L, const float *params, const float *pe, const float *src, int64_t n_seq,
                 float *hvo, void *ws, int64_t ws_bytes, bool train, uint64_t seed, uint64_t step, int64_t seq0, cudaStream_t st);
int t256_backward(const gt_config &c, const Layout &L, const float *params, const float *pe, const float *src, int64_t n_seq,
                  const float *hvo, const float *d_hvo, float *grads, void *ws, int64_t ws_bytes, uint64_t seed, uint64_t step,
                  int64_t seq0, cudaStream_t st);
int t256_train_step(const gt_config &c, const Layout &L, const float *params, const float *pe, const float *src, const float *y,
                    int64_t n_seq, float penalty, float *grads, float *metrics6, float *hvo, void *ws, int64_t ws_bytes,
                    uint64_t seed, uint64_t step, int64_t seq0, cudaStream_t st);
int t256_predict(const gt_config &c, const Layout &L, const float *params, const float *pe, const float *src, int64_t n_seq,
                 float thres, float *hvo_out, void *ws, int64_t ws_bytes, cudaStream_t st);

}  // namespace gt
