// tc_layers.cuh — fused encoder-layer kernels on tcgen05 / TMEM (bf16 operands, fp32 accumulate).
//
// Tile = 4 whole sequences = 128 tokens = the M of every UMMA.  One persistent CTA (256 threads)
// per SM keeps the layer's bf16 weight images resident in shared memory (bulk-TMA staged once per
// launch) and walks tiles; thread t owns token row (t & 127) — the TMEM lane its warp can read —
// and column half (t >> 7) of wide epilogues.
#pragma once
#include "common.cuh"

namespace gt {

constexpr int TC_TILE = 128;
constexpr int TC_MAX_LAYERS = 16;
// whole encoder layer / feed-forward block alone / attention block alone (attention + out-proj + residual + LayerNorm), causal
// self-attention or cross-attention over an encoder-memory tile: the three blocks of a decoder layer
constexpr int TC_MODE_LAYER = 0, TC_MODE_FFN = 1, TC_MODE_ATTN_CAUSAL = 2, TC_MODE_ATTN_CROSS = 3;

// linear1's bias rides in the contraction: the A image of the FFN input carries TC_KAUG extra K columns (column D = 1.0, the
// rest 0) and every W1 chunk image the matching columns (column D = bf16(b1), the rest 0), so the accumulator the epilogue
// drains already holds x1 W1^T + b1 and the epilogue is pack-with-ReLU + dropout mask only.
constexpr int TC_KAUG = 16;
struct TcImg {               // byte offsets of the bf16 operand images of one layer
  uint32_t wqkv, wo, w1, w2, total;
};
__host__ __device__ inline uint32_t tc_w1_chunk_bytes(int D, int FC) { return (uint32_t)(FC * (D + TC_KAUG) * 2); }
__host__ __device__ inline TcImg tc_img(int D, int F) {
  TcImg o;
  o.wqkv = 0;
  o.wo = (uint32_t)(3 * D * D * 2);
  o.w1 = o.wo + (uint32_t)(D * D * 2);
  o.w2 = o.w1 + (uint32_t)(F * (D + TC_KAUG) * 2);
  o.total = o.w2 + (uint32_t)(D * F * 2);
  return o;
}
inline int tc_ffn_chunk(int F) {
  const int cand[5] = {128, 96, 64, 32, 16};
  for (int i = 0; i < 5; ++i)
    if (F % cand[i] == 0) return cand[i];
  return 0;
}

struct TcPrepArgs {
  const float *params;
  uint8_t *img;                        // n_layers * img_stride bytes
  int64_t w_in[TC_MAX_LAYERS], w_out[TC_MAX_LAYERS], w1[TC_MAX_LAYERS], w2[TC_MAX_LAYERS], b1[TC_MAX_LAYERS];
  uint32_t img_stride;
  int n_layers, D, F, FC;
  int dh;                              // d_model = 256 streams only: head dim (128 selects the backward stream without q | k | v recompute)
};

struct TcLayerArgs {
  // forward: x_in -> (u1, u2 saved when non-null) -> x_out.   backward: + dy (grad wrt x_out) -> dx
  const float *x_in, *u1_in, *u2_in, *dy;
  float *u1, *u2, *x_out, *dx;
  const float *mem;                    // TC_MODE_ATTN_CROSS: encoder memory [tokens, D] (keys / values)
  float *dmem;                         // ... and its gradient (accumulated: every decoder layer adds to it)
  const uint8_t *img;
  uint32_t img_bytes;
  const float *bqkv, *bo, *b1, *b2, *g1, *be1, *g2, *be2;
  float *gwqkv, *gbqkv, *gwo, *gbo, *gw1, *gb1, *gw2, *gb2, *gg1, *gbe1, *gg2, *gbe2;   // backward only
  int64_t M;                           // valid token rows
  int n_tiles, F, FC, H, dh;
  Drop d_attn, d1, d_ffn, d2;
  int64_t seq0;
  int mode;                            // TC_MODE_LAYER / TC_MODE_FFN (FFN block: x_in, u2, b1, b2, g2, be2, d_ffn, d2 only)
};

bool tc_shape_supported(const gt_config &c, std::string *why);
int tc_prep_weights(const TcPrepArgs &a, cudaStream_t st);
int tc_layer_fwd(int D, const TcLayerArgs &a, cudaStream_t st);
int tc_layer_bwd(int D, const TcLayerArgs &a, cudaStream_t st);

}  // namespace gt
