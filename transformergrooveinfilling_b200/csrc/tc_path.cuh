// tc_path.cuh — bf16 tensor-core (tcgen05 / TMEM / bulk-TMA) path, precision mode GT_PREC_BF16.
#pragma once
#include "common.cuh"

namespace gt {

int64_t tc_workspace_bytes(const gt_config &c, int64_t n_seq, int mode);
int tc_forward(const gt_config &c, const Layout &L, const float *params, const float *pe, const float *src,
               const float *tgt_in, int64_t n_seq, float *hvo, void *ws, int64_t ws_bytes, bool train, uint64_t seed,
               uint64_t step, int64_t seq0, cudaStream_t st);
int tc_backward(const gt_config &c, const Layout &L, const float *params, const float *pe, const float *src,
                const float *tgt_in, int64_t n_seq, const float *hvo, const float *d_hvo, float *grads, void *ws,
                int64_t ws_bytes, uint64_t seed, uint64_t step, int64_t seq0, cudaStream_t st);
int tc_train_step(const gt_config &c, const Layout &L, const float *params, const float *pe, const float *src,
                  const float *y, int64_t n_seq, float penalty, float *grads, float *metrics6, float *hvo, void *ws,
                  int64_t ws_bytes, uint64_t seed, uint64_t step, int64_t seq0, cudaStream_t st);
int tc_predict(const gt_config &c, const Layout &L, const float *params, const float *pe, const float *src, int64_t n_seq,
               float thres, float *hvo_out, void *ws, int64_t ws_bytes, cudaStream_t st);
// encoder stack of an encoder-decoder model on the fused d_model = 32 kernels (runner.cu hybrid path)
bool tc_encoder_supported(const gt_config &c);
uint32_t tc_enc_img_stride(const gt_config &c);
int tc_enc_prep(const gt_config &c, const Layout &L, const float *params, uint8_t *img, cudaStream_t st);
int tc_enc_layer_fwd(const gt_config &c, const Layout &L, const float *params, uint8_t *img, int l, const float *x_in, float *x_out,
                     float *u1, float *u2, int64_t n_seq, bool train, uint64_t seed, uint64_t step, int64_t seq0, cudaStream_t st);
int tc_enc_layer_bwd(const gt_config &c, const Layout &L, const float *params, float *grads, uint8_t *img, int l, const float *x_in,
                     const float *u1, const float *u2, const float *dy, float *dx, int64_t n_seq, uint64_t seed, uint64_t step,
                     int64_t seq0, cudaStream_t st);
// the three blocks of a decoder layer on the fused kernels (TC_MODE_ATTN_CAUSAL / TC_MODE_ATTN_CROSS / TC_MODE_FFN);
// dec_img: 2 * n_dec * tc_enc_img_stride bytes
bool tc_dec_attn_supported(const gt_config &c);
int tc_dec_attn_fwd(const gt_config &c, const Layout &L, const float *params, uint8_t *dec_img, int l, int cross, const float *x_in,
                    const float *mem, float *x_out, float *u, int64_t n_seq, bool train, uint64_t seed, uint64_t step, int64_t seq0,
                    cudaStream_t st);
int tc_dec_attn_bwd(const gt_config &c, const Layout &L, const float *params, float *grads, uint8_t *dec_img, int l, int cross,
                    const float *x_in, const float *mem, const float *u, const float *dy, float *dx, float *dmem, int64_t n_seq,
                    uint64_t seed, uint64_t step, int64_t seq0, cudaStream_t st);
int tc_dec_prep(const gt_config &c, const Layout &L, const float *params, uint8_t *dec_img, cudaStream_t st);
int tc_dec_ffn_fwd(const gt_config &c, const Layout &L, const float *params, uint8_t *dec_img, int l, const float *x_in, float *x_out,
                   float *u, int64_t n_seq, bool train, uint64_t seed, uint64_t step, int64_t seq0, cudaStream_t st);
int tc_dec_ffn_rows(const gt_config &c, const Layout &L, const float *params, uint8_t *dec_img, int l, const float *x_in, float *x_out,
                    int64_t n_rows, cudaStream_t st);
int tc_dec_ffn_bwd(const gt_config &c, const Layout &L, const float *params, float *grads, uint8_t *dec_img, int l, const float *x_in,
                   const float *u, const float *dy, float *dx, int64_t n_seq, uint64_t seed, uint64_t step, int64_t seq0,
                   cudaStream_t st);
int tc_debug_gemm(const uint16_t *a, const uint16_t *b, float *d, int m, int n, int k, int variant, cudaStream_t st);

}  // namespace gt
