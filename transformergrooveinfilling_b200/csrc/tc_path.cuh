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
int tc_debug_gemm(const uint16_t *a, const uint16_t *b, float *d, int m, int n, int k, int variant, cudaStream_t st);

}  // namespace gt
