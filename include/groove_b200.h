/* groove_b200.h — C ABI of the B200-native Transformer Groove Infilling hot path.
 *
 * The reference (pelinski/TransformerGrooveInfilling) has no FFI: its hot path is pure PyTorch
 * eager code.  Each entry point below therefore cites the reference *Python* interface whose
 * arithmetic it replaces (paths relative to the reference root; BGT = BaseGrooveTransformers).
 * The reference-side binding a maintainer would add is the ctypes stub shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name says host;
 *   - the caller (PyTorch) owns all buffers; the library borrows them for the duration of the
 *     enqueue and keeps no reference (gt_peer_alloc / gt_peer_free are the one explicit allocation pair: an IPC
 *     handle names a whole cudaMalloc allocation);
 *   - every call enqueues asynchronously on the cudaStream_t passed as `stream` (void* here);
 *   - return value 0 = success; non-zero = error, message via gt_last_error() (thread-local);
 *     no C++ exception crosses this boundary; shape / alignment violations are rejected before
 *     any launch;
 *   - all tensors are float32, row-major, batch-first: src [n_seq,32,e_src], y/hvo [n_seq,32,e_tgt] (27 below stands for e_tgt).
 */
#ifndef GROOVE_B200_H
#define GROOVE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GT_ABI_VERSION 1
#define GT_T_STEPS 32          /* train.py:128 `max_len: 32`; BGT/models/utils.py:49 requires T == max_len */

/* precision modes */
#define GT_PREC_FP32 0         /* fp32 SIMT arithmetic everywhere (the 1e-4 "fp32/TF32" parity mode) */
#define GT_PREC_BF16 1         /* bf16 operands / fp32 accumulate on tcgen05 tensor cores (2e-3 mode) */
#define GT_PREC_FP32_TC 2      /* the 1e-4 parity mode ON the tensor cores: every Linear contraction runs on tcgen05 with its fp32
                                  operands split exactly into three bf16 terms (x0 + x1 + x2) and the six products down to 2^-18
                                  contracted with fp32 accumulation ("3 x bf16", the bf16 form of 3xTF32: measured gradient
                                  error 1 - 2e-6 against 5e-7 of the FFMA kernels and 3e-3 of bf16 operands); attention,
                                  LayerNorm, loss and optimizers are the fp32 kernels of GT_PREC_FP32 */

/* which implementation a configuration runs on (gt_path_kind) */
#define GT_PATH_FP32_SIMT   0  /* precision fp32: fp32 FMA kernels */
#define GT_PATH_FUSED_D32   1  /* precision bf16, d_model = 32: fused tcgen05 layer kernels — encoder-only models, and encoder-decoder
                                  models with head_dim 2 / 4 / 8 (every decoder layer = three fused block launches) */
#define GT_PATH_FUSED_D256  2  /* precision bf16, d_model = 256, head_dim 16 / 32, encoder-only: weight-streaming fused kernels */
#define GT_PATH_GEMM_TC     3  /* precision bf16, every other shape: per-op kernels, contractions on gemm_tc, attention on mma.sync
                                  (d_model = 32 encoder-decoder models outside the fused head dims still run their encoder
                                  stack and decoder FFN blocks in the fused kernels) */
#define GT_PATH_GEMM_TC_SPLIT 4 /* precision fp32_tc, every shape: the per-op kernels of GT_PATH_FP32_SIMT with the contractions on
                                  gemm_tc in split form */

/* Mirrors the constructor arguments of GrooveTransformerEncoder / GrooveTransformer
 * (BGT/models/transformer.py:10-11, :87-88) and params["model"] of train.py:115-143. */
typedef struct gt_config {
  int32_t d_model;
  int32_t nhead;
  int32_t dim_ff;
  int32_t n_enc;       /* num_encoder_layers */
  int32_t n_dec;       /* num_decoder_layers; 0 = encoder-only model */
  int32_t e_src;       /* embedding_size_src (16 MSO / 27 symbolic) */
  int32_t e_tgt;       /* embedding_size_tgt = 3 x n_voices (hits | velocities | offsets thirds, BGT/models/io_layers.py:34-40);
                          27 in every set of the reference — the fused stem / tail kernels cover 27, other widths take the
                          generic kernels */
  int32_t precision;   /* GT_PREC_* */
  float   dropout;     /* p of every nn.Dropout on the path */
  int32_t reserved;
} gt_config;

int         gt_version(void);
const char *gt_last_error(void);
/* GT_PATH_* for this configuration, or a negative value for an invalid configuration. */
int         gt_path_kind(const gt_config *cfg);

/* Flat fp32 parameter vector.  Tensors appear in the reference's state_dict order with the `pe`
 * buffers skipped (enumerated in SURVEY.md §8b); each tensor starts on a 16-byte boundary.
 * gt_param_layout writes up to `max_entries` (offset,size) pairs in floats and returns the number
 * of tensors (or a negative error). */
int64_t gt_param_count(const gt_config *cfg);
int     gt_param_layout(const gt_config *cfg, int64_t *offsets, int64_t *sizes, int max_entries);

/* Scratch the caller must provide.  mode 0 = inference forward (ping-pong buffers only),
 * 1 = training (activations saved for backward), 2 = gt_predict (for encoder-decoder models: per-layer key/value
 * caches of the incremental decode; same as 0 for encoder-only models). */
int64_t gt_workspace_bytes(const gt_config *cfg, int64_t n_seq, int mode);

/* Forward pass: BGT/models/transformer.py:108-115 (encoder-only: InputLayer -> Encoder -> OutputLayer)
 * and :35-46 (encoder-decoder; tgt_in is the already right-shifted target, BGT/models/train.py:130-131).
 * hvo[n,32,27] receives channels 0-8 raw hit logits, 9-17 sigmoid, 18-26 0.5*tanh
 * (BGT/models/io_layers.py:36-48).  train!=0 applies dropout (counter-based masks keyed by
 * seed/step/global sequence index seq0+n) and saves activations in `ws` for gt_backward. */
int gt_forward(const gt_config *cfg, const float *params, const float *pe,
               const float *src, const float *tgt_in, int64_t n_seq,
               float *hvo, void *ws, int64_t ws_bytes,
               int train, uint64_t seed, uint64_t step, int64_t seq0, void *stream);

/* Backward of gt_forward(train=1) — what `loss.backward()` does at BGT/models/train.py:138.
 * hvo is the output gt_forward produced; d_hvo[n,32,27] is the gradient w.r.t. those ACTIVATED
 * outputs; gradients are ACCUMULATED (+=)
 * into the flat `grads` vector (same layout as params).  d_src / d_tgt_in are not produced. */
int gt_backward(const gt_config *cfg, const float *params, const float *pe,
                const float *src, const float *tgt_in, int64_t n_seq,
                const float *hvo, const float *d_hvo, float *grads, void *ws, int64_t ws_bytes,
                uint64_t seed, uint64_t step, int64_t seq0, void *stream);

/* calculate_loss, BGT/models/train.py:9-40.  metrics6 = {total, hit_accuracy, hit_perplexity,
 * bce_hits, mse_velocities, mse_offsets}.  If d_hvo != NULL it receives grad_scale * dLoss/d(hvo).
 * `partials` is scratch of gt_loss_scratch_floats(n_seq) floats (two-stage deterministic sum). */
int64_t gt_loss_scratch_floats(int64_t n_seq);
int gt_loss(const float *hvo, const float *y, int64_t n_seq, float hit_loss_penalty,
            float *metrics6, float *d_hvo, float grad_scale, float *partials, void *stream);
/* the same for hvo / y of [n_seq,32,3*n_voices] (train.py:12-13 splits the last axis into thirds); gt_loss = 9 voices */
int gt_loss_voices(const float *hvo, const float *y, int64_t n_seq, int n_voices, float hit_loss_penalty,
                   float *metrics6, float *d_hvo, float grad_scale, float *partials, void *stream);

/* Evaluator metrics — the step right after predict() in the reference's per-epoch evaluation
 * (GrooveEvaluator/GrooveEvaluator/evaluator.py:189-251 get_hits_accuracies / get_velocity_errors /
 * get_micro_timing_errors, fed by evaluator.py:171-186 with np.concatenate(model.predict(...), axis=2)).
 * pred_hvo / gt_hvo are [n_seq,32,3*n_voices] (hits | velocities | offsets).  out receives 3*(n_voices+1) floats:
 * {hit accuracy, velocity MSE, micro-timing MSE} x {voice 0 .. n_voices-1, Overall}.  partials is scratch of
 * gt_eval_scratch_floats(n_seq, n_voices) floats (two-stage deterministic sum). */
int64_t gt_eval_scratch_floats(int64_t n_seq, int n_voices);
int gt_eval_metrics(const float *pred_hvo, const float *gt_hvo, int64_t n_seq, int n_voices,
                    float *out, float *partials, void *stream);

/* One fused training step body (BGT/models/train.py:126-138 without the optimizer): forward with
 * dropout, loss + metrics, backward.  `grads` is ZEROED first, then holds dLoss/dparams.
 * For n_dec>0 the shifted target is built internally from y (train.py:130-131). */
int gt_train_step(const gt_config *cfg, const float *params, const float *pe,
                  const float *src, const float *y, int64_t n_seq, float hit_loss_penalty,
                  float *grads, float *metrics6, float *hvo, void *ws, int64_t ws_bytes,
                  uint64_t seed, uint64_t step, int64_t seq0, void *stream);

/* predict(): BGT/models/transformer.py:117-125 (encoder-only: forward + threshold) and :48-83
 * (encoder-decoder: 32-step autoregressive loop feeding back thresholded hits and raw v,o).
 * hvo_out[n,32,27] channels 0-8 are 0.0/1.0 hits (BGT/models/utils.py:59-69, threshold path only). */
int gt_predict(const gt_config *cfg, const float *params, const float *pe,
               const float *src, int64_t n_seq, float thres,
               float *hvo_out, void *ws, int64_t ws_bytes, void *stream);
/* Encoder-decoder gt_predict runs an incremental decode: the target mask is causal (BGT/models/utils.py:53-56), so
 * caching each decoder layer's self-attention keys/values (and the cross-attention keys/values of the encoder
 * memory) reproduces the reference's 32 full decoder passes with 1/32 of the decoder work.  variant 0 = that decode
 * (workspace mode 2); variant 1 = the reference's literal 32-pass loop (workspace mode 0), kept as its cross-check. */
int gt_predict_variant(const gt_config *cfg, const float *params, const float *pe,
                       const float *src, int64_t n_seq, float thres,
                       float *hvo_out, void *ws, int64_t ws_bytes, int variant, void *stream);

/* torch.optim.SGD(lr).step() / torch.optim.Adam(lr).step() as called at BGT/models/train.py:141,
 * over the flat vectors.  g is multiplied by grad_scale first (1/world after an all-reduce SUM).
 * `step` for Adam is 1-based. */
int gt_sgd_step(float *p, const float *g, int64_t n, float lr, float grad_scale, void *stream);
int gt_adam_step(float *p, const float *g, float *m, float *v, int64_t n, float lr,
                 float beta1, float beta2, float eps, int64_t step, float grad_scale, void *stream);

/* Data-parallel exchange + optimizer as ONE kernel over NVLink peer memory (one process per GPU of one node; SURVEY.md 8e: the
 * gradient sum in front of the optimizer is the path's only exchange step).  gt_peer_alloc cudaMallocs an exchange buffer and
 * returns its 64-byte cudaIpcMemHandle_t (ship it to the peers by any means, e.g. torch.distributed.all_gather_object);
 * gt_peer_open maps a peer's buffer (cudaIpcMemLazyEnablePeerAccess); gt_peer_publish copies the rank's flat gradient into its
 * own buffer at offset_floats (double buffered by step parity by the caller); after a barrier every rank enqueued behind its
 * publish, gt_*_step_peers reads element i of ALL `world` buffers (exchange_bufs[r] = rank r's buffer as mapped in THIS process,
 * own buffer included), adds them in rank order (bit-identical on every rank) and applies the update of gt_sgd_step / gt_adam_step;
 * gsum (may be NULL) receives the summed gradient.  The buffers are the caller's to close / free. */
int gt_peer_alloc(int64_t bytes, void **ptr, uint8_t *handle64);
int gt_peer_open(const uint8_t *handle64, void **ptr);
int gt_peer_close(void *ptr);
int gt_peer_free(void *ptr);
int gt_peer_publish(void *exchange, int64_t offset_floats, const float *grads, int64_t n, void *stream);
int gt_sgd_step_peers(float *p, const void *const *exchange_bufs, int world, int64_t offset_floats, float *gsum, int64_t n,
                      float lr, float grad_scale, void *stream);
int gt_adam_step_peers(float *p, const void *const *exchange_bufs, int world, int64_t offset_floats, float *m, float *v,
                       float *gsum, int64_t n, float lr, float beta1, float beta2, float eps, int64_t step, float grad_scale,
                       void *stream);

/* Several consecutive training steps in ONE call over a dataset that is resident on the device (the body of
 * BGT/models/train.py:118-141 for `n_steps` batches of one epoch): step s gathers rows perm[start + s*batch .. + batch) of
 * data_x [S,32,e_src] / data_y [S,32,27] into xbuf / ybuf, runs gt_train_step on them with dropout step counter step0 + s,
 * and applies the optimizer (optimizer 0: SGD, 1: Adam with torch defaults and 1-based counter adam_t0 + s + 1; m, v may be
 * NULL for SGD).  metrics_out[s*6 .. s*6+6) receives the six calculate_loss values of step s.  Everything is enqueued on
 * `stream`; the caller's host thread is busy only for the launches, which is what lets several sweep members be driven
 * from several host threads at once (transformergrooveinfilling_b200/sweep.py).  perm holds int64 row indices on the device. */
int gt_train_steps(const gt_config *cfg, float *params, const float *pe, const float *data_x, const float *data_y,
                   const int64_t *perm, int64_t start, int64_t batch, int n_steps, float hit_loss_penalty, float *grads,
                   float *metrics_out, float *hvo, float *xbuf, float *ybuf, void *ws, int64_t ws_bytes, int optimizer,
                   float lr, float *m, float *v, int64_t adam_t0, uint64_t seed, uint64_t step0, void *stream);

/* One training step captured into a CUDA graph: [row gather of the batch] + gt_train_step + optimizer + bookkeeping.  The
 * latency-bound regimes (the reference's batch sizes of 16..512, BGT/models/train.py:118-141 once per batch, and several sweep
 * members per GPU) pay ONE graph launch per step instead of ~47 kernel launches.  xbuf / ybuf / metrics6 / hvo / ws / grads are
 * the fixed buffers the graph reads and writes.  `counters` points to four device uint64 the graph reads AND advances:
 *   [0] dropout step counter — the kernels derive their dropout keys from it with site_key's arithmetic, so the masks are those
 *       of gt_train_step(step = counters[0]);           [1] Adam steps taken so far (bias corrections use [1] + 1);
 *   [2] row offset into `perm` of this step's batch (+= n_seq per replay);   [3] metrics slot (+= 1 per replay).
 * data_x / data_y / perm (all three or none): the device-resident dataset and the epoch's permutation (a fixed buffer the caller
 * refills per epoch, DataLoader(shuffle=True) of train.py:153-158); when given, each replay first gathers rows
 * perm[counters[2] .. +n_seq) into xbuf / ybuf, otherwise the caller fills xbuf / ybuf before each launch.  metrics_ring
 * (optional, ring_slots x 6 floats): replay k stores its six calculate_loss values at slot counters[3] % ring_slots.
 * Available on every path (GT_PATH_*), for encoder-only and encoder-decoder models (for n_dec > 0 the shifted target is built
 * from ybuf inside the graph, BGT/models/train.py:130-131).  optimizer: 0 SGD, 1 Adam (torch
 * defaults).  Per-kernel profiling (gt_profile_enable) and bucket events (gt_grad_events_enable) must be off. */
int gt_graph_train_create(const gt_config *cfg, float *params, const float *pe, float *xbuf, float *ybuf, int64_t n_seq,
                          float hit_loss_penalty, float *grads, float *metrics6, float *hvo, void *ws, int64_t ws_bytes,
                          int optimizer, float lr, float *m, float *v, uint64_t seed, unsigned long long *counters,
                          const float *data_x, const float *data_y, const int64_t *perm, float *metrics_ring,
                          int64_t ring_slots, void *stream, void **graph_out);
/* n_replays consecutive steps (each replay advances the counters, so replay k trains on the k-th next batch of the permutation) */
int gt_graph_launch(void *graph, int n_replays, void *stream);
int gt_graph_destroy(void *graph);

/* Data-parallel overlap (BASELINE.json north_star: "bucketed NCCL gradient allreduce ... overlapped with backward").
 * The reference has no distributed code; the flat gradient of gt_backward / gt_train_step is partitioned into
 * contiguous buckets listed in the order backward FINISHES them (output head first, one bucket per decoder / encoder
 * layer in reverse layer order, encoder input layer last) — a descending partition of the flat vector, so consecutive
 * buckets can be merged into one range.  gt_grad_buckets returns their number and (offset,size) in floats.
 * After gt_grad_events_enable(n >= bucket count), every gt_backward / gt_train_step records one library-owned CUDA
 * event per bucket on its stream as soon as the bucket is final; gt_grad_bucket_wait(b, s) makes stream `s` (the
 * caller's communication stream) wait for bucket b of the most recent call, so the caller can launch that bucket's
 * all-reduce while the layers below are still in backward.  gt_grad_events_enable(0) frees the events. */
int gt_grad_buckets(const gt_config *cfg, int64_t *offsets, int64_t *sizes, int max_entries);
int gt_grad_events_enable(int max_buckets);
int gt_grad_bucket_wait(int bucket, void *stream);

/* Launch accounting for bench.py.  gt_launch_count(-1) = kernels launched by this library so far
 * (all classes); gt_profile_enable(class, max_records) brackets every launch of one kernel class with
 * CUDA events on its stream (0 disables); gt_profile_collect sums their elapsed times and resets.
 * Kernel classes: 1 gemm_f32, 2 attention fwd, 3 attention bwd, 4 layernorm, 5 element-wise, 6 loss,
 * 7 optimizer, 16 tc weight prep, 17 tc layer fwd, 18 tc layer bwd, 19 tc head, 20 tc wgrad, 21 tc input,
 * 22 gemm_tc (generic tcgen05 GEMM). */
int64_t gt_launch_count(int kernel_class);
int     gt_profile_enable(int kernel_class, int max_records);
int     gt_profile_collect(double *total_ms, int64_t *launches);
/* gt_profile_enable(-1, n) brackets EVERY class; gt_profile_collect_class reads one class's total without resetting
 * (call it before gt_profile_collect, which resets). */
int     gt_profile_collect_class(int kernel_class, double *total_ms, int64_t *launches);

/* Test hook: fills keep[i] = 1/0 for element indices idx0..idx0+n-1 of dropout site `site`
 * (the generator restated in oracle/groove_oracle.py:dropout_keep). */
int gt_debug_dropout_mask(uint64_t seed, uint64_t step, int32_t site, float p,
                          int64_t idx0, int64_t n, uint8_t *keep, void *stream);

/* Stand-alone tcgen05 tile GEMM used by the unit tests of the tensor-core engine:
 * D[M,N] = A[M,K] (bf16 bits) * B[N,K]^T (bf16 bits), fp32 out; M multiple of 128. */
int gt_debug_tc_gemm(const uint16_t *a, const uint16_t *b, float *d, int m, int n, int k,
                     int variant, void *stream);

/* Test hook for the generic GEMMs: C[m,n] = epi(sum_k A[m*sam + k*sak] * B[n*sbn + k*sbk]) on fp32 device buffers.
 * tc == 1 runs the tcgen05 kernel (operands rounded to bf16, fp32 accumulate), tc == 2 the same kernel in split form
 * (GT_PREC_FP32_TC: three-term bf16 split operands, fp32 results; needs a bound scratch, falls back to the SIMT
 * kernel without one), tc == 0 the fp32 SIMT kernel.
 * flags: bit0 relu, bit1 accumulate (C += v), bit2 atomic (required when split_k_chunk splits K).  bias[N], residual
 * (ld_res) and mask_pos (ld_mask, mask_scale) may be NULL; drop_p > 0 applies the counter-based dropout of site
 * `site` at (seed, step) to element (row0 + m) * N + n. */
int gt_debug_gemm(int tc, const float *a, int64_t sam, int64_t sak, const float *b, int64_t sbn, int64_t sbk,
                  float *c, int64_t ldc, int64_t m, int64_t n, int64_t k, int flags, const float *bias,
                  const float *residual, int64_t ld_res, const float *mask_pos, int64_t ld_mask, float mask_scale,
                  float drop_p, uint64_t seed, uint64_t step, int32_t site, int64_t row0, int64_t split_k_chunk,
                  void *stream);

/* Binds a caller-owned device buffer as the operand-image scratch of this thread's following gt_debug_gemm(tc != 0) calls
 * (the model passes carve the same scratch out of their workspace: the library allocates no device memory).  NULL / 0 unbinds:
 * the GEMM then stages its operands inside the main kernel. */
int gt_debug_gemm_scratch(void *scratch, int64_t bytes);

/* Micro-benchmark of the tensor pipe: n_mma back-to-back 128 x n x 16 bf16 UMMAs per SM from shared-memory
 * operands; out[0] (device float) = clocks per UMMA.  Used by tools/umma_rate.py. */
int gt_debug_umma_rate(int n, int n_mma, int ksteps, float *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* GROOVE_B200_H */
