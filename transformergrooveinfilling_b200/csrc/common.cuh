// common.cuh — shared declarations for the groove_b200 library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <stdlib.h>
#include <utility>

#include <nvtx3/nvToolsExt.h>

#include "../../include/groove_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "groove_b200 targets sm_100a (B200) only"
#endif

namespace gt {

constexpr int T = GT_T_STEPS;     // steps per groove
constexpr float LN_EPS = 1e-5f;   // torch.nn.LayerNorm default

// ---- error plumbing (no exceptions across the C ABI) ---------------------------------------
void set_error(const std::string &msg);
#define GT_FAIL(msg)                                                     \
  do {                                                                   \
    gt::set_error(std::string(__FILE__) + ":" + std::to_string(__LINE__) + ": " + (msg)); \
    return 1;                                                            \
  } while (0)
#define GT_CHECK(cond, msg) \
  do {                      \
    if (!(cond)) GT_FAIL(msg); \
  } while (0)
#define GT_CUDA(expr)                                                     \
  do {                                                                    \
    cudaError_t _e = (expr);                                              \
    if (_e != cudaSuccess) GT_FAIL(std::string(#expr) + " -> " + cudaGetErrorString(_e)); \
  } while (0)
#define GT_TRY(expr)        \
  do {                      \
    int _r = (expr);        \
    if (_r != 0) return _r; \
  } while (0)

// ---- launch accounting + optional per-kernel-class CUDA-event timing (bench.py roofline) --------
enum KernelClass {
  KC_NONE = 0, KC_GEMM_F32 = 1, KC_ATTN_FWD = 2, KC_ATTN_BWD = 3, KC_LN = 4, KC_ELEMWISE = 5, KC_LOSS = 6, KC_OPT = 7,
  KC_TC_PREP = 16, KC_TC_LAYER_FWD = 17, KC_TC_LAYER_BWD = 18, KC_TC_HEAD = 19, KC_TC_WGRAD = 20, KC_TC_INPUT = 21, KC_GEMM_TC = 22,
  KC_MAX = 32
};
// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------------------------
// A kernel launched through launch_pdl may become resident while its predecessor in the stream is still running; it must
// execute griddep_wait() before it touches anything the predecessor writes (a no-op for a plain launch), and lets ITS successor
// in with griddep_launch().  The fused layer kernels call both right after their prologue (TMEM allocation, barrier init,
// bulk-TMA load of the layer's weight images — data written at the start of the step, at least two kernels back): with a grid
// smaller than the GPU (the yaml batch sizes: 8 CTAs) the prologue of layer l + 1 runs on idle SMs while layer l computes.
// launch_dependents AFTER the wait keeps the chain one deep: when kernel l + 1 starts, kernel l - 1 has completed.  GT_PDL=0 disables.
#if defined(__CUDACC__)
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
inline bool pdl_enabled() {
  static const bool on = !(getenv("GT_PDL") && atoi(getenv("GT_PDL")) == 0);
  return on;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

struct LaunchScope {     // RAII: counts the launch; records start/stop events when its class is being profiled
  int cls; cudaStream_t st; int slot;
  LaunchScope(int cls, cudaStream_t st);
  ~LaunchScope();
};

// ---- NVTX ranges per phase (forward / loss / backward / optimizer ...): visible in nsys / ncu --nvtx timelines; a push / pop
// without an attached tool is a null function-pointer check (SURVEY.md §5 "tracing").  GT_NVTX=0 disables them.
bool nvtx_enabled();
struct NvtxRange {
  bool on;
  explicit NvtxRange(const char *name) : on(nvtx_enabled()) { if (on) nvtxRangePushA(name); }
  ~NvtxRange() { if (on) nvtxRangePop(); }
  NvtxRange(const NvtxRange &) = delete;
};
#define GT_NVTX(name) gt::NvtxRange _gt_nvtx_##__LINE__(name)

// ---- counter-based dropout generator (restated bit-exactly in oracle/groove_oracle.py) -----
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
  return x;
}
inline uint32_t site_key_seed(uint64_t seed) {                      // the seed-only prefix of site_key
  uint32_t k = mix32((uint32_t)seed ^ 0x9E3779B9u);
  return mix32(k ^ (uint32_t)(seed >> 32));
}
inline uint32_t site_key(uint64_t seed, uint64_t step, int32_t site) {
  uint32_t k = site_key_seed(seed);
  k = mix32(k + (uint32_t)step * 0x85EBCA6Bu);
  k = mix32(k ^ ((uint32_t)site * 0xC2B2AE35u));
  return k;
}
// A dropout decision is a 14-BIT field compared with a 14-bit threshold: keep <=> field >= thr, thr = round(p * 16384)
// (p quantised to 1 / 16384).  14 bits because two such fields sitting in the halves of a 32-bit word are then two
// non-negative, finite fp16 bit patterns whose fp16 order is their integer order: the hot kernels decide both with ONE packed
// half-precision compare (HSET2 -> 0xFFFF / 0 per half) and apply the result to a packed bf16 pair with one LOP3.
constexpr uint32_t DROP_FIELD_BITS = 14, DROP_FIELD_ONE = 1u << DROP_FIELD_BITS, DROP_FIELD_MASK2 = 0x3FFF3FFFu;
inline uint32_t drop_threshold(float p) {
  double t = (double)p * (double)DROP_FIELD_ONE;
  long r = lrint(t);            // round-half-even like Python's round()
  if (r < 0) r = 0;
  if (r > (long)DROP_FIELD_ONE - 1) r = (long)DROP_FIELD_ONE - 1;
  return (uint32_t)r;
}
inline float drop_scale(uint32_t thr) { return thr == 0 ? 1.0f : (float)((double)DROP_FIELD_ONE / ((double)DROP_FIELD_ONE - (double)thr)); }

struct Drop {           // one dropout site; thr == 0 means "inactive"
  uint32_t key = 0, thr = 0;
  float scale = 1.f;
  // CUDA-graph replay (gt_graph_train_create): the dropout step counter lives in device memory, so the key is derived ON THE
  // DEVICE from (k0 = seed part of site_key, *step_ptr, site) with exactly site_key's arithmetic — same masks as a normal call
  uint32_t k0 = 0, site = 0;
  const unsigned long long *step_ptr = nullptr;
};
__device__ __forceinline__ void drop_resolve(Drop &d);
__device__ __forceinline__ void drop_resolve(Drop &d) {
  if (d.step_ptr != nullptr) {
    const uint32_t k = mix32(d.k0 + (uint32_t)(*d.step_ptr) * 0x85EBCA6Bu);
    d.key = mix32(k ^ (d.site * 0xC2B2AE35u));
  }
}
// How a fused layer kernel sees its argument block (any struct with the four Drops d_attn, d1, d_ffn, d2).  DEVSTEP = false
// (every eager launch) binds the parameter block as it lies in constant memory; DEVSTEP = true (CUDA-graph replay,
// gt_graph_train_create) takes a copy whose four dropout keys are derived from the device-resident step counter.  Keeping the
// eager instantiation free of the copy matters: with it the d_model = 32 backward kernel was 3.3 % slower.
template <class A, bool DEVSTEP> struct DropArgsView;
template <class A> struct DropArgsView<A, false> {
  const A &a;
  __device__ __forceinline__ explicit DropArgsView(const A &p) : a(p) {}
};
template <class A> struct DropArgsView<A, true> {
  A a;
  __device__ __forceinline__ explicit DropArgsView(const A &p) : a(p) {
    drop_resolve(a.d_attn); drop_resolve(a.d1); drop_resolve(a.d_ffn); drop_resolve(a.d2);
  }
};
template <class A> inline bool drop_args_devstep(const A &a) {
  return a.d_attn.step_ptr != nullptr || a.d1.step_ptr != nullptr || a.d_ffn.step_ptr != nullptr || a.d2.step_ptr != nullptr;
}
// when non-null, the pass drivers build Drops that read the step from this device counter (set around graph capture only)
extern thread_local const unsigned long long *g_drop_step_ptr;
inline void drop_fill_devstep(Drop &d, uint64_t seed, int site) {
  if (g_drop_step_ptr != nullptr) { d.k0 = site_key_seed(seed); d.site = (uint32_t)site; d.step_ptr = g_drop_step_ptr; }
}
// Four 14-bit dropout fields (the low 14 bits of each 16-bit half of lo / hi) for the QUAD of elements 4w .. 4w+3 of one site.
// v = (uint32)w ^ ((uint32)(w >> 32) * 0x85EBCA6B).  One xorshift-multiply round, then two 32x32 -> 64 multiplies (IMAD.WIDE)
// whose halves are cross-mixed, then the field mask: 14 instructions per four decisions.  (Cheaper two-multiply variants were
// measured and rejected: adjacent quads of a sequential index correlate at 0.2 - 0.6; this one stays below 0.003.)
__host__ __device__ __forceinline__ void hash_quad(uint32_t v, uint32_t key, uint32_t &lo, uint32_t &hi) {
  uint32_t x = (v * 0x9E3779B1u) ^ key;
  x ^= x >> 16;
  const uint64_t p = (uint64_t)x * 0x7FEB352Du;
  const uint32_t plo = (uint32_t)p, phi = (uint32_t)(p >> 32);
  const uint32_t y = plo ^ phi;
  const uint64_t q = (uint64_t)y * 0x846CA68Bu;
  lo = ((uint32_t)q ^ phi) & DROP_FIELD_MASK2;
  hi = ((uint32_t)(q >> 32) ^ ((y << 16) | (y >> 16))) & DROP_FIELD_MASK2;
}
// decision k (0..3) of a quad: field k of (lo, hi) >= thr  <=> keep
__host__ __device__ __forceinline__ bool quad_keep(uint32_t lo, uint32_t hi, int k, uint32_t thr) {
  const uint32_t word = (k & 2) ? hi : lo;
  return ((k & 1) ? (word >> 16) : (word & 0xFFFFu)) >= thr;
}
__device__ __forceinline__ bool drop_keep(uint32_t key, uint32_t thr, uint64_t idx) {
  const uint64_t w = idx >> 2;
  uint32_t lo, hi;
  hash_quad((uint32_t)w ^ ((uint32_t)(w >> 32) * 0x85EBCA6Bu), key, lo, hi);
  return quad_keep(lo, hi, (int)(idx & 3), thr);
}
// Attention-probability sites index key j of a query row at position key_perm(j): the four probabilities an mma.sync
// accumulator thread owns for one row (keys 8 nt + 2 t + b, nt in {0,1} or {2,3}) then form ONE quad.
__host__ __device__ __forceinline__ int key_perm(int k) { return (k & 16) | (((k >> 1) & 3) << 2) | (((k >> 3) & 1) << 1) | (k & 1); }
__device__ __forceinline__ float ex2_ftz(float x) {      // 2^x, one MUFU (exp2f adds a denormal-range fix-up: 3 more instructions)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// site numbering (oracle: site_id / SITE_IN_*)
constexpr int SITE_IN_ENC = 1, SITE_IN_DEC = 2;
inline int site_id(int stack, int layer, int k) { return 16 + (stack * 64 + layer) * 8 + k; }

// ---- parameter layout -----------------------------------------------------------------------
struct AttnP { int64_t w_in, b_in, w_out, b_out; };
struct LayerP {
  AttnP sa, ca;                       // self-attention, cross-attention (decoder only)
  int64_t w1, b1, w2, b2;             // linear1 [F,d], linear2 [d,F]
  int64_t g1, be1, g2, be2, g3, be3;  // norm1..3 (norm3 decoder only)
};
struct Layout {
  int64_t in_enc_w, in_enc_b, in_dec_w, in_dec_b;
  LayerP enc[64], dec[64];
  int64_t enc_norm_g, enc_norm_b, dec_norm_g, dec_norm_b;
  int64_t out_w, out_b;
  int64_t total;
  int n_tensors;
  int64_t offs[2048], sizes[2048];
};
int build_layout(const gt_config &c, Layout &L);
int validate_config(const gt_config *cfg);

// ---- gradient buckets (data-parallel overlap) -------------------------------------------------
// The flat gradient is partitioned into contiguous ranges listed in the order in which backward FINISHES them
// (head first, input layer last).  Backward drivers call grad_bucket_ready(idx) right after enqueueing the last
// kernel that writes into bucket idx; when events are enabled (gt_grad_events_enable) that records a CUDA event a
// communication stream can wait on, so the all-reduce of a bucket overlaps the backward of the layers below it.
enum BucketKind { BK_HEAD = 0, BK_DEC_LAYER = 1, BK_MID = 2, BK_ENC_LAYER = 3, BK_IN_ENC = 4 };
int grad_bucket_index(const gt_config &c, int kind, int layer);
int grad_bucket_count(const gt_config &c);
void grad_bucket_ready(const gt_config &c, int kind, int layer, cudaStream_t st);

// ---- fp32 SIMT kernels (kernels_simt.cu) ----------------------------------------------------
struct GemmEpi {
  const float *bias = nullptr;       // + bias[n]
  int relu = 0;                      // max(.,0)
  const float *pe = nullptr;         // + pe[(m % 32) * N + n]
  Drop drop;                         // dropout on element ((row0 + m) * N + n)
  int64_t drop_row0 = 0;
  const float *mask_pos = nullptr;   // v = mask_pos[m*ld_mask+n] > 0 ? v * mask_scale : 0
  int64_t ld_mask = 0;
  float mask_scale = 1.f;
  const float *residual = nullptr;   // + residual[m*ld_res+n]
  int64_t ld_res = 0;
  int accumulate = 0;                // C += v
  int atomic = 0;                    // atomicAdd(C, v)   (split-K)
};
// C[m,n] = epi( sum_k A(m,k) * B(n,k) ), A(m,k) = A[m*sam + k*sak], B(n,k) = B[n*sbn + k*sbk]
int gemm_f32(const float *A, int64_t sam, int64_t sak, const float *B, int64_t sbn, int64_t sbk,
             float *C, int64_t ldc, int64_t M, int64_t N, int64_t K, const GemmEpi &epi,
             int64_t split_k_chunk, cudaStream_t st);
// Same contract on the tensor cores (gemm_tc.cu): operands rounded to bf16 while they are staged, fp32 accumulation in
// TMEM, fp32 result.  gemm_tc_supported: each operand has a contiguous index and M, N, K >= 32.
bool gemm_tc_supported(int64_t sam, int64_t sak, int64_t sbn, int64_t sbk, int64_t M, int64_t N, int64_t K);
int gemm_tc(const float *A, int64_t sam, int64_t sak, const float *B, int64_t sbn, int64_t sbk,
            float *C, int64_t ldc, int64_t M, int64_t N, int64_t K, const GemmEpi &epi,
            int64_t split_k_chunk, cudaStream_t st, int split = 0);
// split = 1 (GT_PREC_FP32_TC): every operand is imaged as a bf16 triple x0 + x1 + x2 (exact) and contracted in six passes — fp32
// results on the tensor cores; without a bound scratch the problem runs on gemm_f32.
// operand-image scratch of gemm_tc: caller-owned (a workspace region), bound to the calling thread for the duration of a pass
void gemm_tc_bind_scratch(void *p, int64_t bytes);
int64_t gemm_tc_scratch_bytes(int64_t tokens, int64_t d, int64_t F, int split = 0);
// out[n] += sum_m X[m*ld + n]
int colsum_f32(const float *X, int64_t ld, int64_t M, int N, float *out, cudaStream_t st);

struct AttnArgs {
  const float *q, *k, *v;  int64_t ldq, ldk, ldv;
  float *o;                int64_t ldo;
  const float *d_o;        int64_t ld_do;     // backward only
  float *dq, *dk, *dv;     int64_t ld_dq, ld_dk, ld_dv;
  int64_t n_seq; int H, dh, causal;
  Drop drop; int64_t seq0;
};
int attention_fwd(const AttnArgs &a, cudaStream_t st);
int attention_bwd(const AttnArgs &a, cudaStream_t st);
// the same contract on mma.sync with bf16 operands (attn_mma.cu): head dims 16 / 32 / 64 / 128, even leading dimensions
bool attention_tc_supported(const AttnArgs &a);
int attention_fwd_tc(const AttnArgs &a, cudaStream_t st);
int attention_bwd_tc(const AttnArgs &a, cudaStream_t st);

// y = LN(res + dropout(a)) ; u = res + dropout(a) saved when u != nullptr
int ln_fwd(const float *a, const float *res, const float *gamma, const float *beta, float *u, float *y,
           float *mean, float *rstd, int64_t M, int d, const Drop &drop, int64_t row0, cudaStream_t st);
// du = dLN/du ; da = du * dropmask (only written when da != nullptr) ; dgamma/dbeta accumulated
int ln_bwd(const float *dy, const float *u, const float *mean, const float *rstd, const float *gamma,
           float *du, float *da, float *dgamma, float *dbeta, int64_t M, int d, const Drop &drop,
           int64_t row0, cudaStream_t st);
// x0 = dropout(r + pe)
int pe_dropout_fwd(const float *r, const float *pe, float *x0, int64_t M, int d, const Drop &drop,
                   int64_t row0, cudaStream_t st);
// g = dx0 * dropmask * (r > 0)
int pe_dropout_bwd(const float *dx0, const float *r, float *g, int64_t M, int d, const Drop &drop,
                   int64_t row0, cudaStream_t st);
// in place on logits [M,27]: ch 9..17 sigmoid, 18..26 0.5*tanh ; thres >= 0: ch 0..8 -> (sigmoid > thres)
int head_activation(float *hvo, int64_t M, int e_tgt, float thres, cudaStream_t st);
// dlogits = d_hvo * activation'(hvo)
int head_activation_bwd(const float *d_hvo, const float *hvo, float *dlogits, int64_t M, int e_tgt, cudaStream_t st);
int loss_fwd_bwd(const float *hvo, const float *y, int64_t n_seq, float penalty, float *metrics6, float *d_hvo,
                 float grad_scale, float *partials, cudaStream_t st, int n_voices = 9);
int64_t loss_scratch_floats(int64_t n_seq);
int loss_finalize(const float *partials, int64_t blocks, int64_t M, float *metrics6, cudaStream_t st);
// fused ends of the d_model = 32 path (edge32.cu): input layer + positional encoding + dropout ; final LayerNorm + output
// head + activations (+ calculate_loss) ; and their backward passes, one HBM-bound kernel each
int64_t edge32_loss_partials(int64_t n_seq);
int edge32_stem_fwd(const float *src, int E, const float *W, const float *b, const float *pe, float *x0, int64_t M, const Drop &drop,
                    int64_t row0, cudaStream_t st);
int edge32_stem_bwd(const float *dx0, const float *src, int E, const float *W, const float *b, float *gW, float *gb, int64_t M,
                    const Drop &drop, int64_t row0, cudaStream_t st);
int edge32_tail_fwd(const float *x, const float *gamma, const float *beta, const float *Wout, const float *bout, float *hvo,
                    float *mean, float *rstd, int64_t M, float thres, cudaStream_t st);
int edge32_tail_fwd_loss(const float *x, const float *gamma, const float *beta, const float *Wout, const float *bout, float *hvo,
                         float *mean, float *rstd, int64_t M, const float *y, float penalty, float *dlog, float *partials,
                         float *metrics6, cudaStream_t st);
int edge32_tail_bwd(const float *d_in, const float *hvo, const float *x, const float *mean, const float *rstd, const float *gamma,
                    const float *beta, const float *Wout, float *dx, float *gW, float *gb, float *gg, float *gbe, int64_t M,
                    cudaStream_t st);
// per-voice hit accuracy / velocity MSE / micro-timing MSE (+ Overall) of predictions against ground truth [n_seq,32,3V]
int64_t eval_scratch_floats(int64_t n_seq, int n_voices);
int eval_metrics(const float *pred, const float *gt, int64_t n_seq, int n_voices, float *out, float *partials, cudaStream_t st);
int shift_right(const float *y, float *out, int64_t n_seq, int e, cudaStream_t st);
int gather_rows(const float *data, const int64_t *perm, int64_t start, float *out, int64_t n_rows, int64_t row_floats, cudaStream_t st,
                const unsigned long long *start_ptr = nullptr);      // start_ptr: row offset += *start_ptr (graph replay)
int sgd_step(float *p, const float *g, int64_t n, float lr, float gs, cudaStream_t st);
int adam_step(float *p, const float *g, float *m, float *v, int64_t n, float lr, float b1, float b2, float eps,
              int64_t step, float gs, cudaStream_t st, const unsigned long long *t_ptr = nullptr);   // t_ptr: 1-based step = *t_ptr + 1 (graph replay)
int counter_advance(unsigned long long *counters, int64_t batch, const float *metrics6, float *ring, int64_t ring_slots, cudaStream_t st);
int optimizer_kernels_warm();
int debug_dropout_mask(uint32_t key, uint32_t thr, int64_t idx0, int64_t n, uint8_t *keep, cudaStream_t st);
int attention_decode(const float *q, int64_t ldq, const float *k, const float *v, int64_t ld_key, int64_t ld_seq, int nkeys,
                     float *o, int64_t ldo, int64_t n_seq, int H, int dh, cudaStream_t st);
int add_pe_row(float *x, const float *pe_row, int64_t n_rows, int d, cudaStream_t st);
int decode_feedback(const float *hvo_step, float *tok, float *out, int64_t n_seq, int e, int step_i, float thres, cudaStream_t st);
// one decoder layer of the KV-cached decode for one token per sequence, d_model = 32 (decode32.cu)
bool dec32_supported(const gt_config &c);
int dec32_layer_step(const gt_config &c, const LayerP &p, const float *P, const float *y_in, float *y_out, float *kv_self,
                     const float *kv_cross, int64_t n, int step, bool do_ffn, cudaStream_t st);
int predict_feedback(const float *hvo, float *tgt, float *out, int64_t n_seq, int e, int step_i, float thres, cudaStream_t st);

}  // namespace gt
