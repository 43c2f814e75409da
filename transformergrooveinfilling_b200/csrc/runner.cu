// runner.cu — parameter layout, workspace planning and the fp32 pass orchestration behind the
// C ABI (include/groove_b200.h).  One C call = one whole pass (all layers), so Python pays one
// ctypes call per forward / backward / fused train step.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "tc_path.cuh"
#include "tc256.cuh"

namespace gt {

static thread_local std::string g_err;
thread_local const unsigned long long *g_drop_step_ptr = nullptr;
void set_error(const std::string &msg) { g_err = msg; }
bool nvtx_enabled() {
  static const bool on = !(getenv("GT_NVTX") && atoi(getenv("GT_NVTX")) == 0);
  return on;
}

// ---- launch accounting ------------------------------------------------------------------------
static int64_t g_launches[KC_MAX] = {0};
// Profiling brackets and gradient-bucket events belong to the host thread that enabled them: sweep members drive the library
// from many host threads at once (sweep.py), and one thread's bench / data-parallel bookkeeping must not see another's launches.
static thread_local int g_prof_class = KC_NONE;
static thread_local std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_events;
static thread_local std::vector<int> g_prof_cls;
static thread_local size_t g_prof_used = 0;

LaunchScope::LaunchScope(int c, cudaStream_t s) : cls(c), st(s), slot(-1) {
  __atomic_fetch_add(&g_launches[c], 1, __ATOMIC_RELAXED);     // members of a sweep launch from several host threads
  if ((c == g_prof_class || g_prof_class == -1) && g_prof_used < g_prof_events.size()) {
    slot = (int)g_prof_used++;
    g_prof_cls[slot] = c;
    cudaEventRecord(g_prof_events[slot].first, st);
  }
}
LaunchScope::~LaunchScope() {
  if (slot >= 0) cudaEventRecord(g_prof_events[slot].second, st);
}

// ---- gradient buckets ---------------------------------------------------------------------------
static thread_local std::vector<cudaEvent_t> g_bucket_events;
static thread_local bool g_bucket_events_on = false;

int grad_bucket_count(const gt_config &c) { return c.n_dec > 0 ? c.n_enc + c.n_dec + 3 : c.n_enc + 2; }

int grad_bucket_index(const gt_config &c, int kind, int layer) {
  const int after_dec = c.n_dec > 0 ? c.n_dec + 2 : 1;     // index of the last encoder layer's bucket
  switch (kind) {
    case BK_HEAD: return 0;
    case BK_DEC_LAYER: return 1 + (c.n_dec - 1 - layer);
    case BK_MID: return c.n_dec + 1;
    case BK_ENC_LAYER: return after_dec + (c.n_enc - 1 - layer);
    default: return after_dec + c.n_enc;
  }
}

void grad_bucket_ready(const gt_config &c, int kind, int layer, cudaStream_t st) {
  if (!g_bucket_events_on) return;
  const int i = grad_bucket_index(c, kind, layer);
  if (i >= 0 && i < (int)g_bucket_events.size()) cudaEventRecord(g_bucket_events[i], st);
}

// [offset, offset+size) of every bucket in completion order: a descending partition of the flat vector
static int bucket_ranges(const gt_config &c, const Layout &L, int64_t *offs, int64_t *sizes) {
  std::vector<int64_t> starts;
  if (c.n_dec > 0) {
    starts.push_back(L.dec_norm_g);
    for (int l = c.n_dec - 1; l >= 0; --l) starts.push_back(L.dec[l].sa.w_in);
    starts.push_back(L.enc_norm_g);
  } else {
    starts.push_back(L.enc_norm_g);
  }
  for (int l = c.n_enc - 1; l >= 0; --l) starts.push_back(L.enc[l].sa.w_in);
  starts.push_back(L.in_enc_w);
  int64_t end = L.total;
  for (size_t i = 0; i < starts.size(); ++i) { offs[i] = starts[i]; sizes[i] = end - starts[i]; end = starts[i]; }
  return (int)starts.size();
}

int validate_config(const gt_config *c) {
  GT_CHECK(c != nullptr, "null config");
  GT_CHECK(c->d_model >= 1 && c->d_model <= 512, "d_model must be in [1,512]");
  GT_CHECK(c->nhead >= 1 && c->d_model % c->nhead == 0, "d_model must be divisible by nhead");
  GT_CHECK(c->dim_ff >= 1 && c->dim_ff <= 8192, "dim_feedforward must be in [1,8192]");
  GT_CHECK(c->n_enc >= 1 && c->n_enc <= 64, "num_encoder_layers must be in [1,64]");
  GT_CHECK(c->n_dec >= 0 && c->n_dec <= 64, "num_decoder_layers must be in [0,64]");
  GT_CHECK(c->e_src >= 1 && c->e_src <= 512, "embedding_size_src out of range");
  // BGT/models/io_layers.py:34-40 / train.py:12-13 split the target embedding into thirds (hits | velocities | offsets)
  GT_CHECK(c->e_tgt >= 3 && c->e_tgt <= 192 && c->e_tgt % 3 == 0,
           "embedding_size_tgt must be a multiple of 3 in [3, 192] (n_voices x hit / velocity / offset; the reference's sets use 27)");
  GT_CHECK(c->dropout >= 0.f && c->dropout < 1.f, "dropout must be in [0,1)");
  GT_CHECK(c->precision == GT_PREC_FP32 || c->precision == GT_PREC_BF16 || c->precision == GT_PREC_FP32_TC, "unknown precision mode");
  return 0;
}

int build_layout(const gt_config &c, Layout &L) {
  memset(&L, 0, sizeof(L));
  int64_t off = 0;
  int nt = 0;
  auto take = [&](int64_t n) {
    int64_t o = off;
    L.offs[nt] = o; L.sizes[nt] = n; ++nt;
    off += (n + 3) / 4 * 4;              // 16-byte alignment of every tensor
    return o;
  };
  const int64_t d = c.d_model, F = c.dim_ff;
  auto attn = [&](AttnP &a) { a.w_in = take(3 * d * d); a.b_in = take(3 * d); a.w_out = take(d * d); a.b_out = take(d); };
  L.in_enc_w = take(d * c.e_src); L.in_enc_b = take(d);
  for (int l = 0; l < c.n_enc; ++l) {
    LayerP &p = L.enc[l];
    attn(p.sa);
    p.w1 = take(F * d); p.b1 = take(F); p.w2 = take(d * F); p.b2 = take(d);
    p.g1 = take(d); p.be1 = take(d); p.g2 = take(d); p.be2 = take(d);
  }
  L.enc_norm_g = take(d); L.enc_norm_b = take(d);
  if (c.n_dec > 0) {
    L.in_dec_w = take(d * c.e_tgt); L.in_dec_b = take(d);
    for (int l = 0; l < c.n_dec; ++l) {
      LayerP &p = L.dec[l];
      attn(p.sa); attn(p.ca);
      p.w1 = take(F * d); p.b1 = take(F); p.w2 = take(d * F); p.b2 = take(d);
      p.g1 = take(d); p.be1 = take(d); p.g2 = take(d); p.be2 = take(d); p.g3 = take(d); p.be3 = take(d);
    }
    L.dec_norm_g = take(d); L.dec_norm_b = take(d);
  }
  L.out_w = take(c.e_tgt * d); L.out_b = take(c.e_tgt);
  L.total = off;
  L.n_tensors = nt;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// workspace plan
// ---------------------------------------------------------------------------------------------
struct LayerBuf {
  float *qkv, *ctx, *u1, *m1, *r1, *x1, *hd, *u2, *m2, *r2, *x2;
  float *qc, *kvc, *ctx2, *u3, *m3, *r3, *x3;   // decoder only (x2 = after cross-attn, x3 = layer output)
};
struct Plan {
  float *r0e, *x0e, *r0d, *y0d;
  LayerBuf enc[64], dec[64];
  float *mf_e, *rf_e, *mem;          // encoder final LN
  float *mf_d, *rf_d, *zdec;         // decoder final LN
  float *a;                          // GEMM output feeding an LN  [M,d]
  float *tgt_in;                     // shifted target (fused train step / predict)
  float *full;                       // [M,27] full-pass outputs of one predict() iteration
  float *d_hvo, *loss_partials;      // fused train step
  // backward temporaries
  float *dxa, *dxb, *du, *da, *dh, *dqkv, *dctx, *dlog, *dmem, *dqc, *dkvc, *g0;
  // KV-cached decode (mode 2): per-layer key/value caches + one-token-per-sequence step buffers
  float *kv_self[64], *kv_cross[64];
  uint8_t *enc_img;                  // hybrid encoder (fused d_model = 32 layer kernels inside an encoder-decoder model): weight images
  uint8_t *dec_img;                  // hybrid decoder: W1 / W2 images of the fused feed-forward blocks
  float *s_tok, *s_ya, *s_yb, *s_x1, *s_x2, *s_q, *s_ctx, *s_a, *s_hd, *s_z, *s_hvo;
  uint8_t *gemm_img;                 // bf16 mode: operand images of the generic tcgen05 GEMM (gemm_tc.cu), caller-owned like everything else
  int64_t gemm_img_bytes;
  int64_t bytes;
};

static void make_plan(const gt_config &c, int64_t n_seq, int mode, char *base, Plan &P) {
  memset(&P, 0, sizeof(P));
  const int64_t M = n_seq * T, d = c.d_model, F = c.dim_ff;
  int64_t off = 0;
  auto take = [&](int64_t n) -> float * {
    int64_t o = off;
    off += (n * 4 + 255) / 256 * 256;
    return base ? reinterpret_cast<float *>(base + o) : nullptr;
  };
  const bool train = mode == 1;
  const bool decode = mode == 2 && c.n_dec > 0;
  const bool hybrid = c.n_dec > 0 && tc_encoder_supported(c);
  auto layer_slim = [&](LayerBuf &b) {                // hybrid encoder layer: the fused kernels keep only u1 / u2 and the output
    b.u1 = train ? take(M * d) : nullptr; b.u2 = train ? take(M * d) : nullptr; b.x2 = take(M * d);
  };
  auto layer = [&](LayerBuf &b, bool dec) {
    if (!dec && hybrid) return layer_slim(b);
    b.qkv = take(M * 3 * d); b.ctx = take(M * d);
    b.u1 = train ? take(M * d) : nullptr; b.m1 = train ? take(M) : nullptr; b.r1 = train ? take(M) : nullptr;
    b.x1 = take(M * d); b.hd = (dec && hybrid) ? nullptr : take(M * F);      // hybrid decoder: the hidden activations never leave the SM
    b.u2 = train ? take(M * d) : nullptr; b.m2 = train ? take(M) : nullptr; b.r2 = train ? take(M) : nullptr;
    b.x2 = take(M * d);
    if (dec) {
      b.qc = take(M * d); b.kvc = take(M * 2 * d); b.ctx2 = take(M * d);
      b.u3 = train ? take(M * d) : nullptr; b.m3 = train ? take(M) : nullptr; b.r3 = train ? take(M) : nullptr;
      b.x3 = take(M * d);
    }
  };
  P.r0e = take(M * d); P.x0e = take(M * d);
  if (train) {
    for (int l = 0; l < c.n_enc; ++l) layer(P.enc[l], false);
  } else {
    layer(P.enc[0], false);
    for (int l = 1; l < c.n_enc; ++l) P.enc[l] = P.enc[0];
    if (hybrid && c.n_enc > 1) {                      // the fused layer kernels ping-pong between two output buffers
      float *b2 = take(M * d);
      for (int l = 1; l < c.n_enc; l += 2) P.enc[l].x2 = b2;
    }
  }
  P.mf_e = train ? take(M) : nullptr; P.rf_e = train ? take(M) : nullptr; P.mem = take(M * d);
  if (hybrid) {
    P.enc_img = reinterpret_cast<uint8_t *>(take(((int64_t)tc_enc_img_stride(c) * c.n_enc + 3) / 4));
    P.dec_img = reinterpret_cast<uint8_t *>(take(((int64_t)tc_enc_img_stride(c) * 2 * c.n_dec + 3) / 4));   // decode: FFN blocks only
  }
  if (decode) {
    const int64_t n = n_seq;
    for (int l = 0; l < c.n_dec; ++l) { P.kv_self[l] = take(M * 2 * d); P.kv_cross[l] = take(M * 2 * d); }
    P.s_tok = take(n * c.e_tgt); P.s_ya = take(n * d); P.s_yb = take(n * d); P.s_x1 = take(n * d); P.s_x2 = take(n * d);
    P.s_q = take(n * d); P.s_ctx = take(n * d); P.s_a = take(n * d); P.s_hd = take(n * F); P.s_z = take(n * d);
    P.s_hvo = take(n * c.e_tgt);
  } else if (c.n_dec > 0) {
    P.r0d = take(M * d); P.y0d = take(M * d);
    if (train) {
      for (int l = 0; l < c.n_dec; ++l) layer(P.dec[l], true);
    } else {
      layer(P.dec[0], true);
      for (int l = 1; l < c.n_dec; ++l) P.dec[l] = P.dec[0];
    }
    P.mf_d = train ? take(M) : nullptr; P.rf_d = train ? take(M) : nullptr; P.zdec = take(M * d);
    P.tgt_in = take(M * c.e_tgt);
    if (!train) P.full = take(M * c.e_tgt);
  }
  P.a = take(M * d);
  if (train) {
    P.d_hvo = take(M * c.e_tgt);
    P.loss_partials = take(loss_scratch_floats(n_seq) + edge32_loss_partials(n_seq));
    P.dxa = take(M * d); P.dxb = take(M * d); P.du = take(M * d); P.da = take(M * d);
    P.dh = take(M * F); P.dqkv = take(M * 3 * d); P.dctx = take(M * d); P.dlog = take(M * c.e_tgt);
    P.g0 = take(M * d);
    if (c.n_dec > 0) { P.dmem = take(M * d); P.dqc = take(M * d); P.dkvc = take(M * 2 * d); }
  }
  if (c.precision == GT_PREC_FP32_TC || (c.precision == GT_PREC_BF16 && !(hybrid && tc_dec_attn_supported(c)))) {      // (every block fused: no generic GEMM of width >= 32 left)
    P.gemm_img_bytes = gemm_tc_scratch_bytes(M, d, F, c.precision == GT_PREC_FP32_TC);
    P.gemm_img = reinterpret_cast<uint8_t *>(take((P.gemm_img_bytes + 3) / 4));
  }
  P.bytes = off;
}

// ---------------------------------------------------------------------------------------------
// pass context + small helpers
// ---------------------------------------------------------------------------------------------
struct Ctx {
  gt_config c;
  Layout L;
  const float *P;
  float *G;
  const float *pe;
  int64_t n_seq, M;
  bool train;
  bool tc;                    // precision = bf16 on a shape without fused layer kernels: contractions run on gemm_tc
  bool split;                 // precision = fp32_tc: contractions on gemm_tc in split form (three bf16 terms per fp32 operand), everything else fp32
  uint64_t seed, step;
  int64_t seq0;
  cudaStream_t st;
  Drop drop(int site) const {
    Drop d;
    uint32_t thr = drop_threshold(c.dropout);
    if (!train || thr == 0) return d;
    d.thr = thr; d.key = site_key(seed, step, site); d.scale = drop_scale(thr);
    drop_fill_devstep(d, seed, site);
    return d;
  }
  int64_t row0() const { return seq0 * T; }
};

static const int64_t WGRAD_CHUNK = 2048;

// every contraction of the generic path goes through here: the tcgen05 GEMM when the pass runs in bf16 mode and the shape
// qualifies (gemm_tc_supported), the fp32 SIMT GEMM otherwise (fp32 mode; K = 16 / 27 input layers and the 27-wide head)
static int gemm(const Ctx &x, const float *A, int64_t sam, int64_t sak, const float *B, int64_t sbn, int64_t sbk, float *C,
                int64_t ldc, int64_t M, int64_t N, int64_t K, const GemmEpi &e, int64_t split_k_chunk) {
  if ((x.tc || x.split) && gemm_tc_supported(sam, sak, sbn, sbk, M, N, K))
    return gemm_tc(A, sam, sak, B, sbn, sbk, C, ldc, M, N, K, e, split_k_chunk, x.st, x.split ? 1 : 0);
  return gemm_f32(A, sam, sak, B, sbn, sbk, C, ldc, M, N, K, e, split_k_chunk, x.st);
}

// out[M,N] = epi(X[M,K] W[N,K]^T)
static int linear(const Ctx &x, const float *X, int64_t K, const float *W, float *out, int64_t N, GemmEpi e) {
  return gemm(x, X, K, 1, W, K, 1, out, N, x.M, N, K, e, 0);
}
// dX[M,K] = epi(dY[M,N] W[N,K])
static int linear_dgrad(const Ctx &x, const float *dY, int64_t N, const float *W, int64_t K, float *dX, GemmEpi e) {
  return gemm(x, dY, N, 1, W, 1, K, dX, K, x.M, K, N, e, 0);
}
// dW[N,K] += dY[M,N]^T X[M,K] ; db[N] += colsum(dY)
static int linear_wgrad(const Ctx &x, const float *dY, int64_t ldy, int64_t N, const float *X, int64_t ldx, int64_t K,
                        float *dW, float *db) {
  GemmEpi e; e.atomic = 1;
  // (fp32_tc: 512-token chunks — the tensor core's accumulator truncation grows with the UMMAs per accumulator, gemm_tc.cu)
  GT_TRY(gemm(x, dY, 1, ldy, X, 1, ldx, dW, K, N, K, x.M, e, x.split ? 512 : WGRAD_CHUNK));
  if (db) GT_TRY(colsum_f32(dY, ldy, x.M, (int)N, db, x.st));
  return 0;
}

static AttnArgs attn_args(const Ctx &x, const float *q, int64_t ldq, const float *k, const float *v, int64_t ldkv,
                          float *o, int causal, int site) {
  AttnArgs a;
  memset(&a, 0, sizeof(a));
  a.q = q; a.k = k; a.v = v; a.ldq = ldq; a.ldk = ldkv; a.ldv = ldkv;
  a.o = o; a.ldo = x.c.d_model;
  a.n_seq = x.n_seq; a.H = x.c.nhead; a.dh = x.c.d_model / x.c.nhead; a.causal = causal;
  a.drop = x.drop(site); a.seq0 = x.seq0;
  return a;
}

// bf16 mode: attention on mma.sync (attn_mma.cu) for head dims 16 / 32 / 64 / 128; fp32 mode and other head dims: the fp32 kernels
static int attn_fwd(const Ctx &x, const AttnArgs &a) {
  return (x.tc && attention_tc_supported(a)) ? attention_fwd_tc(a, x.st) : attention_fwd(a, x.st);
}
static int attn_bwd(const Ctx &x, const AttnArgs &a) {
  return (x.tc && attention_tc_supported(a)) ? attention_bwd_tc(a, x.st) : attention_bwd(a, x.st);
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
static int input_layer_fwd(const Ctx &x, const float *src, int E, int64_t w, int64_t b, float *r0, float *x0, int site) {
  GemmEpi e; e.bias = x.P + b; e.relu = 1;
  GT_TRY(linear(x, src, E, x.P + w, r0, x.c.d_model, e));
  return pe_dropout_fwd(r0, x.pe, x0, x.M, x.c.d_model, x.drop(site), x.row0(), x.st);
}

static int ffn_fwd(const Ctx &x, const LayerP &p, const float *xin, float *hd, float *a, int stack, int li) {
  const int d = x.c.d_model, F = x.c.dim_ff;
  GemmEpi e1; e1.bias = x.P + p.b1; e1.relu = 1; e1.drop = x.drop(site_id(stack, li, 2)); e1.drop_row0 = x.row0();
  GT_TRY(linear(x, xin, d, x.P + p.w1, hd, F, e1));
  GemmEpi e2; e2.bias = x.P + p.b2;
  return linear(x, hd, F, x.P + p.w2, a, d, e2);
}

static int enc_layer_fwd(const Ctx &x, const Plan &pl, int li, const float *xin) {
  const LayerP &p = x.L.enc[li];
  const LayerBuf &b = pl.enc[li];
  const int d = x.c.d_model;
  GemmEpi e; e.bias = x.P + p.sa.b_in;
  GT_TRY(linear(x, xin, d, x.P + p.sa.w_in, b.qkv, 3 * d, e));
  GT_TRY(attn_fwd(x, attn_args(x, b.qkv, 3 * d, b.qkv + d, b.qkv + 2 * d, 3 * d, b.ctx, 0, site_id(0, li, 0))));
  GemmEpi eo; eo.bias = x.P + p.sa.b_out;
  GT_TRY(linear(x, b.ctx, d, x.P + p.sa.w_out, pl.a, d, eo));
  GT_TRY(ln_fwd(pl.a, xin, x.P + p.g1, x.P + p.be1, b.u1, b.x1, b.m1, b.r1, x.M, d, x.drop(site_id(0, li, 1)), x.row0(), x.st));
  GT_TRY(ffn_fwd(x, p, b.x1, b.hd, pl.a, 0, li));
  return ln_fwd(pl.a, b.x1, x.P + p.g2, x.P + p.be2, b.u2, b.x2, b.m2, b.r2, x.M, d, x.drop(site_id(0, li, 3)), x.row0(), x.st);
}

static int dec_layer_fwd(const Ctx &x, const Plan &pl, int li, const float *yin) {
  const LayerP &p = x.L.dec[li];
  const LayerBuf &b = pl.dec[li];
  const int d = x.c.d_model;
  if (pl.dec_img != nullptr && tc_dec_attn_supported(x.c)) {
    // hybrid decoder layer = three fused blocks: causal self-attention + LN1, cross-attention + LN2, FFN + LN3
    GT_TRY(tc_dec_attn_fwd(x.c, x.L, x.P, pl.dec_img, li, 0, yin, nullptr, b.x1, b.u1, x.n_seq, x.train, x.seed, x.step, x.seq0, x.st));
    GT_TRY(tc_dec_attn_fwd(x.c, x.L, x.P, pl.dec_img, li, 1, b.x1, pl.mem, b.x2, b.u2, x.n_seq, x.train, x.seed, x.step, x.seq0, x.st));
    return tc_dec_ffn_fwd(x.c, x.L, x.P, pl.dec_img, li, b.x2, b.x3, b.u3, x.n_seq, x.train, x.seed, x.step, x.seq0, x.st);
  }
  GemmEpi e; e.bias = x.P + p.sa.b_in;
  GT_TRY(linear(x, yin, d, x.P + p.sa.w_in, b.qkv, 3 * d, e));
  GT_TRY(attn_fwd(x, attn_args(x, b.qkv, 3 * d, b.qkv + d, b.qkv + 2 * d, 3 * d, b.ctx, 1, site_id(1, li, 0))));
  GemmEpi eo; eo.bias = x.P + p.sa.b_out;
  GT_TRY(linear(x, b.ctx, d, x.P + p.sa.w_out, pl.a, d, eo));
  GT_TRY(ln_fwd(pl.a, yin, x.P + p.g1, x.P + p.be1, b.u1, b.x1, b.m1, b.r1, x.M, d, x.drop(site_id(1, li, 1)), x.row0(), x.st));
  // cross attention: q from the target stream, k/v from the encoder memory
  GemmEpi eq; eq.bias = x.P + p.ca.b_in;
  GT_TRY(linear(x, b.x1, d, x.P + p.ca.w_in, b.qc, d, eq));
  GemmEpi ekv; ekv.bias = x.P + p.ca.b_in + d;
  GT_TRY(linear(x, pl.mem, d, x.P + p.ca.w_in + (int64_t)d * d, b.kvc, 2 * d, ekv));
  GT_TRY(attn_fwd(x, attn_args(x, b.qc, d, b.kvc, b.kvc + d, 2 * d, b.ctx2, 0, site_id(1, li, 4))));
  GemmEpi eco; eco.bias = x.P + p.ca.b_out;
  GT_TRY(linear(x, b.ctx2, d, x.P + p.ca.w_out, pl.a, d, eco));
  GT_TRY(ln_fwd(pl.a, b.x1, x.P + p.g2, x.P + p.be2, b.u2, b.x2, b.m2, b.r2, x.M, d, x.drop(site_id(1, li, 5)), x.row0(), x.st));
  if (pl.dec_img != nullptr)       // hybrid: FFN + residual + LayerNorm3 in one fused kernel (TC_MODE_FFN)
    return tc_dec_ffn_fwd(x.c, x.L, x.P, pl.dec_img, li, b.x2, b.x3, b.u3, x.n_seq, x.train, x.seed, x.step, x.seq0, x.st);
  GT_TRY(ffn_fwd(x, p, b.x2, b.hd, pl.a, 1, li));
  return ln_fwd(pl.a, b.x2, x.P + p.g3, x.P + p.be3, b.u3, b.x3, b.m3, b.r3, x.M, d, x.drop(site_id(1, li, 3)), x.row0(), x.st);
}

static int encoder_fwd(const Ctx &x, const Plan &pl, const float *src) {
  const float *cur = pl.x0e;
  if (pl.enc_img != nullptr) {
    // hybrid: the encoder stack of an encoder-decoder model on the fused stem + fused tcgen05 layer kernels
    GT_TRY(tc_enc_prep(x.c, x.L, x.P, pl.enc_img, x.st));
    GT_TRY(edge32_stem_fwd(src, x.c.e_src, x.P + x.L.in_enc_w, x.P + x.L.in_enc_b, x.pe, pl.x0e, x.M, x.drop(SITE_IN_ENC), x.row0(),
                           x.st));
    for (int l = 0; l < x.c.n_enc; ++l) {
      GT_TRY(tc_enc_layer_fwd(x.c, x.L, x.P, pl.enc_img, l, cur, pl.enc[l].x2, pl.enc[l].u1, pl.enc[l].u2, x.n_seq, x.train, x.seed,
                              x.step, x.seq0, x.st));
      cur = pl.enc[l].x2;
    }
  } else {
    GT_TRY(input_layer_fwd(x, src, x.c.e_src, x.L.in_enc_w, x.L.in_enc_b, pl.r0e, pl.x0e, SITE_IN_ENC));
    for (int l = 0; l < x.c.n_enc; ++l) {
      GT_TRY(enc_layer_fwd(x, pl, l, cur));
      cur = pl.enc[l].x2;
    }
  }
  Drop none;
  return ln_fwd(cur, nullptr, x.P + x.L.enc_norm_g, x.P + x.L.enc_norm_b, nullptr, pl.mem, pl.mf_e, pl.rf_e, x.M,
                x.c.d_model, none, 0, x.st);
}

// fused ends of the hybrid decoder (edge32.cu): the target input layer and the final LayerNorm + head (+ loss)
static bool dec_fused_edges(const Ctx &x, const Plan &pl) { return pl.dec_img != nullptr && x.c.e_tgt == 27; }

static int decoder_fwd(const Ctx &x, const Plan &pl, const float *tgt_in, bool final_ln = true) {
  if (pl.dec_img != nullptr) GT_TRY(tc_dec_prep(x.c, x.L, x.P, pl.dec_img, x.st));
  if (dec_fused_edges(x, pl))
    GT_TRY(edge32_stem_fwd(tgt_in, x.c.e_tgt, x.P + x.L.in_dec_w, x.P + x.L.in_dec_b, x.pe, pl.y0d, x.M, x.drop(SITE_IN_DEC), x.row0(), x.st));
  else
    GT_TRY(input_layer_fwd(x, tgt_in, x.c.e_tgt, x.L.in_dec_w, x.L.in_dec_b, pl.r0d, pl.y0d, SITE_IN_DEC));
  const float *cur = pl.y0d;
  for (int l = 0; l < x.c.n_dec; ++l) {
    GT_TRY(dec_layer_fwd(x, pl, l, cur));
    cur = pl.dec[l].x3;
  }
  if (!final_ln) return 0;
  Drop none;
  return ln_fwd(cur, nullptr, x.P + x.L.dec_norm_g, x.P + x.L.dec_norm_b, nullptr, pl.zdec, pl.mf_d, pl.rf_d, x.M,
                x.c.d_model, none, 0, x.st);
}

static int head_fwd(const Ctx &x, const float *z, float *hvo, float thres) {
  GemmEpi e; e.bias = x.P + x.L.out_b;
  GT_TRY(linear(x, z, x.c.d_model, x.P + x.L.out_w, hvo, x.c.e_tgt, e));
  return head_activation(hvo, x.M, x.c.e_tgt, thres, x.st);
}

// y != nullptr (hybrid decoder with fused ends only): the tail also evaluates calculate_loss and leaves dL/dlogits in pl.dlog
static int forward_all(const Ctx &x, const Plan &pl, const float *src, const float *tgt_in, float *hvo, const float *y = nullptr,
                       float penalty = 0.f, float *metrics6 = nullptr) {
  GT_NVTX("groove.forward");
  GT_TRY(encoder_fwd(x, pl, src));
  if (x.c.n_dec > 0) {
    GT_CHECK(tgt_in != nullptr, "encoder-decoder forward needs the shifted target");
    if (dec_fused_edges(x, pl)) {
      GT_TRY(decoder_fwd(x, pl, tgt_in, false));
      const float *last = pl.dec[x.c.n_dec - 1].x3, *g = x.P + x.L.dec_norm_g, *b = x.P + x.L.dec_norm_b, *w = x.P + x.L.out_w, *bo = x.P + x.L.out_b;
      if (y != nullptr)
        return edge32_tail_fwd_loss(last, g, b, w, bo, hvo, pl.mf_d, pl.rf_d, x.M, y, penalty, pl.dlog, pl.loss_partials, metrics6, x.st);
      return edge32_tail_fwd(last, g, b, w, bo, hvo, pl.mf_d, pl.rf_d, x.M, -1.f, x.st);
    }
    GT_CHECK(y == nullptr, "forward_all: fused loss needs the fused decoder ends");
    GT_TRY(decoder_fwd(x, pl, tgt_in));
    return head_fwd(x, pl.zdec, hvo, -1.f);
  }
  GT_CHECK(y == nullptr, "forward_all: fused loss needs the fused decoder ends");
  return head_fwd(x, pl.mem, hvo, -1.f);
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
// Given du2-like gradient `dout` w.r.t. the FFN block output (post-LN), produce gradient w.r.t. the
// block input in `dx_out`.  xin = block input (saved), u/m/r = saved LN input + stats.
static int ffn_bwd(const Ctx &x, const Plan &pl, const LayerP &p, const float *dout, const float *xin, const float *hd,
                   const float *u, const float *m, const float *r, int64_t g, int64_t be, float *dx_out, int stack, int li) {
  const int d = x.c.d_model, F = x.c.dim_ff;
  GT_TRY(ln_bwd(dout, u, m, r, x.P + g, pl.du, pl.da, x.G + g, x.G + be, x.M, d, x.drop(site_id(stack, li, 3)), x.row0(), x.st));
  // dhd = da W2 ; dh = dhd * (hd>0) * scale
  GemmEpi e; e.mask_pos = hd; e.ld_mask = F; e.mask_scale = x.drop(site_id(stack, li, 2)).scale;
  GT_TRY(linear_dgrad(x, pl.da, d, x.P + p.w2, F, pl.dh, e));
  GT_TRY(linear_wgrad(x, pl.da, d, d, hd, F, F, x.G + p.w2, x.G + p.b2));
  GemmEpi e2; e2.residual = pl.du; e2.ld_res = d;
  GT_TRY(linear_dgrad(x, pl.dh, F, x.P + p.w1, d, dx_out, e2));
  return linear_wgrad(x, pl.dh, F, F, xin, d, d, x.G + p.w1, x.G + p.b1);
}

static int enc_layer_bwd(const Ctx &x, const Plan &pl, int li, const float *xin, const float *dout, float *dx_tmp, float *dx_in) {
  const LayerP &p = x.L.enc[li];
  const LayerBuf &b = pl.enc[li];
  const int d = x.c.d_model;
  GT_TRY(ffn_bwd(x, pl, p, dout, b.x1, b.hd, b.u2, b.m2, b.r2, p.g2, p.be2, dx_tmp, 0, li));
  GT_TRY(ln_bwd(dx_tmp, b.u1, b.m1, b.r1, x.P + p.g1, pl.du, pl.da, x.G + p.g1, x.G + p.be1, x.M, d,
                x.drop(site_id(0, li, 1)), x.row0(), x.st));
  GemmEpi e0;
  GT_TRY(linear_dgrad(x, pl.da, d, x.P + p.sa.w_out, d, pl.dctx, e0));
  GT_TRY(linear_wgrad(x, pl.da, d, d, b.ctx, d, d, x.G + p.sa.w_out, x.G + p.sa.b_out));
  AttnArgs a = attn_args(x, b.qkv, 3 * d, b.qkv + d, b.qkv + 2 * d, 3 * d, nullptr, 0, site_id(0, li, 0));
  a.d_o = pl.dctx; a.ld_do = d;
  a.dq = pl.dqkv; a.dk = pl.dqkv + d; a.dv = pl.dqkv + 2 * d; a.ld_dq = a.ld_dk = a.ld_dv = 3 * d;
  GT_TRY(attn_bwd(x, a));
  GemmEpi e1; e1.residual = pl.du; e1.ld_res = d;
  GT_TRY(linear_dgrad(x, pl.dqkv, 3 * d, x.P + p.sa.w_in, d, dx_in, e1));
  return linear_wgrad(x, pl.dqkv, 3 * d, 3 * d, xin, d, d, x.G + p.sa.w_in, x.G + p.sa.b_in);
}

static int dec_layer_bwd(const Ctx &x, const Plan &pl, int li, const float *yin, const float *dout, float *dx_tmp, float *dx_in) {
  const LayerP &p = x.L.dec[li];
  const LayerBuf &b = pl.dec[li];
  const int d = x.c.d_model;
  if (pl.dec_img != nullptr)
    GT_TRY(tc_dec_ffn_bwd(x.c, x.L, x.P, x.G, pl.dec_img, li, b.x2, b.u3, dout, dx_tmp, x.n_seq, x.seed, x.step, x.seq0, x.st));
  else
    GT_TRY(ffn_bwd(x, pl, p, dout, b.x2, b.hd, b.u3, b.m3, b.r3, p.g3, p.be3, dx_tmp, 1, li));
  if (pl.dec_img != nullptr && tc_dec_attn_supported(x.c)) {
    // fused cross-attention block: dx_tmp (grad wrt x2) -> dx_in (grad wrt x1), dmem += ; then the causal self-attention block
    GT_TRY(tc_dec_attn_bwd(x.c, x.L, x.P, x.G, pl.dec_img, li, 1, b.x1, pl.mem, b.u2, dx_tmp, dx_in, pl.dmem, x.n_seq, x.seed, x.step,
                           x.seq0, x.st));
    return tc_dec_attn_bwd(x.c, x.L, x.P, x.G, pl.dec_img, li, 0, yin, nullptr, b.u1, dx_in, dx_tmp, nullptr, x.n_seq, x.seed, x.step,
                           x.seq0, x.st);
  }
  // cross-attention block
  GT_TRY(ln_bwd(dx_tmp, b.u2, b.m2, b.r2, x.P + p.g2, pl.du, pl.da, x.G + p.g2, x.G + p.be2, x.M, d,
                x.drop(site_id(1, li, 5)), x.row0(), x.st));
  GemmEpi e0;
  GT_TRY(linear_dgrad(x, pl.da, d, x.P + p.ca.w_out, d, pl.dctx, e0));
  GT_TRY(linear_wgrad(x, pl.da, d, d, b.ctx2, d, d, x.G + p.ca.w_out, x.G + p.ca.b_out));
  AttnArgs c = attn_args(x, b.qc, d, b.kvc, b.kvc + d, 2 * d, nullptr, 0, site_id(1, li, 4));
  c.d_o = pl.dctx; c.ld_do = d;
  c.dq = pl.dqc; c.ld_dq = d; c.dk = pl.dkvc; c.dv = pl.dkvc + d; c.ld_dk = c.ld_dv = 2 * d;
  GT_TRY(attn_bwd(x, c));
  // d(x1) = du (residual) + dqc Wq ; dmem += dkvc Wkv
  GemmEpi e1; e1.residual = pl.du; e1.ld_res = d;
  GT_TRY(linear_dgrad(x, pl.dqc, d, x.P + p.ca.w_in, d, dx_in, e1));          // dx_in used as scratch for d(x1)
  GT_TRY(linear_wgrad(x, pl.dqc, d, d, b.x1, d, d, x.G + p.ca.w_in, x.G + p.ca.b_in));
  GemmEpi e2; e2.accumulate = 1;
  GT_TRY(linear_dgrad(x, pl.dkvc, 2 * d, x.P + p.ca.w_in + (int64_t)d * d, d, pl.dmem, e2));
  GT_TRY(linear_wgrad(x, pl.dkvc, 2 * d, 2 * d, pl.mem, d, d, x.G + p.ca.w_in + (int64_t)d * d, x.G + p.ca.b_in + d));
  // causal self-attention block
  GT_TRY(ln_bwd(dx_in, b.u1, b.m1, b.r1, x.P + p.g1, pl.du, pl.da, x.G + p.g1, x.G + p.be1, x.M, d,
                x.drop(site_id(1, li, 1)), x.row0(), x.st));
  GT_TRY(linear_dgrad(x, pl.da, d, x.P + p.sa.w_out, d, pl.dctx, e0));
  GT_TRY(linear_wgrad(x, pl.da, d, d, b.ctx, d, d, x.G + p.sa.w_out, x.G + p.sa.b_out));
  AttnArgs a = attn_args(x, b.qkv, 3 * d, b.qkv + d, b.qkv + 2 * d, 3 * d, nullptr, 1, site_id(1, li, 0));
  a.d_o = pl.dctx; a.ld_do = d;
  a.dq = pl.dqkv; a.dk = pl.dqkv + d; a.dv = pl.dqkv + 2 * d; a.ld_dq = a.ld_dk = a.ld_dv = 3 * d;
  GT_TRY(attn_bwd(x, a));
  GemmEpi e3; e3.residual = pl.du; e3.ld_res = d;
  GT_TRY(linear_dgrad(x, pl.dqkv, 3 * d, x.P + p.sa.w_in, d, dx_tmp, e3));
  GT_TRY(linear_wgrad(x, pl.dqkv, 3 * d, 3 * d, yin, d, d, x.G + p.sa.w_in, x.G + p.sa.b_in));
  return 0;   // result in dx_tmp
}

static int input_layer_bwd(const Ctx &x, const Plan &pl, const float *dx0, const float *r0, const float *src, int E,
                           int64_t w, int64_t b, int site) {
  GT_TRY(pe_dropout_bwd(dx0, r0, pl.g0, x.M, x.c.d_model, x.drop(site), x.row0(), x.st));
  return linear_wgrad(x, pl.g0, x.c.d_model, x.c.d_model, src, E, E, x.G + w, x.G + b);
}

static int backward_all(const Ctx &x, const Plan &pl, const float *src, const float *tgt_in, const float *hvo,
                        const float *d_hvo) {
  GT_NVTX("groove.backward");
  const int d = x.c.d_model, E = x.c.e_tgt;
  Drop none;
  const bool fused_dec_ends = x.c.n_dec > 0 && dec_fused_edges(x, pl);
  float *cur = pl.dxa, *oth = pl.dxb;
  if (!fused_dec_ends) {
    // head
    GT_CHECK(hvo != nullptr, "backward_all: hvo is required on this path");
    GT_TRY(head_activation_bwd(d_hvo, hvo, pl.dlog, x.M, E, x.st));
    const float *z = x.c.n_dec > 0 ? pl.zdec : pl.mem;
    GT_TRY(linear_wgrad(x, pl.dlog, E, E, z, d, d, x.G + x.L.out_w, x.G + x.L.out_b));
    GemmEpi e0;
    GT_TRY(linear_dgrad(x, pl.dlog, E, x.P + x.L.out_w, d, pl.dxa, e0));
  }
  if (x.c.n_dec > 0) {
    GT_CUDA(cudaMemsetAsync(pl.dmem, 0, (size_t)x.M * d * sizeof(float), x.st));
    const float *last = pl.dec[x.c.n_dec - 1].x3;
    if (fused_dec_ends) {
      // hvo == nullptr: d_hvo already holds dL/dlogits (left by the fused tail + loss forward)
      GT_TRY(edge32_tail_bwd(d_hvo, hvo, last, pl.mf_d, pl.rf_d, x.P + x.L.dec_norm_g, x.P + x.L.dec_norm_b, x.P + x.L.out_w, oth,
                             x.G + x.L.out_w, x.G + x.L.out_b, x.G + x.L.dec_norm_g, x.G + x.L.dec_norm_b, x.M, x.st));
    } else {
      GT_TRY(ln_bwd(cur, last, pl.mf_d, pl.rf_d, x.P + x.L.dec_norm_g, oth, nullptr, x.G + x.L.dec_norm_g,
                    x.G + x.L.dec_norm_b, x.M, d, none, 0, x.st));
    }
    grad_bucket_ready(x.c, BK_HEAD, 0, x.st);
    std::swap(cur, oth);
    for (int l = x.c.n_dec - 1; l >= 0; --l) {
      const float *yin = l == 0 ? pl.y0d : pl.dec[l - 1].x3;
      // reads `cur`; `oth` and pl.g0 are temporaries; the result lands in `oth`
      GT_TRY(dec_layer_bwd(x, pl, l, yin, cur, oth, pl.g0));
      grad_bucket_ready(x.c, BK_DEC_LAYER, l, x.st);
      std::swap(cur, oth);
    }
    if (fused_dec_ends)
      GT_TRY(edge32_stem_bwd(cur, tgt_in, E, x.P + x.L.in_dec_w, x.P + x.L.in_dec_b, x.G + x.L.in_dec_w, x.G + x.L.in_dec_b, x.M,
                             x.drop(SITE_IN_DEC), x.row0(), x.st));
    else
      GT_TRY(input_layer_bwd(x, pl, cur, pl.r0d, tgt_in, E, x.L.in_dec_w, x.L.in_dec_b, SITE_IN_DEC));
    // encoder gradient starts from dmem
    GT_CUDA(cudaMemcpyAsync(pl.dxa, pl.dmem, (size_t)x.M * d * sizeof(float), cudaMemcpyDeviceToDevice, x.st));
    cur = pl.dxa; oth = pl.dxb;
  }
  const float *last = pl.enc[x.c.n_enc - 1].x2;
  GT_TRY(ln_bwd(cur, last, pl.mf_e, pl.rf_e, x.P + x.L.enc_norm_g, oth, nullptr, x.G + x.L.enc_norm_g,
                x.G + x.L.enc_norm_b, x.M, d, none, 0, x.st));
  grad_bucket_ready(x.c, x.c.n_dec > 0 ? BK_MID : BK_HEAD, 0, x.st);
  std::swap(cur, oth);
  for (int l = x.c.n_enc - 1; l >= 0; --l) {
    const float *xin = l == 0 ? pl.x0e : pl.enc[l - 1].x2;
    if (pl.enc_img != nullptr)
      GT_TRY(tc_enc_layer_bwd(x.c, x.L, x.P, x.G, pl.enc_img, l, xin, pl.enc[l].u1, pl.enc[l].u2, cur, oth, x.n_seq, x.seed, x.step,
                              x.seq0, x.st));
    else
      GT_TRY(enc_layer_bwd(x, pl, l, xin, cur, pl.g0, oth));
    grad_bucket_ready(x.c, BK_ENC_LAYER, l, x.st);
    std::swap(cur, oth);
  }
  if (pl.enc_img != nullptr)
    GT_TRY(edge32_stem_bwd(cur, src, x.c.e_src, x.P + x.L.in_enc_w, x.P + x.L.in_enc_b, x.G + x.L.in_enc_w, x.G + x.L.in_enc_b, x.M,
                           x.drop(SITE_IN_ENC), x.row0(), x.st));
  else
    GT_TRY(input_layer_bwd(x, pl, cur, pl.r0e, src, x.c.e_src, x.L.in_enc_w, x.L.in_enc_b, SITE_IN_ENC));
  grad_bucket_ready(x.c, BK_IN_ENC, 0, x.st);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// KV-cached autoregressive predict (encoder-decoder).  BGT/models/transformer.py:48-83 runs 32 FULL decoder passes; the
// target mask is causal (BGT/models/utils.py:53-56), so position i of pass i depends only on tokens <= i and an incremental
// decode that caches each layer's self-attention keys/values (and the cross-attention keys/values of the encoder
// memory, which never change) produces the same outputs with 1/32 of the decoder work.
// ---------------------------------------------------------------------------------------------
static int linear_rows(const Ctx &x, const float *X, int64_t K, const float *W, float *out, int64_t ldc, int64_t N, int64_t rows,
                       GemmEpi e) {
  return gemm(x, X, K, 1, W, K, 1, out, ldc, rows, N, K, e, 0);
}

static int predict_decode(const Ctx &x, const Plan &pl, float thres, float *hvo_out) {
  const int d = x.c.d_model, F = x.c.dim_ff, E = x.c.e_tgt, H = x.c.nhead, dh = d / H;
  const int64_t n = x.n_seq;
  Drop none;
  for (int l = 0; l < x.c.n_dec; ++l) {        // cross-attention keys / values of the memory, once per layer
    const LayerP &p = x.L.dec[l];
    GemmEpi e; e.bias = x.P + p.ca.b_in + d;
    GT_TRY(linear(x, pl.mem, d, x.P + p.ca.w_in + (int64_t)d * d, pl.kv_cross[l], 2 * d, e));
  }
  GT_CUDA(cudaMemsetAsync(pl.s_tok, 0, (size_t)n * E * sizeof(float), x.st));
  if (pl.dec_img != nullptr && dec32_supported(x.c)) GT_TRY(tc_dec_prep(x.c, x.L, x.P, pl.dec_img, x.st));
  for (int i = 0; i < T; ++i) {
    GemmEpi ein; ein.bias = x.P + x.L.in_dec_b; ein.relu = 1;
    GT_TRY(linear_rows(x, pl.s_tok, E, x.P + x.L.in_dec_w, pl.s_ya, d, d, n, ein));
    GT_TRY(add_pe_row(pl.s_ya, x.pe + (int64_t)i * d, n, d, x.st));
    float *cur = pl.s_ya, *nxt = pl.s_yb;
    for (int l = 0; l < x.c.n_dec; ++l) {
      const LayerP &p = x.L.dec[l];
      if (dec32_supported(x.c)) {               // d_model = 32: the whole layer for this token in one kernel (decode32.cu) ...
        const bool tc_ffn = pl.dec_img != nullptr;   // ... bf16 mode: its FFN block on the tensor cores (TC_MODE_FFN over the n token rows)
        GT_TRY(dec32_layer_step(x.c, p, x.P, cur, tc_ffn ? pl.s_x2 : nxt, pl.kv_self[l], pl.kv_cross[l], n, i, !tc_ffn, x.st));
        if (tc_ffn) GT_TRY(tc_dec_ffn_rows(x.c, x.L, x.P, pl.dec_img, l, pl.s_x2, nxt, n, x.st));
        std::swap(cur, nxt);
        continue;
      }
      GemmEpi eq; eq.bias = x.P + p.sa.b_in;
      GT_TRY(linear_rows(x, cur, d, x.P + p.sa.w_in, pl.s_q, d, d, n, eq));
      GemmEpi ekv; ekv.bias = x.P + p.sa.b_in + d;        // this token's key | value -> row i of the layer's cache
      GT_TRY(linear_rows(x, cur, d, x.P + p.sa.w_in + (int64_t)d * d, pl.kv_self[l] + (int64_t)i * 2 * d, (int64_t)T * 2 * d, 2 * d, n, ekv));
      GT_TRY(attention_decode(pl.s_q, d, pl.kv_self[l], pl.kv_self[l] + d, 2 * d, (int64_t)T * 2 * d, i + 1, pl.s_ctx, d, n, H, dh, x.st));
      GemmEpi eo; eo.bias = x.P + p.sa.b_out;
      GT_TRY(linear_rows(x, pl.s_ctx, d, x.P + p.sa.w_out, pl.s_a, d, d, n, eo));
      GT_TRY(ln_fwd(pl.s_a, cur, x.P + p.g1, x.P + p.be1, nullptr, pl.s_x1, nullptr, nullptr, n, d, none, 0, x.st));
      GemmEpi ecq; ecq.bias = x.P + p.ca.b_in;
      GT_TRY(linear_rows(x, pl.s_x1, d, x.P + p.ca.w_in, pl.s_q, d, d, n, ecq));
      GT_TRY(attention_decode(pl.s_q, d, pl.kv_cross[l], pl.kv_cross[l] + d, 2 * d, (int64_t)T * 2 * d, T, pl.s_ctx, d, n, H, dh, x.st));
      GemmEpi eco; eco.bias = x.P + p.ca.b_out;
      GT_TRY(linear_rows(x, pl.s_ctx, d, x.P + p.ca.w_out, pl.s_a, d, d, n, eco));
      GT_TRY(ln_fwd(pl.s_a, pl.s_x1, x.P + p.g2, x.P + p.be2, nullptr, pl.s_x2, nullptr, nullptr, n, d, none, 0, x.st));
      GemmEpi e1; e1.bias = x.P + p.b1; e1.relu = 1;
      GT_TRY(linear_rows(x, pl.s_x2, d, x.P + p.w1, pl.s_hd, F, F, n, e1));
      GemmEpi e2; e2.bias = x.P + p.b2;
      GT_TRY(linear_rows(x, pl.s_hd, F, x.P + p.w2, pl.s_a, d, d, n, e2));
      GT_TRY(ln_fwd(pl.s_a, pl.s_x2, x.P + p.g3, x.P + p.be3, nullptr, nxt, nullptr, nullptr, n, d, none, 0, x.st));
      std::swap(cur, nxt);
    }
    GT_TRY(ln_fwd(cur, nullptr, x.P + x.L.dec_norm_g, x.P + x.L.dec_norm_b, nullptr, pl.s_z, nullptr, nullptr, n, d, none, 0, x.st));
    GemmEpi eh; eh.bias = x.P + x.L.out_b;
    GT_TRY(linear_rows(x, pl.s_z, d, x.P + x.L.out_w, pl.s_hvo, E, E, n, eh));
    GT_TRY(head_activation(pl.s_hvo, n, E, -1.f, x.st));
    GT_TRY(decode_feedback(pl.s_hvo, pl.s_tok, hvo_out, n, E, i, thres, x.st));
  }
  return 0;
}

static int make_ctx(Ctx &x, const gt_config *cfg, const float *params, float *grads, const float *pe, int64_t n_seq,
                    bool train, uint64_t seed, uint64_t step, int64_t seq0, void *stream) {
  GT_TRY(validate_config(cfg));
  GT_CHECK(n_seq >= 1, "n_seq must be >= 1");
  GT_CHECK(n_seq <= (int64_t)1 << 24, "n_seq too large for one call");
  GT_CHECK(params != nullptr && pe != nullptr, "null params / pe");
  GT_CHECK(((uintptr_t)params & 15) == 0 && ((uintptr_t)grads & 15) == 0, "params / grads must be 16-byte aligned");
  x.c = *cfg;
  GT_TRY(build_layout(*cfg, x.L));
  x.P = params; x.G = grads; x.pe = pe; x.n_seq = n_seq; x.M = n_seq * T; x.train = train;
  x.seed = seed; x.step = step; x.seq0 = seq0; x.st = (cudaStream_t)stream;
  x.tc = cfg->precision == GT_PREC_BF16;
  x.split = cfg->precision == GT_PREC_FP32_TC;
  return 0;
}

// precision = bf16 has two implementations: the fused tcgen05 layer kernels (tc_layers.cu / tc256*.cu) for the shapes they are
// instantiated for, and the generic path of this file with its contractions on gemm_tc for every other shape.
static bool fused_layers(const gt_config &c) { return c.precision == GT_PREC_BF16 && tc_shape_supported(c, nullptr); }

static int check_ws(const gt_config *cfg, int64_t n_seq, int mode, void *ws, int64_t ws_bytes, Plan &pl) {
  GT_CHECK(ws != nullptr, "null workspace");
  GT_CHECK(((uintptr_t)ws & 255) == 0, "workspace must be 256-byte aligned");
  make_plan(*cfg, n_seq, mode, (char *)ws, pl);
  GT_CHECK(ws_bytes >= pl.bytes, "workspace too small: need " + std::to_string(pl.bytes) + " bytes, got " + std::to_string(ws_bytes));
  gemm_tc_bind_scratch(pl.gemm_img, pl.gemm_img_bytes);      // this thread's GEMMs of the pass image their operands here
  return 0;
}

}  // namespace gt

// =============================================================================================
// C ABI
// =============================================================================================
using namespace gt;

extern "C" {

int gt_version(void) { return GT_ABI_VERSION; }
const char *gt_last_error(void) { return g_err.c_str(); }

int gt_path_kind(const gt_config *cfg) {
  if (validate_config(cfg)) return -1;
  if (cfg->precision == GT_PREC_FP32_TC) return GT_PATH_GEMM_TC_SPLIT;
  if (cfg->precision != GT_PREC_BF16) return GT_PATH_FP32_SIMT;
  // encoder-decoder, d_model = 32, head dim 2 / 4 / 8: every block of every layer runs in the fused tcgen05 kernels
  if (cfg->n_dec > 0 && tc_encoder_supported(*cfg) && tc_dec_attn_supported(*cfg) && cfg->e_tgt == 27) return GT_PATH_FUSED_D32;
  if (!tc_shape_supported(*cfg, nullptr)) return GT_PATH_GEMM_TC;
  return cfg->d_model == 256 ? GT_PATH_FUSED_D256 : GT_PATH_FUSED_D32;
}

int64_t gt_param_count(const gt_config *cfg) {
  if (validate_config(cfg)) return -1;
  static thread_local Layout L;
  build_layout(*cfg, L);
  return L.total;
}

int gt_param_layout(const gt_config *cfg, int64_t *offsets, int64_t *sizes, int max_entries) {
  if (validate_config(cfg)) return -1;
  static thread_local Layout L;
  build_layout(*cfg, L);
  if (max_entries < L.n_tensors) { set_error("gt_param_layout: max_entries too small"); return -1; }
  for (int i = 0; i < L.n_tensors; ++i) { offsets[i] = L.offs[i]; sizes[i] = L.sizes[i]; }
  return L.n_tensors;
}

int64_t gt_workspace_bytes(const gt_config *cfg, int64_t n_seq, int mode) {
  if (validate_config(cfg)) return -1;
  if (n_seq < 1 || mode < 0 || mode > 2) { set_error("gt_workspace_bytes: bad n_seq / mode"); return -1; }
  if (fused_layers(*cfg)) return tc_workspace_bytes(*cfg, n_seq, mode == 2 ? 0 : mode);
  static thread_local Plan pl;
  make_plan(*cfg, n_seq, mode, nullptr, pl);
  return pl.bytes;
}

int gt_forward(const gt_config *cfg, const float *params, const float *pe, const float *src, const float *tgt_in,
               int64_t n_seq, float *hvo, void *ws, int64_t ws_bytes, int train, uint64_t seed, uint64_t step,
               int64_t seq0, void *stream) {
  GT_NVTX("gt_forward");
  static thread_local Ctx x;
  GT_TRY(make_ctx(x, cfg, params, nullptr, pe, n_seq, train != 0, seed, step, seq0, stream));
  GT_CHECK(src != nullptr && hvo != nullptr, "null src / hvo");
  if (fused_layers(*cfg))
    return tc_forward(*cfg, x.L, params, pe, src, tgt_in, n_seq, hvo, ws, ws_bytes, train != 0, seed, step, seq0, x.st);
  static thread_local Plan pl;
  GT_TRY(check_ws(cfg, n_seq, train ? 1 : 0, ws, ws_bytes, pl));
  return forward_all(x, pl, src, tgt_in, hvo);
}

int gt_backward(const gt_config *cfg, const float *params, const float *pe, const float *src, const float *tgt_in,
                int64_t n_seq, const float *hvo, const float *d_hvo, float *grads, void *ws, int64_t ws_bytes, uint64_t seed,
                uint64_t step, int64_t seq0, void *stream) {
  GT_NVTX("gt_backward");
  static thread_local Ctx x;
  GT_TRY(make_ctx(x, cfg, params, grads, pe, n_seq, true, seed, step, seq0, stream));
  GT_CHECK(src != nullptr && hvo != nullptr && d_hvo != nullptr && grads != nullptr, "null src / hvo / d_hvo / grads");
  if (fused_layers(*cfg))
    return tc_backward(*cfg, x.L, params, pe, src, tgt_in, n_seq, hvo, d_hvo, grads, ws, ws_bytes, seed, step, seq0, x.st);
  static thread_local Plan pl;
  GT_TRY(check_ws(cfg, n_seq, 1, ws, ws_bytes, pl));
  return backward_all(x, pl, src, tgt_in, hvo, d_hvo);
}

int64_t gt_loss_scratch_floats(int64_t n_seq) { return loss_scratch_floats(n_seq); }

int gt_loss_voices(const float *hvo, const float *y, int64_t n_seq, int n_voices, float hit_loss_penalty, float *metrics6,
                   float *d_hvo, float grad_scale, float *partials, void *stream) {
  GT_CHECK(hvo && y && metrics6 && partials, "gt_loss: null pointer");
  GT_CHECK(n_seq >= 1, "gt_loss: empty batch");
  GT_CHECK(n_voices >= 1 && n_voices <= 64, "gt_loss: n_voices must be in [1, 64]");
  return loss_fwd_bwd(hvo, y, n_seq, hit_loss_penalty, metrics6, d_hvo, grad_scale, partials, (cudaStream_t)stream, n_voices);
}

int gt_loss(const float *hvo, const float *y, int64_t n_seq, float hit_loss_penalty, float *metrics6, float *d_hvo,
            float grad_scale, float *partials, void *stream) {
  return gt_loss_voices(hvo, y, n_seq, 9, hit_loss_penalty, metrics6, d_hvo, grad_scale, partials, stream);
}

int64_t gt_eval_scratch_floats(int64_t n_seq, int n_voices) {
  if (n_seq < 1 || n_voices < 1 || n_voices > 32) { set_error("gt_eval_scratch_floats: bad n_seq / n_voices"); return -1; }
  return eval_scratch_floats(n_seq, n_voices);
}

int gt_eval_metrics(const float *pred_hvo, const float *gt_hvo, int64_t n_seq, int n_voices, float *out, float *partials,
                    void *stream) {
  GT_CHECK(pred_hvo && gt_hvo && out && partials, "gt_eval_metrics: null pointer");
  GT_CHECK(n_seq >= 1, "gt_eval_metrics: empty batch");
  return eval_metrics(pred_hvo, gt_hvo, n_seq, n_voices, out, partials, (cudaStream_t)stream);
}

int gt_train_step(const gt_config *cfg, const float *params, const float *pe, const float *src, const float *y,
                  int64_t n_seq, float hit_loss_penalty, float *grads, float *metrics6, float *hvo, void *ws,
                  int64_t ws_bytes, uint64_t seed, uint64_t step, int64_t seq0, void *stream) {
  GT_NVTX("gt_train_step");
  static thread_local Ctx x;
  GT_TRY(make_ctx(x, cfg, params, grads, pe, n_seq, true, seed, step, seq0, stream));
  GT_CHECK(src && y && grads && metrics6 && hvo, "gt_train_step: null pointer");
  if (fused_layers(*cfg))
    return tc_train_step(*cfg, x.L, params, pe, src, y, n_seq, hit_loss_penalty, grads, metrics6, hvo, ws, ws_bytes, seed,
                         step, seq0, x.st);
  static thread_local Plan pl;
  GT_TRY(check_ws(cfg, n_seq, 1, ws, ws_bytes, pl));
  GT_CUDA(cudaMemsetAsync(grads, 0, (size_t)x.L.total * sizeof(float), x.st));
  const float *tgt_in = nullptr;
  if (cfg->n_dec > 0) {
    GT_TRY(shift_right(y, pl.tgt_in, n_seq, cfg->e_tgt, x.st));
    tgt_in = pl.tgt_in;
  }
  if (cfg->n_dec > 0 && dec_fused_edges(x, pl)) {
    GT_TRY(forward_all(x, pl, src, tgt_in, hvo, y, hit_loss_penalty, metrics6));
    return backward_all(x, pl, src, tgt_in, nullptr, pl.dlog);
  }
  GT_TRY(forward_all(x, pl, src, tgt_in, hvo));
  GT_TRY(loss_fwd_bwd(hvo, y, n_seq, hit_loss_penalty, metrics6, pl.d_hvo, 1.f, pl.loss_partials, x.st, cfg->e_tgt / 3));
  return backward_all(x, pl, src, tgt_in, hvo, pl.d_hvo);
}

int gt_train_steps(const gt_config *cfg, float *params, const float *pe, const float *data_x, const float *data_y,
                   const int64_t *perm, int64_t start, int64_t batch, int n_steps, float hit_loss_penalty, float *grads,
                   float *metrics_out, float *hvo, float *xbuf, float *ybuf, void *ws, int64_t ws_bytes, int optimizer,
                   float lr, float *m, float *v, int64_t adam_t0, uint64_t seed, uint64_t step0, void *stream) {
  GT_TRY(validate_config(cfg));
  GT_CHECK(data_x && data_y && perm && xbuf && ybuf && metrics_out && params && grads, "gt_train_steps: null pointer");
  GT_CHECK(batch >= 1 && n_steps >= 0 && start >= 0, "gt_train_steps: bad batch / n_steps / start");
  GT_CHECK(optimizer == 0 || (optimizer == 1 && m && v), "gt_train_steps: optimizer must be 0 (SGD) or 1 (Adam, with m and v)");
  static thread_local Layout L;
  GT_TRY(build_layout(*cfg, L));
  cudaStream_t st = (cudaStream_t)stream;
  for (int s = 0; s < n_steps; ++s) {
    GT_TRY(gather_rows(data_x, perm, start + (int64_t)s * batch, xbuf, batch, (int64_t)T * cfg->e_src, st));
    GT_TRY(gather_rows(data_y, perm, start + (int64_t)s * batch, ybuf, batch, (int64_t)T * cfg->e_tgt, st));
    GT_TRY(gt_train_step(cfg, params, pe, xbuf, ybuf, batch, hit_loss_penalty, grads, metrics_out + (int64_t)s * 6, hvo, ws, ws_bytes, seed,
                         step0 + (uint64_t)s, 0, stream));
    if (optimizer == 0) GT_TRY(sgd_step(params, grads, L.total, lr, 1.f, st));
    else GT_TRY(adam_step(params, grads, m, v, L.total, lr, 0.9f, 0.999f, 1e-8f, adam_t0 + s + 1, 1.f, st));
  }
  return 0;
}

struct GtGraph { cudaGraph_t graph; cudaGraphExec_t exec; };

int gt_graph_train_create(const gt_config *cfg, float *params, const float *pe, float *xbuf, float *ybuf, int64_t n_seq,
                          float hit_loss_penalty, float *grads, float *metrics6, float *hvo, void *ws, int64_t ws_bytes,
                          int optimizer, float lr, float *m, float *v, uint64_t seed, unsigned long long *counters,
                          const float *data_x, const float *data_y, const int64_t *perm, float *metrics_ring,
                          int64_t ring_slots, void *stream, void **graph_out) {
  GT_TRY(validate_config(cfg));
  GT_CHECK(graph_out && counters && xbuf && ybuf && params && grads && metrics6 && hvo, "gt_graph_train_create: null pointer");
  // every dropout-carrying kernel of every path derives its dropout keys from the device counter when asked to (drop_resolve /
  // DropArgsView): encoder-only and encoder-decoder models, fp32 and bf16, fused or per-op.  The fused d_model = 32 ends
  // (edge32.cu) are instantiated for the reference's two input widths.
  GT_CHECK(!fused_layers(*cfg) || cfg->d_model != 32 || ((cfg->e_src == 16 || cfg->e_src == 27) && cfg->e_tgt == 27),
           "gt_graph_train_create: the fused d_model = 32 path needs embedding_size_src 16 or 27");
  GT_CHECK(optimizer == 0 || (optimizer == 1 && m && v), "gt_graph_train_create: optimizer must be 0 (SGD) or 1 (Adam, with m and v)");
  GT_CHECK((data_x == nullptr) == (data_y == nullptr) && (data_x == nullptr) == (perm == nullptr),
           "gt_graph_train_create: data_x, data_y and perm come together (or all null)");
  GT_CHECK(metrics_ring == nullptr || ring_slots >= 1, "gt_graph_train_create: a metrics ring needs ring_slots >= 1");
  GT_CHECK(g_prof_class == KC_NONE && !g_bucket_events_on,
           "gt_graph_train_create: per-kernel profiling and gradient-bucket events must be off while the step is captured");
  static thread_local Layout L;
  GT_TRY(build_layout(*cfg, L));
  cudaStream_t st = (cudaStream_t)stream;
  // eager warm-up: loads every kernel of the step and sets its attributes outside the capture (it only writes grads / metrics / hvo /
  // the workspace, which the first replay overwrites; xbuf / ybuf are read as they are)
  GT_TRY(gt_train_step(cfg, params, pe, xbuf, ybuf, n_seq, hit_loss_penalty, grads, metrics6, hvo, ws, ws_bytes, seed, 0, 0, stream));
  GT_TRY(optimizer_kernels_warm());
  GT_CUDA(cudaStreamSynchronize(st));
  GT_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  int rc = 0;
  if (data_x != nullptr) {
    rc = gather_rows(data_x, perm, 0, xbuf, n_seq, (int64_t)T * cfg->e_src, st, counters + 2);
    if (rc == 0) rc = gather_rows(data_y, perm, 0, ybuf, n_seq, (int64_t)T * cfg->e_tgt, st, counters + 2);
  }
  g_drop_step_ptr = counters;
  if (rc == 0) rc = gt_train_step(cfg, params, pe, xbuf, ybuf, n_seq, hit_loss_penalty, grads, metrics6, hvo, ws, ws_bytes, seed, 0, 0, stream);
  g_drop_step_ptr = nullptr;
  if (rc == 0) rc = optimizer == 0 ? sgd_step(params, grads, L.total, lr, 1.f, st)
                                   : adam_step(params, grads, m, v, L.total, lr, 0.9f, 0.999f, 1e-8f, 1, 1.f, st, counters + 1);
  if (rc == 0) rc = counter_advance(counters, n_seq, metrics6, metrics_ring, ring_slots, st);
  cudaGraph_t graph = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(st, &graph);
  if (rc != 0) { if (graph) cudaGraphDestroy(graph); cudaGetLastError(); return rc; }
  GT_CHECK(ce == cudaSuccess && graph != nullptr, std::string("gt_graph_train_create: capture failed: ") + cudaGetErrorString(ce));
  cudaGraphExec_t exec = nullptr;
  const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
  if (ie != cudaSuccess) { cudaGraphDestroy(graph); GT_FAIL(std::string("gt_graph_train_create: instantiate failed: ") + cudaGetErrorString(ie)); }
  GtGraph *h = new GtGraph{graph, exec};
  *graph_out = h;
  return 0;
}
int gt_graph_launch(void *graph, int n_replays, void *stream) {
  GT_CHECK(graph != nullptr && n_replays >= 0, "gt_graph_launch: null graph or negative replay count");
  for (int i = 0; i < n_replays; ++i) GT_CUDA(cudaGraphLaunch(static_cast<GtGraph *>(graph)->exec, (cudaStream_t)stream));
  return 0;
}
int gt_graph_destroy(void *graph) {
  if (graph == nullptr) return 0;
  GtGraph *h = static_cast<GtGraph *>(graph);
  cudaGraphExecDestroy(h->exec);
  cudaGraphDestroy(h->graph);
  delete h;
  return 0;
}

int gt_predict(const gt_config *cfg, const float *params, const float *pe, const float *src, int64_t n_seq, float thres,
               float *hvo_out, void *ws, int64_t ws_bytes, void *stream) {
  return gt_predict_variant(cfg, params, pe, src, n_seq, thres, hvo_out, ws, ws_bytes, 0, stream);
}

int gt_predict_variant(const gt_config *cfg, const float *params, const float *pe, const float *src, int64_t n_seq, float thres,
                       float *hvo_out, void *ws, int64_t ws_bytes, int variant, void *stream) {
  GT_NVTX("gt_predict_variant");
  static thread_local Ctx x;
  GT_TRY(make_ctx(x, cfg, params, nullptr, pe, n_seq, false, 0, 0, 0, stream));
  GT_CHECK(src && hvo_out, "gt_predict: null pointer");
  GT_CHECK(thres >= 0.f && thres <= 1.f, "gt_predict: threshold must be in [0,1]");
  if (fused_layers(*cfg))
    return tc_predict(*cfg, x.L, params, pe, src, n_seq, thres, hvo_out, ws, ws_bytes, x.st);
  static thread_local Plan pl;
  GT_TRY(check_ws(cfg, n_seq, variant == 1 ? 0 : 2, ws, ws_bytes, pl));
  GT_TRY(encoder_fwd(x, pl, src));
  if (cfg->n_dec == 0) return head_fwd(x, pl.mem, hvo_out, thres);
  if (variant != 1) return predict_decode(x, pl, thres, hvo_out);
  // variant 1 — the reference's literal loop, BGT/models/transformer.py:48-83: 32 full decoder passes; step i feeds
  // (thresholded h, raw v, raw o) to i+1.  Kept as the cross-check of the KV-cached decode.
  GT_CUDA(cudaMemsetAsync(pl.tgt_in, 0, (size_t)x.M * cfg->e_tgt * sizeof(float), x.st));
  float *full = pl.full;
  for (int i = 0; i < T; ++i) {
    GT_TRY(decoder_fwd(x, pl, pl.tgt_in));
    GT_TRY(head_fwd(x, pl.zdec, full, -1.f));
    GT_TRY(predict_feedback(full, pl.tgt_in, hvo_out, n_seq, cfg->e_tgt, i, thres, x.st));
  }
  return 0;
}

int gt_sgd_step(float *p, const float *g, int64_t n, float lr, float grad_scale, void *stream) {
  GT_CHECK(p && g && n >= 0, "gt_sgd_step: bad arguments");
  return sgd_step(p, g, n, lr, grad_scale, (cudaStream_t)stream);
}
int gt_adam_step(float *p, const float *g, float *m, float *v, int64_t n, float lr, float beta1, float beta2, float eps,
                 int64_t step, float grad_scale, void *stream) {
  GT_CHECK(p && g && m && v && n >= 0, "gt_adam_step: bad arguments");
  return adam_step(p, g, m, v, n, lr, beta1, beta2, eps, step, grad_scale, (cudaStream_t)stream);
}

int gt_debug_dropout_mask(uint64_t seed, uint64_t step, int32_t site, float p, int64_t idx0, int64_t n, uint8_t *keep,
                          void *stream) {
  GT_CHECK(keep && n >= 0, "gt_debug_dropout_mask: bad arguments");
  return debug_dropout_mask(site_key(seed, step, site), drop_threshold(p), idx0, n, keep, (cudaStream_t)stream);
}

int gt_grad_buckets(const gt_config *cfg, int64_t *offsets, int64_t *sizes, int max_entries) {
  if (validate_config(cfg)) return -1;
  static thread_local Layout L;
  build_layout(*cfg, L);
  if (!offsets || !sizes || max_entries < grad_bucket_count(*cfg)) { set_error("gt_grad_buckets: max_entries too small"); return -1; }
  return bucket_ranges(*cfg, L, offsets, sizes);
}

int gt_grad_events_enable(int max_buckets) {
  for (auto &e : g_bucket_events) cudaEventDestroy(e);
  g_bucket_events.clear();
  g_bucket_events_on = false;
  if (max_buckets <= 0) return 0;
  g_bucket_events.resize((size_t)max_buckets);
  for (auto &e : g_bucket_events) GT_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  g_bucket_events_on = true;
  return 0;
}

int gt_grad_bucket_wait(int bucket, void *stream) {
  GT_CHECK(g_bucket_events_on && bucket >= 0 && bucket < (int)g_bucket_events.size(), "gt_grad_bucket_wait: bucket events are not enabled / bad index");
  GT_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, g_bucket_events[bucket], 0));
  return 0;
}

int64_t gt_launch_count(int kernel_class) {
  if (kernel_class < 0) { int64_t t = 0; for (int i = 0; i < KC_MAX; ++i) t += g_launches[i]; return t; }
  return kernel_class < KC_MAX ? g_launches[kernel_class] : 0;
}

int gt_profile_enable(int kernel_class, int max_records) {
  for (auto &e : g_prof_events) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
  g_prof_events.clear();
  g_prof_used = 0;
  g_prof_class = KC_NONE;
  if (kernel_class == 0 || kernel_class < -1 || kernel_class >= KC_MAX || max_records <= 0) return 0;
  g_prof_events.resize((size_t)max_records);
  g_prof_cls.assign((size_t)max_records, 0);
  for (auto &e : g_prof_events) { GT_CUDA(cudaEventCreate(&e.first)); GT_CUDA(cudaEventCreate(&e.second)); }
  g_prof_class = kernel_class;
  return 0;
}

int gt_profile_collect(double *total_ms, int64_t *launches) {
  GT_CHECK(total_ms && launches, "gt_profile_collect: null pointer");
  double tot = 0;
  for (size_t i = 0; i < g_prof_used; ++i) {
    GT_CUDA(cudaEventSynchronize(g_prof_events[i].second));
    float ms = 0.f;
    GT_CUDA(cudaEventElapsedTime(&ms, g_prof_events[i].first, g_prof_events[i].second));
    tot += ms;
  }
  *total_ms = tot;
  *launches = (int64_t)g_prof_used;
  g_prof_used = 0;
  return 0;
}

int gt_profile_collect_class(int kernel_class, double *total_ms, int64_t *launches) {
  GT_CHECK(total_ms && launches, "gt_profile_collect_class: null pointer");
  double tot = 0;
  int64_t n = 0;
  for (size_t i = 0; i < g_prof_used; ++i) {
    if (g_prof_cls[i] != kernel_class) continue;
    GT_CUDA(cudaEventSynchronize(g_prof_events[i].second));
    float ms = 0.f;
    GT_CUDA(cudaEventElapsedTime(&ms, g_prof_events[i].first, g_prof_events[i].second));
    tot += ms;
    ++n;
  }
  *total_ms = tot;
  *launches = n;
  return 0;
}

int gt_debug_gemm(int tc, const float *a, int64_t sam, int64_t sak, const float *b, int64_t sbn, int64_t sbk, float *c, int64_t ldc,
                  int64_t m, int64_t n, int64_t k, int flags, const float *bias, const float *residual, int64_t ld_res,
                  const float *mask_pos, int64_t ld_mask, float mask_scale, float drop_p, uint64_t seed, uint64_t step, int32_t site,
                  int64_t row0, int64_t split_k_chunk, void *stream) {
  GT_CHECK(a && b && c && m >= 0 && n >= 0 && k >= 1, "gt_debug_gemm: bad arguments");
  GemmEpi e;
  e.bias = bias; e.relu = flags & 1; e.accumulate = (flags >> 1) & 1; e.atomic = (flags >> 2) & 1;
  e.residual = residual; e.ld_res = ld_res; e.mask_pos = mask_pos; e.ld_mask = ld_mask; e.mask_scale = mask_scale;
  const uint32_t thr = drop_threshold(drop_p);
  if (thr) { e.drop.thr = thr; e.drop.key = site_key(seed, step, site); e.drop.scale = drop_scale(thr); e.drop_row0 = row0; }
  if (tc) return gemm_tc(a, sam, sak, b, sbn, sbk, c, ldc, m, n, k, e, split_k_chunk, (cudaStream_t)stream, tc == 2 ? 1 : 0);
  return gemm_f32(a, sam, sak, b, sbn, sbk, c, ldc, m, n, k, e, split_k_chunk, (cudaStream_t)stream);
}

int gt_debug_gemm_scratch(void *scratch, int64_t bytes) {
  gemm_tc_bind_scratch(scratch, bytes);
  return 0;
}

int gt_debug_umma_rate(int n, int n_mma, int ksteps, float *out, void *stream) {
  return t256_debug_umma_rate(n, n_mma, ksteps, out, (cudaStream_t)stream);
}
int gt_debug_tc_gemm(const uint16_t *a, const uint16_t *b, float *d, int m, int n, int k, int variant, void *stream) {
  return tc_debug_gemm(a, b, d, m, n, k, variant, (cudaStream_t)stream);
}

}  // extern "C"
