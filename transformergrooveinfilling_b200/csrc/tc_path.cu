// tc_path.cu — bf16 tensor-core path (tcgen05 / TMEM / bulk TMA).  See DESIGN.md §"bf16 path".
#include <string.h>

#include "tc_layers.cuh"
#include "tc256.cuh"
#include "tc_path.cuh"
#include "umma.cuh"

namespace gt {
using namespace umma;

// =============================================================================================
// Stand-alone tile GEMM: D[M,N] = A[M,K] B[N,K]^T.  One CTA per 128 rows; the whole B and the CTA's
// A tile are staged into the canonical K-major layout by the threads themselves, one elected thread
// issues K/16 UMMAs, completion arrives on an mbarrier, every warp drains its 32 TMEM lanes.
// Used by tests/test_tc_engine.py to validate descriptors, TMEM addressing and the epilogue mapping.
// variant bit0: swap LBO/SBO ; bit1: use M=64 tiles (two halves)
// =============================================================================================
__global__ void __launch_bounds__(128) tc_debug_gemm_kernel(const uint16_t *__restrict__ A, const uint16_t *__restrict__ B,
                                                           float *__restrict__ D, int M, int N, int K, int variant) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  uint8_t *sA = smem;                            // 128 x K bf16
  uint8_t *sB = smem + (size_t)128 * K * 2;      // N   x K bf16
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  const int m0 = blockIdx.x * 128;
  uint32_t ncols = 32;
  while ((int)ncols < N) ncols <<= 1;

  if (warp == 0) tmem_alloc(&tmem_slot, ncols);
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }

  // stage operands: 16-byte chunks.  bit2 / bit3: the operand is given TRANSPOSED in global memory
  // ([K, M] / [K, N] row-major) and staged with k as the row index -> consumed as an MN-major operand.
  const bool a_t = variant & 4, b_t = variant & 8;
  if (!a_t) {
    for (int kb = 0; kb < K / 8; ++kb) {
      uint4 v = *reinterpret_cast<const uint4 *>(A + (size_t)(m0 + tid) * K + kb * 8);
      *reinterpret_cast<uint4 *>(sA + kmajor_off(tid, kb * 8, 128)) = v;
    }
  } else {
    for (int k = tid; k < K; k += 128)
      for (int mb = 0; mb < 16; ++mb) {
        uint4 v = *reinterpret_cast<const uint4 *>(A + (size_t)k * M + m0 + mb * 8);
        *reinterpret_cast<uint4 *>(sA + kmajor_off(k, mb * 8, K)) = v;
      }
  }
  if (!b_t) {
    for (int r = tid; r < N; r += 128)
      for (int kb = 0; kb < K / 8; ++kb) {
        uint4 v = *reinterpret_cast<const uint4 *>(B + (size_t)r * K + kb * 8);
        *reinterpret_cast<uint4 *>(sB + kmajor_off(r, kb * 8, N)) = v;
      }
  } else {
    for (int k = tid; k < K; k += 128)
      for (int nb = 0; nb < N / 8; ++nb) {
        uint4 v = *reinterpret_cast<const uint4 *>(B + (size_t)k * N + nb * 8);
        *reinterpret_cast<uint4 *>(sB + kmajor_off(k, nb * 8, K)) = v;
      }
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (tid == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N, a_t ? 1 : 0, b_t ? 1 : 0);
    // K-major image [R rows x K]: K-adjacent core matrices (R/8)*128 B apart (LBO), row-adjacent 128 B (SBO).
    // MN-major use of an image [K rows x R]: k-adjacent cores 128 B apart (LBO), mn-adjacent (K/8)*128 B (SBO).
    uint32_t a_lbo = a_t ? 128u : (128 / 8) * 128u, a_sbo = a_t ? (uint32_t)(K / 8) * 128u : 128u;
    uint32_t b_lbo = b_t ? 128u : (uint32_t)(N / 8) * 128u, b_sbo = b_t ? (uint32_t)(K / 8) * 128u : 128u;
    if (variant & 1) { uint32_t t = a_lbo; a_lbo = a_sbo; a_sbo = t; t = b_lbo; b_lbo = b_sbo; b_sbo = t; }
    for (int k16 = 0; k16 < K / 16; ++k16) {
      uint32_t a_off = a_t ? (uint32_t)(k16 * 2) * 128u : (uint32_t)(k16 * 2) * (128 / 8) * 128u;
      uint32_t b_off = b_t ? (uint32_t)(k16 * 2) * 128u : (uint32_t)(k16 * 2) * (uint32_t)(N / 8) * 128u;
      uint64_t ad = make_desc(smem_u32(sA) + a_off, a_lbo, a_sbo);
      uint64_t bd = make_desc(smem_u32(sB) + b_off, b_lbo, b_sbo);
      mma_bf16_ss(tmem, ad, bd, idesc, k16 > 0 ? 1u : 0u);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  fence_after_sync();

  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
    tmem_ld_wait();
    float *out = D + (size_t)(m0 + warp * 32 + lane) * N + c0;
#pragma unroll
    for (int j = 0; j < 16; ++j) out[j] = v[j];
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, ncols);
}

int tc_debug_gemm(const uint16_t *a, const uint16_t *b, float *d, int m, int n, int k, int variant, cudaStream_t st) {
  GT_CHECK(a && b && d, "tc_debug_gemm: null pointer");
  GT_CHECK(m > 0 && m % 128 == 0, "tc_debug_gemm: M must be a positive multiple of 128");
  GT_CHECK(n >= 16 && n <= 256 && n % 16 == 0, "tc_debug_gemm: N must be a multiple of 16 in [16,256]");
  GT_CHECK(k >= 16 && k % 16 == 0, "tc_debug_gemm: K must be a positive multiple of 16");
  size_t smem = (size_t)(128 + n) * k * 2;
  GT_CHECK(smem <= 200 * 1024, "tc_debug_gemm: operands do not fit in shared memory");
  GT_CUDA(cudaFuncSetAttribute(tc_debug_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  tc_debug_gemm_kernel<<<m / 128, 128, smem, st>>>(a, b, d, m, n, k, variant);
  GT_CUDA(cudaGetLastError());
  return 0;
}

// =============================================================================================
// model passes (encoder-only): the L encoder layers run in the fused tcgen05 kernels of
// tc_layers.cu; the input layer, the final LayerNorm + output head and the loss (a few % of the
// step, K = 16..32 contractions) reuse the fp32 kernels of kernels_simt.cu.
// =============================================================================================
struct TcPlan {
  uint8_t *img;
  float *r0, *x[TC_MAX_LAYERS + 1], *u1[TC_MAX_LAYERS], *u2[TC_MAX_LAYERS];
  float *z, *mf, *rf;
  float *d_hvo, *loss_partials, *dlog, *dxa, *dxb, *g0;
  int64_t bytes;
  uint32_t img_stride;
};

// the fused stem / tail kernels of edge32.cu cover the reference's two source widths (16 MSO features, 27 hvo channels)
static bool tc_fused_edges(const gt_config &c) { return c.d_model == 32 && (c.e_src == 16 || c.e_src == 27) && c.e_tgt == 27; }

static void tc_make_plan(const gt_config &c, int64_t n_seq, int mode, char *base, TcPlan &P) {
  memset(&P, 0, sizeof(P));
  const int64_t M = n_seq * T, d = c.d_model;
  int64_t off = 0;
  auto take = [&](int64_t nbytes) -> char * {
    int64_t o = off;
    off += (nbytes + 255) / 256 * 256;
    return base ? base + o : nullptr;
  };
  P.img_stride = (tc_img(c.d_model, c.dim_ff).total + 255u) & ~255u;
  P.img = reinterpret_cast<uint8_t *>(take((int64_t)P.img_stride * c.n_enc));
  const bool train = mode == 1;
  const bool fused = tc_fused_edges(c);        // edge32.cu: no r0 / z / g0 / d_hvo intermediates
  if (!fused) P.r0 = reinterpret_cast<float *>(take(M * d * 4));
  if (train) {
    for (int l = 0; l <= c.n_enc; ++l) P.x[l] = reinterpret_cast<float *>(take(M * d * 4));
    for (int l = 0; l < c.n_enc; ++l) {
      P.u1[l] = reinterpret_cast<float *>(take(M * d * 4));
      P.u2[l] = reinterpret_cast<float *>(take(M * d * 4));
    }
    P.mf = reinterpret_cast<float *>(take(M * 4));
    P.rf = reinterpret_cast<float *>(take(M * 4));
  } else {
    float *a = reinterpret_cast<float *>(take(M * d * 4)), *b = reinterpret_cast<float *>(take(M * d * 4));
    for (int l = 0; l <= c.n_enc; ++l) P.x[l] = (l & 1) ? b : a;
  }
  if (!fused) P.z = reinterpret_cast<float *>(take(M * d * 4));
  if (train) {
    if (!fused) P.d_hvo = reinterpret_cast<float *>(take(M * c.e_tgt * 4));
    P.loss_partials = reinterpret_cast<float *>(take((fused ? edge32_loss_partials(n_seq) : loss_scratch_floats(n_seq)) * 4));
    P.dlog = reinterpret_cast<float *>(take(M * c.e_tgt * 4));
    P.dxa = reinterpret_cast<float *>(take(M * d * 4));
    P.dxb = reinterpret_cast<float *>(take(M * d * 4));
    if (!fused) P.g0 = reinterpret_cast<float *>(take(M * d * 4));
  }
  P.bytes = off;
}

struct TcCtx {
  gt_config c;
  const Layout *L;
  const float *P;
  float *G;
  const float *pe;
  int64_t n_seq, M;
  bool train;
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
};

static int tc_check(const gt_config &c, int64_t n_seq, int mode, void *ws, int64_t ws_bytes, TcPlan &pl) {
  std::string why;
  GT_CHECK(tc_shape_supported(c, &why), "precision=bf16 is not available for this configuration (" + why + "); use precision=fp32");
  GT_CHECK(ws != nullptr && ((uintptr_t)ws & 255) == 0, "workspace must be non-null and 256-byte aligned");
  tc_make_plan(c, n_seq, mode, (char *)ws, pl);
  GT_CHECK(ws_bytes >= pl.bytes, "workspace too small: need " + std::to_string(pl.bytes) + " bytes");
  return 0;
}

int64_t tc_workspace_bytes(const gt_config &c, int64_t n_seq, int mode) {
  if (c.d_model == 256) return t256_workspace_bytes(c, n_seq, mode);
  std::string why;
  if (!tc_shape_supported(c, &why)) {
    set_error("precision=bf16 is not available for this configuration (" + why + "); use precision=fp32");
    return -1;
  }
  static thread_local TcPlan pl;
  tc_make_plan(c, n_seq, mode, nullptr, pl);
  return pl.bytes;
}

static TcLayerArgs tc_layer_args(const TcCtx &x, const TcPlan &pl, int l) {
  TcLayerArgs a;
  memset(&a, 0, sizeof(a));
  const LayerP &p = x.L->enc[l];
  a.img = pl.img + (size_t)l * pl.img_stride;
  a.img_bytes = tc_img(x.c.d_model, x.c.dim_ff).total;
  a.bqkv = x.P + p.sa.b_in; a.bo = x.P + p.sa.b_out; a.b1 = x.P + p.b1; a.b2 = x.P + p.b2;
  a.g1 = x.P + p.g1; a.be1 = x.P + p.be1; a.g2 = x.P + p.g2; a.be2 = x.P + p.be2;
  if (x.G) {
    a.gwqkv = x.G + p.sa.w_in; a.gbqkv = x.G + p.sa.b_in; a.gwo = x.G + p.sa.w_out; a.gbo = x.G + p.sa.b_out;
    a.gw1 = x.G + p.w1; a.gb1 = x.G + p.b1; a.gw2 = x.G + p.w2; a.gb2 = x.G + p.b2;
    a.gg1 = x.G + p.g1; a.gbe1 = x.G + p.be1; a.gg2 = x.G + p.g2; a.gbe2 = x.G + p.be2;
  }
  a.M = x.M; a.n_tiles = (int)((x.n_seq + 3) / 4);
  a.F = x.c.dim_ff; a.FC = tc_ffn_chunk(x.c.dim_ff); a.H = x.c.nhead; a.dh = x.c.d_model / x.c.nhead;
  a.d_attn = x.drop(site_id(0, l, 0)); a.d1 = x.drop(site_id(0, l, 1)); a.d_ffn = x.drop(site_id(0, l, 2));
  a.d2 = x.drop(site_id(0, l, 3));
  a.seq0 = x.seq0;
  return a;
}

static int tc_prep(const TcCtx &x, const TcPlan &pl) {
  TcPrepArgs a;
  memset(&a, 0, sizeof(a));
  a.params = x.P; a.img = pl.img; a.img_stride = pl.img_stride; a.n_layers = x.c.n_enc;
  a.D = x.c.d_model; a.F = x.c.dim_ff; a.FC = tc_ffn_chunk(x.c.dim_ff);
  for (int l = 0; l < x.c.n_enc; ++l) {
    a.w_in[l] = x.L->enc[l].sa.w_in; a.w_out[l] = x.L->enc[l].sa.w_out; a.w1[l] = x.L->enc[l].w1; a.w2[l] = x.L->enc[l].w2;
    a.b1[l] = x.L->enc[l].b1;
  }
  return tc_prep_weights(a, x.st);
}

// y != nullptr: the tail also evaluates calculate_loss (metrics6) and leaves dL/dlogits in pl.dlog (fused edges only)
static int tc_forward_all(const TcCtx &x, const TcPlan &pl, const float *src, float *hvo, bool save, float thres,
                          const float *y = nullptr, float penalty = 0.f, float *metrics6 = nullptr) {
  GT_NVTX("groove.forward");
  const int d = x.c.d_model;
  const bool fused = tc_fused_edges(x.c);
  GT_TRY(tc_prep(x, pl));
  if (fused) {
    GT_TRY(edge32_stem_fwd(src, x.c.e_src, x.P + x.L->in_enc_w, x.P + x.L->in_enc_b, x.pe, pl.x[0], x.M, x.drop(SITE_IN_ENC),
                           x.seq0 * T, x.st));
  } else {
    GemmEpi e; e.bias = x.P + x.L->in_enc_b; e.relu = 1;
    GT_TRY(gemm_f32(src, x.c.e_src, 1, x.P + x.L->in_enc_w, x.c.e_src, 1, pl.r0, d, x.M, d, x.c.e_src, e, 0, x.st));
    GT_TRY(pe_dropout_fwd(pl.r0, x.pe, pl.x[0], x.M, d, x.drop(SITE_IN_ENC), x.seq0 * T, x.st));
  }
  for (int l = 0; l < x.c.n_enc; ++l) {
    TcLayerArgs a = tc_layer_args(x, pl, l);
    a.x_in = pl.x[l]; a.x_out = pl.x[l + 1];
    a.u1 = save ? pl.u1[l] : nullptr; a.u2 = save ? pl.u2[l] : nullptr;
    GT_TRY(tc_layer_fwd(d, a, x.st));
  }
  if (fused) {
    const float *xl = pl.x[x.c.n_enc], *g = x.P + x.L->enc_norm_g, *b = x.P + x.L->enc_norm_b, *w = x.P + x.L->out_w, *bo = x.P + x.L->out_b;
    if (y != nullptr)
      return edge32_tail_fwd_loss(xl, g, b, w, bo, hvo, pl.mf, pl.rf, x.M, y, penalty, pl.dlog, pl.loss_partials, metrics6, x.st);
    return edge32_tail_fwd(xl, g, b, w, bo, hvo, pl.mf, pl.rf, x.M, thres, x.st);
  }
  GT_CHECK(y == nullptr, "tc_forward_all: fused loss needs the fused edge kernels");
  Drop none;
  GT_TRY(ln_fwd(pl.x[x.c.n_enc], nullptr, x.P + x.L->enc_norm_g, x.P + x.L->enc_norm_b, nullptr, pl.z, pl.mf, pl.rf, x.M, d,
                none, 0, x.st));
  GemmEpi eh; eh.bias = x.P + x.L->out_b;
  GT_TRY(gemm_f32(pl.z, d, 1, x.P + x.L->out_w, d, 1, hvo, x.c.e_tgt, x.M, x.c.e_tgt, d, eh, 0, x.st));
  return head_activation(hvo, x.M, x.c.e_tgt, thres, x.st);
}

static int tc_wgrad(const TcCtx &x, const float *dY, int64_t N, const float *X, int64_t K, float *dW, float *db) {
  GemmEpi e; e.atomic = 1;
  GT_TRY(gemm_f32(dY, 1, N, X, 1, K, dW, K, N, K, x.M, e, 2048, x.st));
  return colsum_f32(dY, N, x.M, (int)N, db, x.st);
}

// hvo == nullptr: d_hvo already holds dL/dlogits (left by the fused tail + loss forward)
static int tc_backward_all(const TcCtx &x, const TcPlan &pl, const float *src, const float *hvo, const float *d_hvo) {
  GT_NVTX("groove.backward");
  const int d = x.c.d_model, E = x.c.e_tgt;
  const bool fused = tc_fused_edges(x.c);
  Drop none;
  if (fused) {
    GT_TRY(edge32_tail_bwd(d_hvo, hvo, pl.x[x.c.n_enc], pl.mf, pl.rf, x.P + x.L->enc_norm_g, x.P + x.L->enc_norm_b, x.P + x.L->out_w,
                           pl.dxb, x.G + x.L->out_w, x.G + x.L->out_b, x.G + x.L->enc_norm_g, x.G + x.L->enc_norm_b, x.M, x.st));
  } else {
    GT_TRY(head_activation_bwd(d_hvo, hvo, pl.dlog, x.M, E, x.st));
    GT_TRY(tc_wgrad(x, pl.dlog, E, pl.z, d, x.G + x.L->out_w, x.G + x.L->out_b));
    GemmEpi e0;
    GT_TRY(gemm_f32(pl.dlog, E, 1, x.P + x.L->out_w, 1, d, pl.dxa, d, x.M, d, E, e0, 0, x.st));
    GT_TRY(ln_bwd(pl.dxa, pl.x[x.c.n_enc], pl.mf, pl.rf, x.P + x.L->enc_norm_g, pl.dxb, nullptr, x.G + x.L->enc_norm_g,
                  x.G + x.L->enc_norm_b, x.M, d, none, 0, x.st));
  }
  grad_bucket_ready(x.c, BK_HEAD, 0, x.st);
  float *cur = pl.dxb, *oth = pl.dxa;
  for (int l = x.c.n_enc - 1; l >= 0; --l) {
    TcLayerArgs a = tc_layer_args(x, pl, l);
    a.x_in = pl.x[l]; a.u1_in = pl.u1[l]; a.u2_in = pl.u2[l]; a.dy = cur; a.dx = oth;
    GT_TRY(tc_layer_bwd(d, a, x.st));
    grad_bucket_ready(x.c, BK_ENC_LAYER, l, x.st);
    float *t = cur; cur = oth; oth = t;
  }
  if (fused) {
    GT_TRY(edge32_stem_bwd(cur, src, x.c.e_src, x.P + x.L->in_enc_w, x.P + x.L->in_enc_b, x.G + x.L->in_enc_w, x.G + x.L->in_enc_b,
                           x.M, x.drop(SITE_IN_ENC), x.seq0 * T, x.st));
  } else {
    GT_TRY(pe_dropout_bwd(cur, pl.r0, pl.g0, x.M, d, x.drop(SITE_IN_ENC), x.seq0 * T, x.st));
    GT_TRY(tc_wgrad(x, pl.g0, d, src, x.c.e_src, x.G + x.L->in_enc_w, x.G + x.L->in_enc_b));
  }
  grad_bucket_ready(x.c, BK_IN_ENC, 0, x.st);
  return 0;
}

static void tc_ctx(TcCtx &x, const gt_config &c, const Layout &L, const float *params, float *grads, const float *pe,
                   int64_t n_seq, bool train, uint64_t seed, uint64_t step, int64_t seq0, cudaStream_t st) {
  x.c = c; x.L = &L; x.P = params; x.G = grads; x.pe = pe; x.n_seq = n_seq; x.M = n_seq * T; x.train = train;
  x.seed = seed; x.step = step; x.seq0 = seq0; x.st = st;
}

// ---- encoder stack of an encoder-DECODER model (runner.cu hybrid path) ------------------------------------------------
// The encoder of GrooveTransformer is the encoder-only stack (BGT/models/transformer.py:25-29 vs :101-104), so in bf16
// mode with d_model = 32 its layers run in the fused kernels while the decoder runs per-op on gemm_tc.  The caller owns the
// activation buffers (row-major fp32 [tokens, 32]) and the image block (n_enc * tc_enc_img_stride bytes).
static gt_config enc_only(const gt_config &c) { gt_config e = c; e.n_dec = 0; return e; }
bool tc_encoder_supported(const gt_config &c) {
  return c.precision == GT_PREC_BF16 && c.d_model == 32 && tc_shape_supported(enc_only(c), nullptr) && tc_fused_edges(c);
}
uint32_t tc_enc_img_stride(const gt_config &c) { return (tc_img(c.d_model, c.dim_ff).total + 255u) & ~255u; }

static void tc_ext_ctx(TcCtx &x, TcPlan &pl, const gt_config &c, const Layout &L, const float *params, float *grads, uint8_t *img,
                       int64_t n_seq, bool train, uint64_t seed, uint64_t step, int64_t seq0, cudaStream_t st) {
  tc_ctx(x, c, L, params, grads, nullptr, n_seq, train, seed, step, seq0, st);
  memset(&pl, 0, sizeof(pl));
  pl.img = img; pl.img_stride = tc_enc_img_stride(c);
}
int tc_enc_prep(const gt_config &c, const Layout &L, const float *params, uint8_t *img, cudaStream_t st) {
  static thread_local TcPlan pl;
  TcCtx x;
  tc_ext_ctx(x, pl, c, L, params, nullptr, img, 1, false, 0, 0, 0, st);
  return tc_prep(x, pl);
}
int tc_enc_layer_fwd(const gt_config &c, const Layout &L, const float *params, uint8_t *img, int l, const float *x_in, float *x_out,
                     float *u1, float *u2, int64_t n_seq, bool train, uint64_t seed, uint64_t step, int64_t seq0, cudaStream_t st) {
  static thread_local TcPlan pl;
  TcCtx x;
  tc_ext_ctx(x, pl, c, L, params, nullptr, img, n_seq, train, seed, step, seq0, st);
  TcLayerArgs a = tc_layer_args(x, pl, l);
  a.x_in = x_in; a.x_out = x_out; a.u1 = u1; a.u2 = u2;
  return tc_layer_fwd(c.d_model, a, st);
}
int tc_enc_layer_bwd(const gt_config &c, const Layout &L, const float *params, float *grads, uint8_t *img, int l, const float *x_in,
                     const float *u1, const float *u2, const float *dy, float *dx, int64_t n_seq, uint64_t seed, uint64_t step,
                     int64_t seq0, cudaStream_t st) {
  static thread_local TcPlan pl;
  TcCtx x;
  tc_ext_ctx(x, pl, c, L, params, grads, img, n_seq, true, seed, step, seq0, st);
  TcLayerArgs a = tc_layer_args(x, pl, l);
  a.x_in = x_in; a.u1_in = u1; a.u2_in = u2; a.dy = dy; a.dx = dx;
  return tc_layer_bwd(c.d_model, a, st);
}

// ---- feed-forward block of a DECODER layer (third block: x3 = LN3(x2 + drop(FFN(x2)))) on the fused kernels, TC_MODE_FFN ----
// dec_img: 2 * n_dec * tc_enc_img_stride bytes: block l = (self-attention in/out projections, W1, W2) of decoder layer l,
// block n_dec + l = (cross-attention in/out projections, ...) of decoder layer l
int tc_dec_prep(const gt_config &c, const Layout &L, const float *params, uint8_t *dec_img, cudaStream_t st) {
  for (int pass = 0; pass < 2; ++pass) {
    TcPrepArgs a;
    memset(&a, 0, sizeof(a));
    a.params = params; a.img = dec_img + (size_t)pass * c.n_dec * tc_enc_img_stride(c); a.img_stride = tc_enc_img_stride(c);
    a.n_layers = c.n_dec; a.D = c.d_model; a.F = c.dim_ff; a.FC = tc_ffn_chunk(c.dim_ff);
    for (int l = 0; l < c.n_dec; ++l) {
      const AttnP &ap = pass == 0 ? L.dec[l].sa : L.dec[l].ca;
      a.w_in[l] = ap.w_in; a.w_out[l] = ap.w_out; a.w1[l] = L.dec[l].w1; a.w2[l] = L.dec[l].w2;
      a.b1[l] = L.dec[l].b1;
    }
    GT_TRY(tc_prep_weights(a, st));
  }
  return 0;
}
bool tc_dec_attn_supported(const gt_config &c) { const int dh = c.d_model / c.nhead; return dh == 2 || dh == 4 || dh == 8; }
// attention block of a decoder layer: cross = 0: causal self-attention + LayerNorm1 ; cross = 1: cross-attention over `mem` + LayerNorm2
static TcLayerArgs tc_dec_attn_args(const TcCtx &x, uint8_t *dec_img, int l, int cross) {
  TcLayerArgs a;
  memset(&a, 0, sizeof(a));
  const LayerP &p = x.L->dec[l];
  const AttnP &ap = cross ? p.ca : p.sa;
  a.mode = cross ? TC_MODE_ATTN_CROSS : TC_MODE_ATTN_CAUSAL;
  a.img = dec_img + (size_t)(cross ? x.c.n_dec + l : l) * tc_enc_img_stride(x.c);
  a.img_bytes = tc_img(x.c.d_model, x.c.dim_ff).w1;            // in-projection + out-projection images only
  a.bqkv = x.P + ap.b_in; a.bo = x.P + ap.b_out;
  a.g1 = x.P + (cross ? p.g2 : p.g1); a.be1 = x.P + (cross ? p.be2 : p.be1);
  if (x.G) {
    a.gwqkv = x.G + ap.w_in; a.gbqkv = x.G + ap.b_in; a.gwo = x.G + ap.w_out; a.gbo = x.G + ap.b_out;
    a.gg1 = x.G + (cross ? p.g2 : p.g1); a.gbe1 = x.G + (cross ? p.be2 : p.be1);
  }
  a.M = x.M; a.n_tiles = (int)((x.n_seq + 3) / 4);
  a.F = x.c.dim_ff; a.FC = tc_ffn_chunk(x.c.dim_ff); a.H = x.c.nhead; a.dh = x.c.d_model / x.c.nhead;
  a.d_attn = x.drop(site_id(1, l, cross ? 4 : 0)); a.d1 = x.drop(site_id(1, l, cross ? 5 : 1));
  a.seq0 = x.seq0;
  return a;
}
int tc_dec_attn_fwd(const gt_config &c, const Layout &L, const float *params, uint8_t *dec_img, int l, int cross, const float *x_in,
                    const float *mem, float *x_out, float *u, int64_t n_seq, bool train, uint64_t seed, uint64_t step, int64_t seq0,
                    cudaStream_t st) {
  TcCtx x;
  tc_ctx(x, c, L, params, nullptr, nullptr, n_seq, train, seed, step, seq0, st);
  TcLayerArgs a = tc_dec_attn_args(x, dec_img, l, cross);
  a.x_in = x_in; a.mem = mem; a.x_out = x_out; a.u1 = u;
  return tc_layer_fwd(c.d_model, a, st);
}
int tc_dec_attn_bwd(const gt_config &c, const Layout &L, const float *params, float *grads, uint8_t *dec_img, int l, int cross,
                    const float *x_in, const float *mem, const float *u, const float *dy, float *dx, float *dmem, int64_t n_seq,
                    uint64_t seed, uint64_t step, int64_t seq0, cudaStream_t st) {
  TcCtx x;
  tc_ctx(x, c, L, params, grads, nullptr, n_seq, true, seed, step, seq0, st);
  TcLayerArgs a = tc_dec_attn_args(x, dec_img, l, cross);
  a.x_in = x_in; a.mem = mem; a.u1_in = u; a.dy = dy; a.dx = dx; a.dmem = dmem;
  return tc_layer_bwd(c.d_model, a, st);
}
static TcLayerArgs tc_dec_ffn_args(const TcCtx &x, uint8_t *dec_img, int l) {
  TcLayerArgs a;
  memset(&a, 0, sizeof(a));
  const LayerP &p = x.L->dec[l];
  a.mode = TC_MODE_FFN;
  a.img = dec_img + (size_t)l * tc_enc_img_stride(x.c);
  a.img_bytes = tc_img(x.c.d_model, x.c.dim_ff).total;
  a.b1 = x.P + p.b1; a.b2 = x.P + p.b2; a.g2 = x.P + p.g3; a.be2 = x.P + p.be3;
  if (x.G) { a.gw1 = x.G + p.w1; a.gb1 = x.G + p.b1; a.gw2 = x.G + p.w2; a.gb2 = x.G + p.b2; a.gg2 = x.G + p.g3; a.gbe2 = x.G + p.be3; }
  a.M = x.M; a.n_tiles = (int)((x.n_seq + 3) / 4);
  a.F = x.c.dim_ff; a.FC = tc_ffn_chunk(x.c.dim_ff); a.H = x.c.nhead; a.dh = x.c.d_model / x.c.nhead;
  a.d_ffn = x.drop(site_id(1, l, 2)); a.d2 = x.drop(site_id(1, l, 3));
  a.seq0 = x.seq0;
  return a;
}
int tc_dec_ffn_fwd(const gt_config &c, const Layout &L, const float *params, uint8_t *dec_img, int l, const float *x_in, float *x_out,
                   float *u, int64_t n_seq, bool train, uint64_t seed, uint64_t step, int64_t seq0, cudaStream_t st) {
  TcCtx x;
  tc_ctx(x, c, L, params, nullptr, nullptr, n_seq, train, seed, step, seq0, st);
  TcLayerArgs a = tc_dec_ffn_args(x, dec_img, l);
  a.x_in = x_in; a.x_out = x_out; a.u2 = u;
  return tc_layer_fwd(c.d_model, a, st);
}
// the same block over an arbitrary set of token rows (KV-cached decode: one token per sequence), eval mode
int tc_dec_ffn_rows(const gt_config &c, const Layout &L, const float *params, uint8_t *dec_img, int l, const float *x_in, float *x_out,
                    int64_t n_rows, cudaStream_t st) {
  TcCtx x;
  tc_ctx(x, c, L, params, nullptr, nullptr, 1, false, 0, 0, 0, st);
  TcLayerArgs a = tc_dec_ffn_args(x, dec_img, l);
  a.M = n_rows; a.n_tiles = (int)((n_rows + TC_TILE - 1) / TC_TILE);
  a.x_in = x_in; a.x_out = x_out; a.u2 = nullptr;
  return tc_layer_fwd(c.d_model, a, st);
}
int tc_dec_ffn_bwd(const gt_config &c, const Layout &L, const float *params, float *grads, uint8_t *dec_img, int l, const float *x_in,
                   const float *u, const float *dy, float *dx, int64_t n_seq, uint64_t seed, uint64_t step, int64_t seq0,
                   cudaStream_t st) {
  TcCtx x;
  tc_ctx(x, c, L, params, grads, nullptr, n_seq, true, seed, step, seq0, st);
  TcLayerArgs a = tc_dec_ffn_args(x, dec_img, l);
  a.x_in = x_in; a.u2_in = u; a.dy = dy; a.dx = dx;
  return tc_layer_bwd(c.d_model, a, st);
}

int tc_forward(const gt_config &c, const Layout &L, const float *params, const float *pe, const float *src, const float *,
               int64_t n_seq, float *hvo, void *ws, int64_t ws_bytes, bool train, uint64_t seed, uint64_t step, int64_t seq0,
               cudaStream_t st) {
  if (c.d_model == 256) return t256_forward(c, L, params, pe, src, n_seq, hvo, ws, ws_bytes, train, seed, step, seq0, st);
  static thread_local TcPlan pl;
  // gt_forward(train=1) always saves activations (it is the autograd forward); eval forwards do not
  GT_TRY(tc_check(c, n_seq, train ? 1 : 0, ws, ws_bytes, pl));
  TcCtx x;
  tc_ctx(x, c, L, params, nullptr, pe, n_seq, train, seed, step, seq0, st);
  return tc_forward_all(x, pl, src, hvo, train, -1.f);
}

int tc_backward(const gt_config &c, const Layout &L, const float *params, const float *pe, const float *src, const float *,
                int64_t n_seq, const float *hvo, const float *d_hvo, float *grads, void *ws, int64_t ws_bytes, uint64_t seed,
                uint64_t step, int64_t seq0, cudaStream_t st) {
  if (c.d_model == 256) return t256_backward(c, L, params, pe, src, n_seq, hvo, d_hvo, grads, ws, ws_bytes, seed, step, seq0, st);
  static thread_local TcPlan pl;
  GT_TRY(tc_check(c, n_seq, 1, ws, ws_bytes, pl));
  TcCtx x;
  tc_ctx(x, c, L, params, grads, pe, n_seq, true, seed, step, seq0, st);
  return tc_backward_all(x, pl, src, hvo, d_hvo);
}

int tc_train_step(const gt_config &c, const Layout &L, const float *params, const float *pe, const float *src, const float *y,
                  int64_t n_seq, float penalty, float *grads, float *metrics6, float *hvo, void *ws, int64_t ws_bytes,
                  uint64_t seed, uint64_t step, int64_t seq0, cudaStream_t st) {
  if (c.d_model == 256)
    return t256_train_step(c, L, params, pe, src, y, n_seq, penalty, grads, metrics6, hvo, ws, ws_bytes, seed, step, seq0, st);
  static thread_local TcPlan pl;
  GT_TRY(tc_check(c, n_seq, 1, ws, ws_bytes, pl));
  TcCtx x;
  tc_ctx(x, c, L, params, grads, pe, n_seq, true, seed, step, seq0, st);
  GT_CUDA(cudaMemsetAsync(grads, 0, (size_t)L.total * sizeof(float), st));
  if (tc_fused_edges(c)) {
    GT_TRY(tc_forward_all(x, pl, src, hvo, true, -1.f, y, penalty, metrics6));
    return tc_backward_all(x, pl, src, nullptr, pl.dlog);
  }
  GT_TRY(tc_forward_all(x, pl, src, hvo, true, -1.f));
  GT_TRY(loss_fwd_bwd(hvo, y, n_seq, penalty, metrics6, pl.d_hvo, 1.f, pl.loss_partials, st, c.e_tgt / 3));
  return tc_backward_all(x, pl, src, hvo, pl.d_hvo);
}

int tc_predict(const gt_config &c, const Layout &L, const float *params, const float *pe, const float *src, int64_t n_seq,
               float thres, float *hvo_out, void *ws, int64_t ws_bytes, cudaStream_t st) {
  if (c.d_model == 256) return t256_predict(c, L, params, pe, src, n_seq, thres, hvo_out, ws, ws_bytes, st);
  static thread_local TcPlan pl;
  GT_TRY(tc_check(c, n_seq, 0, ws, ws_bytes, pl));
  TcCtx x;
  tc_ctx(x, c, L, params, nullptr, pe, n_seq, false, 0, 0, 0, st);
  return tc_forward_all(x, pl, src, hvo_out, false, thres);
}

}  // namespace gt
