// tc256_path.cu — model passes of the bf16 tensor-core path for d_model = 256 (encoder-only models).
// The L encoder layers run in the fused weight-streaming kernels of tc256.cu on tile-native
// activation layouts; the input layer, the final LayerNorm + output head and the loss (K = 16..32
// contractions, ~1 % of the FLOPs) reuse the row-major fp32 kernels of kernels_simt.cu, with a layout
// conversion kernel at each boundary of the stack.
#include <string.h>

#include "tc256.cuh"

namespace gt {

struct T256Plan {
  uint8_t *img;                                   // stage streams, one block per layer
  uint32_t img_stride;
  float *r0, *x0rm, *xLrm, *z, *mf, *rf;          // row-major boundary buffers
  uint8_t *ximg[TC_MAX_LAYERS + 1];               // bf16 images of the residual stream
  uint8_t *u1img[TC_MAX_LAYERS], *u2img[TC_MAX_LAYERS];
  uint8_t *x1img[TC_MAX_LAYERS], *ctximg[TC_MAX_LAYERS], *himg[TC_MAX_LAYERS];
  uint8_t *qkvimg[TC_MAX_LAYERS];                 // head_dim 128: saved q | k | v group images
  float2 *ln1stat[TC_MAX_LAYERS], *ln2stat[TC_MAX_LAYERS];   // (mean, rstd) per token row of the two LayerNorms, saved by the forward
  float *d_hvo, *loss_partials, *dlog, *dxrm, *dxrm2, *g0, *dxa, *dxb;
  uint8_t *da2img, *da1img, *dhimg, *dqkvimg, *dctx_scratch, *wg_jobs;
  uint8_t *dlimg, *srcimg;                        // fused edges: [128 x 32] bf16 images per tile
  float *edge_scratch;                            // fused edges: T [256][32] + sums [64] (tail), Tin [256][32] (stem)
  int64_t bytes;
  int n_tiles;
};

// the fused stem / tail kernels of edge256.cu cover the reference's two source widths (16 MSO features, 27 hvo channels)
static bool t256_fused_edges(const gt_config &c) { return (c.e_src == 16 || c.e_src == 27) && c.e_tgt == 27; }

static void t256_make_plan(const gt_config &c, int64_t n_seq, int mode, char *base, T256Plan &P) {
  memset(&P, 0, sizeof(P));
  const int64_t M = n_seq * T, d = c.d_model;
  const int n_tiles = (int)((n_seq + 3) / 4);
  const int64_t tf = (int64_t)n_tiles * T256_TILE_F32 * 4, ti = (int64_t)n_tiles * T256_TILE_IMG;
  const int64_t th = (int64_t)n_tiles * 128 * c.dim_ff * 2;
  int64_t off = 0;
  auto take = [&](int64_t nbytes) -> char * {
    int64_t o = off;
    off += (nbytes + 255) / 256 * 256;
    return base ? base + o : nullptr;
  };
  P.n_tiles = n_tiles;
  const int dh = c.d_model / c.nhead;
  P.img_stride = ((t256_img_bytes(c.dim_ff, dh) + 255u) & ~255u) * T256_REP;
  P.img = reinterpret_cast<uint8_t *>(take((int64_t)P.img_stride * c.n_enc));
  const bool train = mode == 1;
  const bool fused = t256_fused_edges(c);
  if (!fused) {
    P.r0 = reinterpret_cast<float *>(take(M * d * 4));
    P.x0rm = reinterpret_cast<float *>(take(M * d * 4));
    P.xLrm = reinterpret_cast<float *>(take(M * d * 4));
    P.z = reinterpret_cast<float *>(take(M * d * 4));
  }
  P.mf = reinterpret_cast<float *>(take(M * 4));
  P.rf = reinterpret_cast<float *>(take(M * 4));
  if (train) {
    for (int l = 0; l <= c.n_enc; ++l) P.ximg[l] = reinterpret_cast<uint8_t *>(take(ti));
    for (int l = 0; l < c.n_enc; ++l) {
      P.u1img[l] = reinterpret_cast<uint8_t *>(take(ti));
      P.u2img[l] = reinterpret_cast<uint8_t *>(take(ti));
      P.x1img[l] = reinterpret_cast<uint8_t *>(take(ti));
      P.ctximg[l] = reinterpret_cast<uint8_t *>(take(ti));
      P.himg[l] = reinterpret_cast<uint8_t *>(take(th));
      P.ln1stat[l] = reinterpret_cast<float2 *>(take((int64_t)n_tiles * 128 * 8));
      P.ln2stat[l] = reinterpret_cast<float2 *>(take((int64_t)n_tiles * 128 * 8));
      if (dh == 128) P.qkvimg[l] = reinterpret_cast<uint8_t *>(take((int64_t)n_tiles * T256_G * T256_QKV_GROUP_IMG));
    }
    if (!fused) P.d_hvo = reinterpret_cast<float *>(take(M * c.e_tgt * 4));
    P.loss_partials = reinterpret_cast<float *>(take((loss_scratch_floats(n_seq) + edge256_loss_partials()) * 4));
    P.dlog = reinterpret_cast<float *>(take(M * c.e_tgt * 4));
    if (!fused) {
      P.dxrm = reinterpret_cast<float *>(take(M * d * 4));
      P.dxrm2 = reinterpret_cast<float *>(take(M * d * 4));
      P.g0 = reinterpret_cast<float *>(take(M * d * 4));
    } else {
      P.dlimg = reinterpret_cast<uint8_t *>(take((int64_t)n_tiles * 8192));
      P.srcimg = reinterpret_cast<uint8_t *>(take((int64_t)n_tiles * 8192));
      P.edge_scratch = reinterpret_cast<float *>(take((int64_t)(2 * 256 * 32 + 64) * 4));
    }
    P.dxa = reinterpret_cast<float *>(take(tf));
    P.dxb = reinterpret_cast<float *>(take(tf));
    P.da2img = reinterpret_cast<uint8_t *>(take(ti));
    P.da1img = reinterpret_cast<uint8_t *>(take(ti));
    P.dhimg = reinterpret_cast<uint8_t *>(take(th));
    P.dqkvimg = reinterpret_cast<uint8_t *>(take(3 * ti));
    P.dctx_scratch = reinterpret_cast<uint8_t *>(take((int64_t)160 * T256_TILE_IMG));
    P.wg_jobs = reinterpret_cast<uint8_t *>(take(T256_WG_JOBBUF));
  } else {
    uint8_t *ia = reinterpret_cast<uint8_t *>(take(ti)), *ib = reinterpret_cast<uint8_t *>(take(ti));
    for (int l = 0; l <= c.n_enc; ++l) P.ximg[l] = (l & 1) ? ib : ia;
  }
  P.bytes = off;
}

struct T256Ctx {
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

static int t256_check(const gt_config &c, int64_t n_seq, int mode, void *ws, int64_t ws_bytes, T256Plan &pl) {
  std::string why;
  GT_CHECK(t256_shape_supported(c, &why), "precision=bf16 is not available for this configuration (" + why + "); use precision=fp32");
  GT_CHECK(ws != nullptr && ((uintptr_t)ws & 255) == 0, "workspace must be non-null and 256-byte aligned");
  t256_make_plan(c, n_seq, mode, (char *)ws, pl);
  GT_CHECK(ws_bytes >= pl.bytes, "workspace too small: need " + std::to_string(pl.bytes) + " bytes");
  return 0;
}

int64_t t256_workspace_bytes(const gt_config &c, int64_t n_seq, int mode) {
  std::string why;
  if (!t256_shape_supported(c, &why)) {
    set_error("precision=bf16 is not available for this configuration (" + why + "); use precision=fp32");
    return -1;
  }
  static thread_local T256Plan pl;
  t256_make_plan(c, n_seq, mode, nullptr, pl);
  return pl.bytes;
}

static T256Args t256_layer_args(const T256Ctx &x, const T256Plan &pl, int l) {
  T256Args a;
  memset(&a, 0, sizeof(a));
  const LayerP &p = x.L->enc[l];
  a.img = pl.img + (size_t)l * pl.img_stride;
  a.img_rep_stride = pl.img_stride / T256_REP;
  a.bqkv = x.P + p.sa.b_in; a.bo = x.P + p.sa.b_out; a.b1 = x.P + p.b1; a.b2 = x.P + p.b2;
  a.g1 = x.P + p.g1; a.be1 = x.P + p.be1; a.g2 = x.P + p.g2; a.be2 = x.P + p.be2;
  if (x.G) {
    a.gbqkv = x.G + p.sa.b_in; a.gbo = x.G + p.sa.b_out; a.gb1 = x.G + p.b1; a.gb2 = x.G + p.b2;
    a.gg1 = x.G + p.g1; a.gbe1 = x.G + p.be1; a.gg2 = x.G + p.g2; a.gbe2 = x.G + p.be2;
  }
  a.M = x.M; a.n_tiles = pl.n_tiles;
  a.F = x.c.dim_ff; a.H = x.c.nhead; a.dh = x.c.d_model / x.c.nhead;
  a.d_attn = x.drop(site_id(0, l, 0)); a.d1 = x.drop(site_id(0, l, 1)); a.d_ffn = x.drop(site_id(0, l, 2));
  a.d2 = x.drop(site_id(0, l, 3));
  a.seq0 = x.seq0;
  return a;
}

static int t256_prep(const T256Ctx &x, const T256Plan &pl) {
  TcPrepArgs a;
  memset(&a, 0, sizeof(a));
  a.params = x.P; a.img = pl.img; a.img_stride = pl.img_stride; a.n_layers = x.c.n_enc;
  a.D = x.c.d_model; a.F = x.c.dim_ff; a.FC = 64; a.dh = x.c.d_model / x.c.nhead;
  for (int l = 0; l < x.c.n_enc; ++l) {
    a.w_in[l] = x.L->enc[l].sa.w_in; a.w_out[l] = x.L->enc[l].sa.w_out; a.w1[l] = x.L->enc[l].w1; a.w2[l] = x.L->enc[l].w2;
  }
  return t256_prep_weights(a, x.st);
}

// y != nullptr: the tail also evaluates calculate_loss (metrics6) and leaves dL/dlogits in pl.dlog (fused edges only)
static int t256_forward_all(const T256Ctx &x, const T256Plan &pl, const float *src, float *hvo, bool save, float thres,
                            const float *y = nullptr, float penalty = 0.f, float *metrics6 = nullptr) {
  GT_NVTX("groove.forward");
  const int d = x.c.d_model, L = x.c.n_enc;
  const bool fused = t256_fused_edges(x.c);
  GT_TRY(t256_prep(x, pl));
  if (fused) {
    GT_TRY(edge256_stem_fwd(src, x.c.e_src, x.P + x.L->in_enc_w, x.P + x.L->in_enc_b, x.pe, pl.ximg[0], x.M, pl.n_tiles,
                            x.drop(SITE_IN_ENC), x.seq0 * T, x.st));
  } else {
    GemmEpi e; e.bias = x.P + x.L->in_enc_b; e.relu = 1;
    GT_TRY(gemm_f32(src, x.c.e_src, 1, x.P + x.L->in_enc_w, x.c.e_src, 1, pl.r0, d, x.M, d, x.c.e_src, e, 0, x.st));
    GT_TRY(pe_dropout_fwd(pl.r0, x.pe, pl.x0rm, x.M, d, x.drop(SITE_IN_ENC), x.seq0 * T, x.st));
    GT_TRY(t256_to_image(pl.x0rm, pl.ximg[0], x.M, pl.n_tiles, x.st));
  }
  for (int l = 0; l < L; ++l) {
    T256Args a = t256_layer_args(x, pl, l);
    a.x_img_in = pl.ximg[l]; a.x_img_out = pl.ximg[l + 1];
    if (save) { a.u1_img = pl.u1img[l]; a.u2_img = pl.u2img[l]; a.x1_img = pl.x1img[l]; a.ctx_img = pl.ctximg[l]; a.h_img = pl.himg[l]; a.qkv_img = pl.qkvimg[l]; a.ln1_stat = pl.ln1stat[l]; a.ln2_stat = pl.ln2stat[l]; }
    GT_TRY(t256_layer_fwd(a, x.st));
  }
  if (fused)
    return edge256_tail_fwd(pl.ximg[L], x.P + x.L->enc_norm_g, x.P + x.L->enc_norm_b, x.P + x.L->out_w, x.P + x.L->out_b, hvo, pl.mf, pl.rf,
                            x.M, pl.n_tiles, thres, y, penalty, pl.dlog, pl.loss_partials, metrics6, x.st);
  GT_CHECK(y == nullptr, "t256_forward_all: fused loss needs the fused edge kernels");
  GT_TRY(t256_from_image(pl.ximg[L], pl.xLrm, x.M, x.st));
  Drop none;
  GT_TRY(ln_fwd(pl.xLrm, nullptr, x.P + x.L->enc_norm_g, x.P + x.L->enc_norm_b, nullptr, pl.z, pl.mf, pl.rf, x.M, d, none, 0, x.st));
  GemmEpi eh; eh.bias = x.P + x.L->out_b;
  GT_TRY(gemm_f32(pl.z, d, 1, x.P + x.L->out_w, d, 1, hvo, x.c.e_tgt, x.M, x.c.e_tgt, d, eh, 0, x.st));
  return head_activation(hvo, x.M, x.c.e_tgt, thres, x.st);
}

static int t256_wgrad_f32(const T256Ctx &x, const float *dY, int64_t N, const float *X, int64_t K, float *dW, float *db) {
  GemmEpi e; e.atomic = 1;
  GT_TRY(gemm_f32(dY, 1, N, X, 1, K, dW, K, N, K, x.M, e, 2048, x.st));
  return colsum_f32(dY, N, x.M, (int)N, db, x.st);
}

static int t256_backward_all(const T256Ctx &x, const T256Plan &pl, const float *src, const float *hvo, const float *d_hvo) {
  GT_NVTX("groove.backward");
  const int d = x.c.d_model, E = x.c.e_tgt, L = x.c.n_enc;
  const bool fused = t256_fused_edges(x.c);
  Drop none;
  if (fused) {
    // hvo == nullptr: d_hvo already holds dL/dlogits (left by the fused tail + loss forward)
    GT_TRY(edge256_tail_bwd(d_hvo, hvo, pl.ximg[L], pl.mf, pl.rf, x.P + x.L->enc_norm_g, x.P + x.L->enc_norm_b, x.P + x.L->out_w, pl.dxa,
                            pl.dlimg, pl.edge_scratch, pl.wg_jobs, x.G + x.L->out_w, x.G + x.L->out_b, x.G + x.L->enc_norm_g,
                            x.G + x.L->enc_norm_b, x.M, pl.n_tiles, x.st));
    grad_bucket_ready(x.c, BK_HEAD, 0, x.st);
  } else {
    GT_TRY(head_activation_bwd(d_hvo, hvo, pl.dlog, x.M, E, x.st));
    GT_TRY(t256_wgrad_f32(x, pl.dlog, E, pl.z, d, x.G + x.L->out_w, x.G + x.L->out_b));
    GemmEpi e0;
    GT_TRY(gemm_f32(pl.dlog, E, 1, x.P + x.L->out_w, 1, d, pl.dxrm, d, x.M, d, E, e0, 0, x.st));
    GT_TRY(ln_bwd(pl.dxrm, pl.xLrm, pl.mf, pl.rf, x.P + x.L->enc_norm_g, pl.dxrm2, nullptr, x.G + x.L->enc_norm_g,
                  x.G + x.L->enc_norm_b, x.M, d, none, 0, x.st));
    grad_bucket_ready(x.c, BK_HEAD, 0, x.st);
    GT_TRY(t256_to_tiled(pl.dxrm2, pl.dxa, x.M, pl.n_tiles, x.st));
  }
  float *cur = pl.dxa, *oth = pl.dxb;
  for (int l = L - 1; l >= 0; --l) {
    const LayerP &p = x.L->enc[l];
    T256Args a = t256_layer_args(x, pl, l);
    a.x_img_in = pl.ximg[l]; a.u1_img = pl.u1img[l]; a.u2_img = pl.u2img[l]; a.ln1_stat = pl.ln1stat[l]; a.ln2_stat = pl.ln2stat[l]; a.dy = cur; a.dx = oth;
    a.x1_img = pl.x1img[l]; a.ctx_img = pl.ctximg[l]; a.h_img = pl.himg[l]; a.qkv_img = pl.qkvimg[l];
    a.da2_img = pl.da2img; a.da1_img = pl.da1img; a.dh_img = pl.dhimg; a.dqkv_img = pl.dqkvimg; a.dctx_scratch = pl.dctx_scratch;
    GT_TRY(t256_layer_bwd(a, x.st));
    T256WgradArgs w;
    memset(&w, 0, sizeof(w));
    w.dqkv_img = pl.dqkvimg; w.x_img = pl.ximg[l]; w.da1_img = pl.da1img; w.ctx_img = pl.ctximg[l]; w.dh_img = pl.dhimg;
    w.x1_img = pl.x1img[l]; w.da2_img = pl.da2img; w.h_img = pl.himg[l];
    w.gwqkv = x.G + p.sa.w_in; w.gwo = x.G + p.sa.w_out; w.gw1 = x.G + p.w1; w.gw2 = x.G + p.w2;
    w.n_tiles = pl.n_tiles; w.F = x.c.dim_ff;
    GT_TRY(t256_wgrad(w, pl.wg_jobs, x.st));
    grad_bucket_ready(x.c, BK_ENC_LAYER, l, x.st);
    float *t = cur; cur = oth; oth = t;
  }
  if (fused) {
    // the g image reuses the da2 image buffer: the last layer-backward / weight-gradient launches that read it are done (stream order)
    GT_TRY(edge256_stem_bwd(cur, src, x.c.e_src, x.P + x.L->in_enc_w, x.P + x.L->in_enc_b, pl.da2img, pl.srcimg,
                            pl.edge_scratch + 256 * 32 + 64, pl.wg_jobs, x.G + x.L->in_enc_w, x.G + x.L->in_enc_b, x.M, pl.n_tiles,
                            x.drop(SITE_IN_ENC), x.seq0 * T, x.st));
  } else {
    GT_TRY(t256_from_tiled(cur, pl.dxrm, x.M, x.st));
    GT_TRY(pe_dropout_bwd(pl.dxrm, pl.r0, pl.g0, x.M, d, x.drop(SITE_IN_ENC), x.seq0 * T, x.st));
    GT_TRY(t256_wgrad_f32(x, pl.g0, d, src, x.c.e_src, x.G + x.L->in_enc_w, x.G + x.L->in_enc_b));
  }
  grad_bucket_ready(x.c, BK_IN_ENC, 0, x.st);
  return 0;
}

static void t256_ctx(T256Ctx &x, const gt_config &c, const Layout &L, const float *params, float *grads, const float *pe,
                     int64_t n_seq, bool train, uint64_t seed, uint64_t step, int64_t seq0, cudaStream_t st) {
  x.c = c; x.L = &L; x.P = params; x.G = grads; x.pe = pe; x.n_seq = n_seq; x.M = n_seq * T; x.train = train;
  x.seed = seed; x.step = step; x.seq0 = seq0; x.st = st;
}

int t256_forward(const gt_config &c, const Layout &L, const float *params, const float *pe, const float *src, int64_t n_seq,
                 float *hvo, void *ws, int64_t ws_bytes, bool train, uint64_t seed, uint64_t step, int64_t seq0, cudaStream_t st) {
  static thread_local T256Plan pl;
  GT_TRY(t256_check(c, n_seq, train ? 1 : 0, ws, ws_bytes, pl));
  T256Ctx x;
  t256_ctx(x, c, L, params, nullptr, pe, n_seq, train, seed, step, seq0, st);
  return t256_forward_all(x, pl, src, hvo, train, -1.f);
}

int t256_backward(const gt_config &c, const Layout &L, const float *params, const float *pe, const float *src, int64_t n_seq,
                  const float *hvo, const float *d_hvo, float *grads, void *ws, int64_t ws_bytes, uint64_t seed, uint64_t step,
                  int64_t seq0, cudaStream_t st) {
  static thread_local T256Plan pl;
  GT_TRY(t256_check(c, n_seq, 1, ws, ws_bytes, pl));
  T256Ctx x;
  t256_ctx(x, c, L, params, grads, pe, n_seq, true, seed, step, seq0, st);
  return t256_backward_all(x, pl, src, hvo, d_hvo);
}

int t256_train_step(const gt_config &c, const Layout &L, const float *params, const float *pe, const float *src, const float *y,
                    int64_t n_seq, float penalty, float *grads, float *metrics6, float *hvo, void *ws, int64_t ws_bytes,
                    uint64_t seed, uint64_t step, int64_t seq0, cudaStream_t st) {
  static thread_local T256Plan pl;
  GT_TRY(t256_check(c, n_seq, 1, ws, ws_bytes, pl));
  T256Ctx x;
  t256_ctx(x, c, L, params, grads, pe, n_seq, true, seed, step, seq0, st);
  GT_CUDA(cudaMemsetAsync(grads, 0, (size_t)L.total * sizeof(float), st));
  if (t256_fused_edges(c)) {
    GT_TRY(t256_forward_all(x, pl, src, hvo, true, -1.f, y, penalty, metrics6));
    return t256_backward_all(x, pl, src, nullptr, pl.dlog);
  }
  GT_TRY(t256_forward_all(x, pl, src, hvo, true, -1.f));
  GT_TRY(loss_fwd_bwd(hvo, y, n_seq, penalty, metrics6, pl.d_hvo, 1.f, pl.loss_partials, st, c.e_tgt / 3));
  return t256_backward_all(x, pl, src, hvo, pl.d_hvo);
}

int t256_predict(const gt_config &c, const Layout &L, const float *params, const float *pe, const float *src, int64_t n_seq,
                 float thres, float *hvo_out, void *ws, int64_t ws_bytes, cudaStream_t st) {
  static thread_local T256Plan pl;
  GT_TRY(t256_check(c, n_seq, 0, ws, ws_bytes, pl));
  T256Ctx x;
  t256_ctx(x, c, L, params, nullptr, pe, n_seq, false, 0, 0, 0, st);
  return t256_forward_all(x, pl, src, hvo_out, false, thres);
}

}  // namespace gt
