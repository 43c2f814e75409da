// tc_layers.cu — fused per-layer kernels of the bf16 tensor-core path.  See tc_layers.cuh.
#include "tc_layers.cuh"
#include "tc256.cuh"
#include "tc_attn32.cuh"
#include "umma.cuh"

namespace gt {
using namespace umma;

bool tc_shape_supported(const gt_config &c, std::string *why) {
  auto no = [&](const char *m) { if (why) *why = m; return false; };
  if (c.d_model == 256) return t256_shape_supported(c, why);
  if (c.n_dec != 0) return no("encoder-decoder models run in precision=fp32 (the fused tcgen05 layer kernels cover the encoder stack)");
  if (c.d_model != 32) return no("fused tcgen05 layer kernels are instantiated for d_model=32 and d_model=256");
  if (c.n_enc > TC_MAX_LAYERS) return no("more than 16 layers");
  if (c.dim_ff % 16 != 0 || c.dim_ff > 512) return no("dim_feedforward must be a multiple of 16 and <= 512 (TMEM-resident weight gradients)");
  if (tc_ffn_chunk(c.dim_ff) == 0) return no("dim_feedforward has no valid chunking");
  int dh = c.d_model / c.nhead;
  if (dh != 1 && (dh & 1)) return no("head dim must be 1 or even");
  return true;
}

// =============================================================================================
// weight prep: fp32 master parameters -> bf16 canonical K-major operand images (one block per layer)
// =============================================================================================
__device__ __forceinline__ uint4 pack8(const float *s) {
  float4 a = *reinterpret_cast<const float4 *>(s), b = *reinterpret_cast<const float4 *>(s + 4);
  return make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
}

__global__ void tc_prep_kernel(TcPrepArgs a) {
  // launch_pdl (common.cuh): the parameters are the optimizer kernel's output.  No early launch_dependents here: the layer kernels
  // read THESE images before their own griddepcontrol.wait, so whatever follows may only start when this kernel's CTAs have exited
  griddep_wait();
  const int l = blockIdx.y;
  const int D = a.D, F = a.F, FC = a.FC;
  const TcImg o = tc_img(D, F);
  uint8_t *img = a.img + (size_t)l * a.img_stride;
  const int kb_d = D / 8, kb_f = F / 8;
  const int n_qkv = 3 * D * kb_d, n_wo = D * kb_d, n_w1 = F * kb_d, n_w2 = D * kb_f, n_b1 = F * (TC_KAUG / 8);
  const int total = n_qkv + n_wo + n_w1 + n_w2 + n_b1;
  const uint32_t w1_chunk = tc_w1_chunk_bytes(D, FC);
  for (int id = blockIdx.x * blockDim.x + threadIdx.x; id < total; id += gridDim.x * blockDim.x) {
    int i = id;
    if (i < n_qkv) {
      int r = i / kb_d, kb = i % kb_d;
      *reinterpret_cast<uint4 *>(img + o.wqkv + kmajor_off(r, kb * 8, 3 * D)) = pack8(a.params + a.w_in[l] + (int64_t)r * D + kb * 8);
      continue;
    }
    i -= n_qkv;
    if (i < n_wo) {
      int r = i / kb_d, kb = i % kb_d;
      *reinterpret_cast<uint4 *>(img + o.wo + kmajor_off(r, kb * 8, D)) = pack8(a.params + a.w_out[l] + (int64_t)r * D + kb * 8);
      continue;
    }
    i -= n_wo;
    if (i < n_w1) {                     // W1 [F, D]: chunk c = rows c*FC.. -> image [FC x D]
      int gr = i / kb_d, kb = i % kb_d, c = gr / FC, r = gr % FC;
      *reinterpret_cast<uint4 *>(img + o.w1 + (size_t)c * w1_chunk + kmajor_off(r, kb * 8, FC)) =
          pack8(a.params + a.w1[l] + (int64_t)gr * D + kb * 8);
      continue;
    }
    i -= n_w1;
    if (i < n_b1) {                     // the TC_KAUG bias columns of every W1 chunk image: column D = bf16(b1), the rest 0
      int gr = i / (TC_KAUG / 8), sl = i % (TC_KAUG / 8), c = gr / FC, r = gr % FC;
      const uint32_t w0 = sl == 0 ? pack_bf16(a.params[a.b1[l] + gr], 0.f) : 0u;
      *reinterpret_cast<uint4 *>(img + o.w1 + (size_t)c * w1_chunk + kmajor_off(r, D + sl * 8, FC)) = make_uint4(w0, 0u, 0u, 0u);
      continue;
    }
    i -= n_b1;
    {                                   // W2 [D, F]: chunk c = columns c*FC.. -> image [D x FC]
      int j = i / kb_f, gk = (i % kb_f) * 8, c = gk / FC, kk = gk % FC;
      *reinterpret_cast<uint4 *>(img + o.w2 + (size_t)c * D * FC * 2 + kmajor_off(j, kk, D)) =
          pack8(a.params + a.w2[l] + (int64_t)j * F + gk);
    }
  }
}

int tc_prep_weights(const TcPrepArgs &a, cudaStream_t st) {
  dim3 grid(16, a.n_layers);
  { LaunchScope _ls(KC_TC_PREP, st);
    GT_CUDA(launch_pdl(tc_prep_kernel, grid, dim3(256), 0, st, a)); }
  GT_CUDA(cudaGetLastError());
  return 0;
}

// =============================================================================================
// shared device helpers
// =============================================================================================
// A operand: image of a [128 x K] tile.  B operand: image of an [N x K] weight (chunk).
__device__ __forceinline__ uint64_t desc_a128(uint32_t base, int k16) { return make_desc(base + (uint32_t)k16 * 2u * 2048u, 2048u, 128u); }
__device__ __forceinline__ uint64_t desc_b(uint32_t base, int N, int k16) {
  const uint32_t kstride = (uint32_t)(N >> 3) * 128u;
  return make_desc(base + (uint32_t)k16 * 2u * kstride, kstride, 128u);
}
// rows [n0, n0 + N') of an [N x K] weight image as a B operand (N' goes into the instruction descriptor)
__device__ __forceinline__ uint64_t desc_b_rows(uint32_t base, int N, int n0, int k16) {
  const uint32_t kstride = (uint32_t)(N >> 3) * 128u;
  return make_desc(base + (uint32_t)(n0 >> 3) * 128u + (uint32_t)k16 * 2u * kstride, kstride, 128u);
}
// four dropout decisions from one 64-bit hash (idx4 must be a multiple of 4): returns scale or 0 for each
__device__ __forceinline__ void drop4(const Drop &d, uint64_t idx4, float &m0, float &m1, float &m2, float &m3) {
  if (d.thr == 0) { m0 = 1.f; m1 = 1.f; m2 = 1.f; m3 = 1.f; return; }
  const uint64_t w = idx4 >> 2;
  uint32_t lo, hi;
  hash_quad((uint32_t)w ^ ((uint32_t)(w >> 32) * 0x85EBCA6Bu), d.key, lo, hi);
  m0 = ((lo & 0xFFFFu) >= d.thr) ? d.scale : 0.f;
  m1 = ((lo >> 16) >= d.thr) ? d.scale : 0.f;
  m2 = ((hi & 0xFFFFu) >= d.thr) ? d.scale : 0.f;
  m3 = ((hi >> 16) >= d.thr) ? d.scale : 0.f;
}

struct SmemPlan {
  uint32_t w, par, stat, xa, q, k, v, ctx, h, h2, total;
};
__host__ __device__ inline uint32_t al128(uint32_t x) { return (x + 127u) & ~127u; }
// mma = attention on mma.sync (tc_attn32.cuh): q | k | v are bf16 row-major token images instead of fp32 rows / compact heads
__host__ __device__ inline constexpr bool attn_mma(int DH) { return DH == 2 || DH == 4 || DH == 8; }
__host__ __device__ inline SmemPlan fwd_smem(int D, int F, int FC, bool mma) {
  SmemPlan s;
  s.w = 0;
  s.par = al128(tc_img(D, F).total);
  s.stat = al128(s.par + (uint32_t)(9 * D + F) * 4u);      // LayerNorm row statistics exchanged between the 4 column parts: [128 rows][4] float2
  s.xa = al128(s.stat + 128u * 4u * 8u);
  s.q = al128(s.xa + 128u * (D + TC_KAUG) * 2u);     // (xa: x image + the TC_KAUG ones / zero columns) ; q: fp32 rows, stride D+4
  if (mma) {
    s.k = s.q + A32_IMG;
    s.v = s.k + A32_IMG;
    s.ctx = al128(s.v + A32_IMG);
  } else {
    s.k = al128(s.q + 128u * (D + 4u) * 4u);         // fp32 k, compact per (sequence, head): [4][H][32][dh]
    s.v = s.k + 128u * D * 4u;
    s.ctx = al128(s.v + 128u * D * 4u);
  }
  s.h = al128(s.ctx + 128u * D * 2u);
  s.h2 = al128(s.h + 128u * (uint32_t)(FC < 32 ? 32 : FC) * 2u);        // >= 8 KB: the cross-attention memory tile is staged in h
  s.total = al128(s.h2 + 128u * (uint32_t)(FC < 32 ? 32 : FC) * 2u);    // h | h2: the hidden image is double buffered across FFN chunks
  return s;
}

// ---- attention score / context inner products for one (sequence, head) pair; lane = query row --------------
// K and V of the pair are 32 x DH fp32, contiguous and 16-byte aligned (read as warp-wide broadcasts).
template <int DH>
__device__ __forceinline__ void attn_scores(const float *q, const float *kh, int dh, float qs, float (&sc)[32]) {
  if constexpr (DH == 2) {
    const float2 q2 = *reinterpret_cast<const float2 *>(q);
    const float q0 = q2.x * qs, q1 = q2.y * qs;
    const float4 *k4 = reinterpret_cast<const float4 *>(kh);
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) {
      const float4 kk = k4[jj];
      sc[2 * jj] = fmaf(q1, kk.y, q0 * kk.x);
      sc[2 * jj + 1] = fmaf(q1, kk.w, q0 * kk.z);
    }
  } else if constexpr (DH > 0 && DH % 4 == 0) {
    float qv[DH];
#pragma unroll
    for (int c = 0; c < DH; c += 4) {
      const float4 t = *reinterpret_cast<const float4 *>(q + c);
      qv[c] = t.x * qs; qv[c + 1] = t.y * qs; qv[c + 2] = t.z * qs; qv[c + 3] = t.w * qs;
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < DH; c += 4) {
        const float4 kk = *reinterpret_cast<const float4 *>(kh + j * DH + c);
        acc = fmaf(qv[c], kk.x, acc); acc = fmaf(qv[c + 1], kk.y, acc); acc = fmaf(qv[c + 2], kk.z, acc); acc = fmaf(qv[c + 3], kk.w, acc);
      }
      sc[j] = acc;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) sc[j] = 0.f;
    for (int c = 0; c < dh; ++c) {
      const float qc = q[c] * qs;
#pragma unroll
      for (int j = 0; j < 32; ++j) sc[j] = fmaf(qc, kh[j * dh + c], sc[j]);
    }
  }
}
// out[c] = sum_j p[j] * vh[j][c]; written as bf16 into the K-major image `img` at (row, col0 + c)
template <int DH>
__device__ __forceinline__ void attn_context(const float (&p)[32], const float *vh, int dh, uint8_t *img, int row, int col0) {
  if constexpr (DH == 2) {
    const float4 *v4 = reinterpret_cast<const float4 *>(vh);
    float o0 = 0.f, o1 = 0.f;
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) {
      const float4 vv = v4[jj];
      o0 = fmaf(p[2 * jj], vv.x, o0); o1 = fmaf(p[2 * jj], vv.y, o1);
      o0 = fmaf(p[2 * jj + 1], vv.z, o0); o1 = fmaf(p[2 * jj + 1], vv.w, o1);
    }
    *reinterpret_cast<uint32_t *>(img + kmajor_off(row, col0, 128)) = pack_bf16(o0, o1);
  } else if constexpr (DH > 0 && DH % 4 == 0) {
    float o[DH];
#pragma unroll
    for (int c = 0; c < DH; ++c) o[c] = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
#pragma unroll
      for (int c = 0; c < DH; c += 4) {
        const float4 vv = *reinterpret_cast<const float4 *>(vh + j * DH + c);
        o[c] = fmaf(p[j], vv.x, o[c]); o[c + 1] = fmaf(p[j], vv.y, o[c + 1]); o[c + 2] = fmaf(p[j], vv.z, o[c + 2]); o[c + 3] = fmaf(p[j], vv.w, o[c + 3]);
      }
    }
#pragma unroll
    for (int c = 0; c < DH; c += 2) *reinterpret_cast<uint32_t *>(img + kmajor_off(row, col0 + c, 128)) = pack_bf16(o[c], o[c + 1]);
  } else {
    for (int c = 0; c < dh; ++c) {
      float o = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) o = fmaf(p[j], vh[j * dh + c], o);
      *reinterpret_cast<__nv_bfloat16 *>(img + kmajor_off(row, col0 + c, 128)) = __float2bfloat16_rn(o);
    }
  }
}
// write one token row's 32 k (or v) values (all heads) into the compact per-(sequence, head) layout
template <int D, int DH>
__device__ __forceinline__ void store_kv_row(float *dst, const float (&v)[D], int row, int H, int dh) {
  const int s = row >> 5, j = row & 31;
  float *base = dst + s * 32 * D;                    // [H][32][dh] of this sequence
  if constexpr (DH == 2) {
#pragma unroll
    for (int h = 0; h < D / 2; ++h) *reinterpret_cast<float2 *>(base + h * 64 + j * 2) = make_float2(v[2 * h], v[2 * h + 1]);
  } else if constexpr (DH > 0 && DH % 4 == 0) {
#pragma unroll
    for (int h = 0; h < D / DH; ++h)
#pragma unroll
      for (int c = 0; c < DH; c += 4)
        *reinterpret_cast<float4 *>(base + h * 32 * DH + j * DH + c) = make_float4(v[h * DH + c], v[h * DH + c + 1], v[h * DH + c + 2], v[h * DH + c + 3]);
  } else {
#pragma unroll
    for (int c = 0; c < D; ++c) {
      const int h = c / dh, cc = c - h * dh;
      base[h * 32 * dh + j * dh + cc] = v[c];
    }
  }
  (void)H;
}

// The K columns [D, D + TC_KAUG) of an FFN-input A image: column D = 1.0 (multiplies the bias row of the W1 chunk images), rest 0.
// Written once per kernel: the tile loop only rewrites columns [0, D).
__device__ __forceinline__ void write_kaug_columns(uint8_t *img, int tid, int D) {
  if (tid < 128 * (TC_KAUG / 8)) {
    const int r = tid & 127, sl = tid >> 7;
    *reinterpret_cast<uint4 *>(img + kmajor_off(r, D + sl * 8, 128)) = make_uint4(sl == 0 ? 0x00003F80u : 0u, 0u, 0u, 0u);
  }
}
// FFN hidden epilogue of one 16-column block: the accumulator already holds x1 W1^T + b1 (TC_KAUG); ReLU is fused into the bf16
// pack, the dropout decisions are packed compares ANDed onto the pairs.  The 1 / (1 - p) factor is NOT applied here: it is one
// multiply per OUTPUT column of linear2 (forward: on the FFN2 accumulator; backward: folded into the da2 image).
__device__ __forceinline__ void ffn_hidden_block(uint32_t taddr, uint8_t *dst_row /* image + kmajor_off(row, cb, 128) */, const Drop &d,
                                                 uint32_t wlo, uint32_t xhi) {
  float v[16];
  tmem_ld16(taddr, v);
  tmem_ld_wait();
  uint32_t pk[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) pk[j] = pack_bf16_relu(v[2 * j], v[2 * j + 1]);
  if (d.thr) {
    const uint32_t thr2 = d.thr | (d.thr << 16);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t lo, hi;
      hash_quad((wlo + (uint32_t)q) ^ xhi, d.key, lo, hi);
      pk[2 * q] &= keep2(lo, thr2);
      pk[2 * q + 1] &= keep2(hi, thr2);
    }
  }
  *reinterpret_cast<uint4 *>(dst_row) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  *reinterpret_cast<uint4 *>(dst_row + 2048) = make_uint4(pk[4], pk[5], pk[6], pk[7]);      // next 8-column slab of a 128-row image
}
// Column sums over the 128 token rows of `nslab` 8-column slabs of a K-major [128 x 8 nslab] bf16 image, on the tensor cores:
// warp w takes slabs w, w + 16, ...; per slab 8 x mma.m16n8k16 with an all-ones A operand (the B fragments come straight
// from ldmatrix.trans), fp32 sums land in lanes 0..3 and are added to dst[8 slab + ...] (shared-memory partials).
__device__ __forceinline__ void colsum_image(const uint8_t *img, int nslab, float *dst, int warp, int lane, int nwarps) {
  const uint32_t ones = 0x3F803F80u;
  for (int sl = warp; sl < nslab; sl += nwarps) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int kq = 0; kq < 4; ++kq) {
      uint32_t b[4];
      ldmatrix_x4_trans(b, img + (uint32_t)sl * 2048u + (uint32_t)(32 * kq + (lane & 7) + 8 * (lane >> 3)) * 16u);
      mma16816(acc, ones, ones, ones, ones, b[0], b[1]);
      mma16816(acc, ones, ones, ones, ones, b[2], b[3]);
    }
    if (lane < 4) { atomicAdd(dst + sl * 8 + 2 * lane, acc[0]); atomicAdd(dst + sl * 8 + 2 * lane + 1, acc[1]); }
  }
}

// =============================================================================================
// forward: x_in -> QKV (UMMA) -> attention (SIMT) -> out-proj (UMMA) -> +res, LN1 -> FFN1 (UMMA,
// chunked) -> relu/dropout -> FFN2 (UMMA, accumulating in TMEM) -> +res, LN2 -> x_out
// torch/nn/modules/transformer.py:951-956 (post-norm encoder layer)
// 512 threads: thread = (token row = tid & 127, column part = tid >> 7).
// =============================================================================================
constexpr int FWD_THREADS = 512;

// The kernels see their arguments through ArgsView: DEVSTEP = false (every eager launch) binds the parameter block as it lies in
// constant memory; DEVSTEP = true (CUDA-graph replay, gt_graph_train_create) takes a copy whose four dropout keys are derived
// from the device-resident step counter.  Keeping the eager instantiation free of the copy matters: with it the backward
// kernel was 3.3 % slower (14.35 vs 13.89 ms per C2 step).
template <bool DEVSTEP> struct ArgsView;
template <> struct ArgsView<false> {
  const TcLayerArgs &a;
  __device__ __forceinline__ explicit ArgsView(const TcLayerArgs &p) : a(p) {}
};
template <> struct ArgsView<true> {
  TcLayerArgs a;
  __device__ __forceinline__ explicit ArgsView(const TcLayerArgs &p) : a(p) {
    drop_resolve(a.d_attn); drop_resolve(a.d1); drop_resolve(a.d_ffn); drop_resolve(a.d2);
  }
};
__host__ inline bool args_devstep(const TcLayerArgs &a) {
  return a.d_attn.step_ptr != nullptr || a.d1.step_ptr != nullptr || a.d_ffn.step_ptr != nullptr || a.d2.step_ptr != nullptr;
}

// MODE 0: whole encoder layer.  MODE 1 (TC_MODE_FFN): the feed-forward block alone — x_out = LN(x_in + drop(FFN(x_in))) with
// the block's LayerNorm in g2 / be2 — used for the third block of a decoder layer (torch/nn/modules/transformer.py:1143-1153).
template <int D, int DH, int MODE, bool DEVSTEP = false>
__global__ void __launch_bounds__(FWD_THREADS, 1) tc_layer_fwd_kernel(const TcLayerArgs a_in) {
  static_assert(D == 32, "q|k|v epilogue assigns one 32-column projection per thread part");
  const ArgsView<DEVSTEP> view(a_in);
  const TcLayerArgs &a = view.a;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_w, bar_mma, bar_h;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = tid & 127, part = tid >> 7;
  const int F = a.F, FC = a.FC, H = DH > 0 ? D / DH : a.H, dh = DH > 0 ? DH : a.dh, nchunk = F / FC;
  constexpr bool MMA = attn_mma(DH);
  constexpr bool ATTN_ONLY = MODE >= TC_MODE_ATTN_CAUSAL, CAUSAL = MODE == TC_MODE_ATTN_CAUSAL, CROSS = MODE == TC_MODE_ATTN_CROSS;
  static_assert(!ATTN_ONLY || MMA, "attention-only blocks use the mma.sync attention path");
  const SmemPlan sp = fwd_smem(D, F, FC, MMA);
  const TcImg io = tc_img(D, F);
  uint8_t *sW = smem + sp.w;
  float *sPar = reinterpret_cast<float *>(smem + sp.par);
  float *p_bqkv = sPar, *p_bo = sPar + 3 * D, *p_b1 = p_bo + D, *p_b2 = p_b1 + F, *p_g1 = p_b2 + D, *p_be1 = p_g1 + D,
        *p_g2 = p_be1 + D, *p_be2 = p_g2 + D;
  uint8_t *sXa = smem + sp.xa;
  float *sQ = reinterpret_cast<float *>(smem + sp.q);
  float *sK = reinterpret_cast<float *>(smem + sp.k);
  float *sV = reinterpret_cast<float *>(smem + sp.v);
  uint8_t *sCtx = smem + sp.ctx;
  uint8_t *sH = smem + sp.h;
  float2 *sStat = reinterpret_cast<float2 *>(smem + sp.stat);
  constexpr int LQ = D + 4;
  const int c0 = part * 8;                             // this thread's 8 columns of its token row in the LayerNorm phases

  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) { mbar_init(&bar_w, 1); mbar_init(&bar_mma, 1); mbar_init(&bar_h, 1); fence_mbar_init(); }
  if constexpr (MODE != TC_MODE_FFN)
    for (int i = tid; i < 3 * D; i += FWD_THREADS) p_bqkv[i] = a.bqkv[i];
  if constexpr (!ATTN_ONLY) write_kaug_columns(sXa, tid, D);      // linear1's bias rides in the contraction (tc_layers.cuh: TC_KAUG)
  (void)p_b1;
  if (tid < D) {
    if constexpr (MODE != TC_MODE_FFN) { p_bo[tid] = a.bo[tid]; p_g1[tid] = a.g1[tid]; p_be1[tid] = a.be1[tid]; }
    if constexpr (!ATTN_ONLY) { p_b2[tid] = a.b2[tid]; p_g2[tid] = a.g2[tid]; p_be2[tid] = a.be2[tid]; }
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  if (tid == 0) {          // stage the layer's operand images with bulk TMA, one transaction barrier
    mbar_expect_tx(&bar_w, a.img_bytes);
    for (uint32_t off = 0; off < a.img_bytes; off += 32768u) {
      uint32_t n = a.img_bytes - off < 32768u ? a.img_bytes - off : 32768u;
      tma_load_1d(sW + off, a.img + off, n, &bar_w);
    }
  }
  const uint32_t tmem = tmem_slot;
  const uint32_t t_big = tmem, t_small = tmem + 128, t_big2 = tmem + 256;      // t_big | t_big2: FFN1 accumulators of even / odd chunks
  const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t aXa = smem_u32(sXa), aCtx = smem_u32(sCtx), aH = smem_u32(sH), aW = smem_u32(sW);
  mbar_wait(&bar_w, 0);
  uint32_t ph = 0, phh = 0;
  // Row statistics of a LayerNorm: every thread holds 8 of its row's 32 columns; the 4 warps that share 32 token rows (warp & 3
  // equal: they also share the TMEM lanes) exchange (sum, sum of squares) through shared memory behind a 128-thread named
  // barrier — all 16 warps work in the LayerNorm phases and no block-wide barrier is added.
  auto row_stats = [&](float s1, float s2, float &mean, float &rstd) {
    sStat[row * 4 + part] = make_float2(s1, s2);
    named_bar_sync(1 + (warp & 3), 128);
    const float4 p01 = *reinterpret_cast<const float4 *>(sStat + row * 4), p23 = *reinterpret_cast<const float4 *>(sStat + row * 4 + 2);
    mean = ((p01.x + p01.z) + (p23.x + p23.z)) * (1.f / D);
    rstd = rsqrtf(fmaxf(((p01.y + p01.w) + (p23.y + p23.w)) * (1.f / D) - mean * mean, 0.f) + LN_EPS);
    named_bar_sync(1 + (warp & 3), 128);               // the four slots may be rewritten (next LayerNorm phase)
  };
  const float attn_scale = rsqrtf((float)dh) * 1.4426950408889634f;   // 1/sqrt(dh) * log2(e), folded into q
  const float keep_scale = a.d_attn.scale;

  // this thread's 8-column chunk of the NEXT tile's input row (and memory row), fetched a whole tile time ahead: the tile
  // would otherwise start with a load of rows nobody has touched yet
  float4 xn0 = make_float4(0, 0, 0, 0), xn1 = xn0;
  uint4 mnext = make_uint4(0, 0, 0, 0);
  auto fetch_tile = [&](int t) {
    const int64_t r = (int64_t)t * TC_TILE + row;
    xn0 = make_float4(0, 0, 0, 0); xn1 = xn0; mnext = make_uint4(0, 0, 0, 0);
    if (t < a.n_tiles && r < a.M) {
      xn0 = *reinterpret_cast<const float4 *>(a.x_in + r * D + c0);
      xn1 = *reinterpret_cast<const float4 *>(a.x_in + r * D + c0 + 4);
      if constexpr (CROSS) mnext = pack8(a.mem + r * D + c0);
    }
  };
  // everything above touched only this layer's parameters; the activations below are the previous kernel's output (common.cuh: PDL)
  griddep_wait();
  griddep_launch();
  fetch_tile(blockIdx.x);
  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const int64_t grow = (int64_t)tile * TC_TILE + row;
    const bool valid = grow < a.M;
    // ---- P0: x_in tile -> bf16 A operand (8 columns per thread); the fp32 chunk stays in registers: it is this thread's residual ----
    float x1[8] = {xn0.x, xn0.y, xn0.z, xn0.w, xn1.x, xn1.y, xn1.z, xn1.w};     // x_in chunk now; LayerNorm1 output (FFN input / residual) after P5
    {
      *reinterpret_cast<uint4 *>(sXa + kmajor_off(row, c0, 128)) =
          make_uint4(pack_bf16(x1[0], x1[1]), pack_bf16(x1[2], x1[3]), pack_bf16(x1[4], x1[5]), pack_bf16(x1[6], x1[7]));
      if constexpr (CROSS) *reinterpret_cast<uint4 *>(sH + kmajor_off(row, c0, 128)) = mnext;   // keys / values come from the encoder memory tile
      fetch_tile(tile + (int)gridDim.x);
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    if constexpr (MODE != TC_MODE_FFN) {
    // ---- P1: QKV = x Wqkv^T  (cross: q = x Wq^T, k | v = mem Wkv^T: rows [D, 3D) of the packed in-projection) ----
    if (tid == 0) {
      fence_after_sync();
      if constexpr (CROSS) {
        const uint32_t idq = make_idesc_bf16(128, D), idkv = make_idesc_bf16(128, 2 * D);
#pragma unroll
        for (int k = 0; k < D / 16; ++k) mma_bf16_ss(t_big, desc_a128(aXa, k), desc_b_rows(aW + io.wqkv, 3 * D, 0, k), idq, k > 0);
#pragma unroll
        for (int k = 0; k < D / 16; ++k) mma_bf16_ss(t_big + (uint32_t)D, desc_a128(aH, k), desc_b_rows(aW + io.wqkv, 3 * D, D, k), idkv, k > 0);
      } else {
        const uint32_t idesc = make_idesc_bf16(128, 3 * D);
#pragma unroll
        for (int k = 0; k < D / 16; ++k) mma_bf16_ss(t_big, desc_a128(aXa, k), desc_b(aW + io.wqkv, 3 * D, k), idesc, k > 0);
      }
      mma_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, ph); ph ^= 1;
    fence_after_sync();
    // ---- P2: + bias; part 0 -> q rows, part 1 -> k, part 2 -> v (compact per head) ----
    if (part < 3) {
      float v[D];
      tmem_ld16(t_big + lane_off + (uint32_t)(part * D), v);
      tmem_ld16(t_big + lane_off + (uint32_t)(part * D + 16), v + 16);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < D; c += 4) {
        const float4 b = *reinterpret_cast<const float4 *>(p_bqkv + part * D + c);
        v[c] += b.x; v[c + 1] += b.y; v[c + 2] += b.z; v[c + 3] += b.w;
      }
      if constexpr (MMA) {
        a32_store_row(smem + (part == 0 ? sp.q : part == 1 ? sp.k : sp.v), row, v, part == 0 ? attn_scale : 1.f);
      } else if (part == 0) {
#pragma unroll
        for (int c = 0; c < D; c += 4) *reinterpret_cast<float4 *>(sQ + row * LQ + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
      } else {
        store_kv_row<D, DH>(part == 1 ? sK : sV, v, row, H, dh);
      }
    }
    fence_before_sync();
    __syncthreads();
    // ---- P3: attention, one warp per (sequence, head) ----
    if constexpr (MMA) {
      for (int p = warp; p < 4 * H; p += FWD_THREADS / 32) {
        const int s = p / H, h = p - s * H;
        const uint64_t w_pair = (uint64_t)((((a.seq0 + (int64_t)tile * 4 + s) * H + h) * 32) * 8);
        a32_attn_fwd<DH, CAUSAL>(smem + sp.q, smem + sp.k, smem + sp.v, sCtx, s, h, lane, a.d_attn, w_pair);
      }
    } else
    for (int p = warp; p < 4 * H; p += FWD_THREADS / 32) {      // lane = query row
      const int s = p / H, h = p - s * H;
      const int r = s * 32 + lane;
      const float *kh = sK + s * 32 * D + h * 32 * dh;
      const float *vh = sV + s * 32 * D + h * 32 * dh;
      float sc[32];
      attn_scores<DH>(sQ + r * LQ + h * dh, kh, dh, attn_scale, sc);
      float mx = sc[0];
#pragma unroll
      for (int j = 1; j < 32; ++j) mx = fmaxf(mx, sc[j]);
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) { sc[j] = ex2_ftz(sc[j] - mx); sum += sc[j]; }
      const float inv = keep_scale / sum;
      if (a.d_attn.thr) {
        const uint64_t w0 = (uint64_t)((((a.seq0 + (int64_t)tile * 4 + s) * H + h) * 32 + lane) * 8);    // quad index (idx >> 2) of position 0
        const uint32_t xhi = (uint32_t)(w0 >> 32) * 0x85EBCA6Bu, wlo = (uint32_t)w0;
#pragma unroll
        for (int m = 0; m < 8; ++m) {                 // quad m = positions 4m..4m+3 = keys k0, k0+1, k0+8, k0+9 (common.cuh: key_perm)
          const int k0 = (m >> 2) * 16 + (m & 3) * 2;
          uint32_t lo, hi;
          hash_quad((wlo + (uint32_t)m) ^ xhi, a.d_attn.key, lo, hi);
          sc[k0] = ((lo & 0xFFFFu) >= a.d_attn.thr) ? sc[k0] * inv : 0.f;
          sc[k0 + 1] = ((lo >> 16) >= a.d_attn.thr) ? sc[k0 + 1] * inv : 0.f;
          sc[k0 + 8] = ((hi & 0xFFFFu) >= a.d_attn.thr) ? sc[k0 + 8] * inv : 0.f;
          sc[k0 + 9] = ((hi >> 16) >= a.d_attn.thr) ? sc[k0 + 9] * inv : 0.f;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) sc[j] *= inv;
      }
      attn_context<DH>(sc, vh, dh, sCtx, r, h * dh);
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    // ---- P4: out-proj ----
    if (tid == 0) {
      fence_after_sync();
      const uint32_t idesc = make_idesc_bf16(128, D);
#pragma unroll
      for (int k = 0; k < D / 16; ++k) mma_bf16_ss(t_small, desc_a128(aCtx, k), desc_b(aW + io.wo, D, k), idesc, k > 0);
      mma_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, ph); ph ^= 1;
    fence_after_sync();
    // ---- P5: + bias, dropout, + residual, LayerNorm1 (thread = (row, 8 columns)) ----
    {
      float u[8];
      tmem_ld8(t_small + lane_off + (uint32_t)c0, u);
      tmem_ld_wait();
      const uint64_t e0 = (uint64_t)((a.seq0 * 32 + grow) * D + c0);
      float m[8];
      drop4(a.d1, e0, m[0], m[1], m[2], m[3]);
      drop4(a.d1, e0 + 4, m[4], m[5], m[6], m[7]);
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        u[j] = x1[j] + (u[j] + p_bo[c0 + j]) * m[j];
        s1 += u[j]; s2 = fmaf(u[j], u[j], s2);
      }
      if (a.u1 && valid) {
        *reinterpret_cast<float4 *>(a.u1 + grow * D + c0) = make_float4(u[0], u[1], u[2], u[3]);
        *reinterpret_cast<float4 *>(a.u1 + grow * D + c0 + 4) = make_float4(u[4], u[5], u[6], u[7]);
      }
      float mu, rs;
      row_stats(s1, s2, mu, rs);
#pragma unroll
      for (int j = 0; j < 8; ++j) x1[j] = (u[j] - mu) * rs * p_g1[c0 + j] + p_be1[c0 + j];
      if constexpr (ATTN_ONLY) {                       // attention block alone: its LayerNorm output is the block output
        if (valid) {
          *reinterpret_cast<float4 *>(a.x_out + grow * D + c0) = make_float4(x1[0], x1[1], x1[2], x1[3]);
          *reinterpret_cast<float4 *>(a.x_out + grow * D + c0 + 4) = make_float4(x1[4], x1[5], x1[6], x1[7]);
        }
      } else {
        *reinterpret_cast<uint4 *>(sXa + kmajor_off(row, c0, 128)) =
            make_uint4(pack_bf16(x1[0], x1[1]), pack_bf16(x1[2], x1[3]), pack_bf16(x1[4], x1[5]), pack_bf16(x1[6], x1[7]));
      }
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    }  // MODE != TC_MODE_FFN
    if constexpr (ATTN_ONLY) continue;
    // ---- P6: FFN, hidden dimension in chunks of FC; FFN2 accumulates in TMEM across chunks ----
    // FFN1 accumulators and hidden images are double buffered: FFN1(c + 1) is issued BEFORE the epilogue of chunk c starts, so
    // the tensor pipe works under the epilogue and no MMA round trip is exposed inside the loop.  One commit covers everything
    // issued before it: the wait for FFN1(c) also proves FFN2(c - 2) retired, i.e. hidden image (c & 1) may be rewritten.
    const uint32_t idesc1 = make_idesc_bf16(128, FC), idesc2 = make_idesc_bf16(128, D);
    const uint32_t w1_chunk = tc_w1_chunk_bytes(D, FC);
    const uint32_t h_bytes = sp.h2 - sp.h;
    if (tid == 0) {
      fence_after_sync();
#pragma unroll
      for (int k = 0; k < (D + TC_KAUG) / 16; ++k) mma_bf16_ss(t_big, desc_a128(aXa, k), desc_b(aW + io.w1, FC, k), idesc1, k > 0);
      mma_commit(&bar_h);
    }
    const int nblk = FC / 16;                          // 16-column blocks of the chunk, dealt round-robin to the 4 parts
    for (int c = 0; c < nchunk; ++c) {
      const uint32_t tb = (c & 1) ? t_big2 : t_big;
      mbar_wait(&bar_h, phh); phh ^= 1;               // FFN1(c) complete
      fence_after_sync();
      if (tid == 0 && c + 1 < nchunk) {
#pragma unroll
        for (int k = 0; k < (D + TC_KAUG) / 16; ++k)
          mma_bf16_ss((c & 1) ? t_big : t_big2, desc_a128(aXa, k), desc_b(aW + io.w1 + (uint32_t)(c + 1) * w1_chunk, FC, k), idesc1, k > 0);
        mma_commit(&bar_h);
      }
      {
        uint8_t *sHc = sH + (c & 1) * h_bytes;
        const uint64_t w0 = (uint64_t)((a.seq0 * 32 + grow) * F + c * FC) >> 2;        // multiple of 4 (F, FC multiples of 16)
        for (int b = part; b < nblk; b += 4) {
          const int cb = b * 16;
          const uint64_t wb = w0 + (uint64_t)(cb >> 2);                                 // + (0..3) never carries
          ffn_hidden_block(tb + lane_off + (uint32_t)cb, sHc + kmajor_off(row, cb, 128), a.d_ffn, (uint32_t)wb,
                           (uint32_t)(wb >> 32) * 0x85EBCA6Bu);
        }
      }
      fence_async_smem();
      fence_before_sync();
      __syncthreads();
      if (tid == 0) {
        fence_after_sync();
        const uint32_t aHc = aH + (uint32_t)(c & 1) * h_bytes;
        for (int k = 0; k < FC / 16; ++k)
          mma_bf16_ss(t_small, desc_a128(aHc, k), desc_b(aW + io.w2 + (uint32_t)c * D * FC * 2u, D, k), idesc2, (c | k) > 0);
        if (c + 1 == nchunk) mma_commit(&bar_mma);
      }
    }
    mbar_wait(&bar_mma, ph); ph ^= 1;                 // last FFN2 complete
    fence_after_sync();
    // ---- P8: + bias, dropout, + residual, LayerNorm2 -> x_out (thread = (row, 8 columns)) ----
    {
      float f[8];
      tmem_ld8(t_small + lane_off + (uint32_t)c0, f);
      tmem_ld_wait();
      const uint64_t e0 = (uint64_t)((a.seq0 * 32 + grow) * D + c0);
      const float fs = a.d_ffn.scale;                  // 1 / (1 - p) of the hidden dropout: applied here, once per output column
      float m[8];
      drop4(a.d2, e0, m[0], m[1], m[2], m[3]);
      drop4(a.d2, e0 + 4, m[4], m[5], m[6], m[7]);
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        f[j] = x1[j] + fmaf(f[j], fs, p_b2[c0 + j]) * m[j];
        s1 += f[j]; s2 = fmaf(f[j], f[j], s2);
      }
      if (a.u2 && valid) {
        *reinterpret_cast<float4 *>(a.u2 + grow * D + c0) = make_float4(f[0], f[1], f[2], f[3]);
        *reinterpret_cast<float4 *>(a.u2 + grow * D + c0 + 4) = make_float4(f[4], f[5], f[6], f[7]);
      }
      float mu, rs;
      row_stats(s1, s2, mu, rs);
      if (valid) {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = (f[j] - mu) * rs * p_g2[c0 + j] + p_be2[c0 + j];
        *reinterpret_cast<float4 *>(a.x_out + grow * D + c0) = make_float4(f[0], f[1], f[2], f[3]);
        *reinterpret_cast<float4 *>(a.x_out + grow * D + c0 + 4) = make_float4(f[4], f[5], f[6], f[7]);
      }
    }
    fence_before_sync();
    __syncthreads();
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int DH, int MODE = 0>
static int launch_fwd(const TcLayerArgs &a, uint32_t smem, int grid, cudaStream_t st) {
  if (args_devstep(a)) {                     // graph replay: dropout keys derived on the device from the step counter
    GT_CUDA(cudaFuncSetAttribute(tc_layer_fwd_kernel<32, DH, MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    { LaunchScope _ls(KC_TC_LAYER_FWD, st);
      GT_CUDA(launch_pdl(tc_layer_fwd_kernel<32, DH, MODE, true>, dim3(grid), dim3(FWD_THREADS), smem, st, a)); }
    GT_CUDA(cudaGetLastError());
    return 0;
  }
  GT_CUDA(cudaFuncSetAttribute(tc_layer_fwd_kernel<32, DH, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  { LaunchScope _ls(KC_TC_LAYER_FWD, st);
    GT_CUDA(launch_pdl(tc_layer_fwd_kernel<32, DH, MODE>, dim3(grid), dim3(FWD_THREADS), smem, st, a)); }
  GT_CUDA(cudaGetLastError());
  return 0;
}

int tc_layer_fwd(int D, const TcLayerArgs &a, cudaStream_t st) {
  GT_CHECK(D == 32, "tc_layer_fwd: d_model not instantiated");
  const SmemPlan sp = fwd_smem(D, a.F, a.FC, attn_mma(a.dh));
  GT_CHECK(sp.total <= 227 * 1024, "tc_layer_fwd: shared memory budget exceeded");
  int grid = a.n_tiles < num_sms() ? a.n_tiles : num_sms();
  if (a.mode == TC_MODE_FFN) return launch_fwd<2, TC_MODE_FFN>(a, fwd_smem(D, a.F, a.FC, true).total, grid, st);
  if (a.mode == TC_MODE_ATTN_CAUSAL || a.mode == TC_MODE_ATTN_CROSS) {
    const bool causal = a.mode == TC_MODE_ATTN_CAUSAL;
    switch (a.dh) {
      case 2: return causal ? launch_fwd<2, TC_MODE_ATTN_CAUSAL>(a, sp.total, grid, st) : launch_fwd<2, TC_MODE_ATTN_CROSS>(a, sp.total, grid, st);
      case 4: return causal ? launch_fwd<4, TC_MODE_ATTN_CAUSAL>(a, sp.total, grid, st) : launch_fwd<4, TC_MODE_ATTN_CROSS>(a, sp.total, grid, st);
      case 8: return causal ? launch_fwd<8, TC_MODE_ATTN_CAUSAL>(a, sp.total, grid, st) : launch_fwd<8, TC_MODE_ATTN_CROSS>(a, sp.total, grid, st);
      default: GT_FAIL("tc_layer_fwd: attention-only blocks need head dim 2, 4 or 8");
    }
  }
  switch (a.dh) {
    case 2: return launch_fwd<2>(a, sp.total, grid, st);
    case 4: return launch_fwd<4>(a, sp.total, grid, st);
    case 8: return launch_fwd<8>(a, sp.total, grid, st);
    default: return launch_fwd<0>(a, sp.total, grid, st);
  }
}

// =============================================================================================
// backward of one encoder layer.  Recomputes LN statistics, q|k|v, the attention probabilities and
// the FFN hidden activations tile by tile; ALL weight gradients of the layer accumulate in TMEM
// across the CTA's tiles (dW1, dW2^T, dWqkv, dWo: 320 columns) via MN-major UMMAs that contract
// over the 128 tokens of the tile, and are flushed once per CTA with fp32 atomics.
//
// TMEM map (512 columns): [0,128) working (q|k|v / H chunk / dH chunk)  [128,160) dx accumulators
// [160,192) dctx   [192,320) dW1 chunks   [320,448) dW2^T chunks   [448,480) dWqkv   [480,512) dWo
// =============================================================================================
struct BwdSmem {
  uint32_t w, par, gpar, stat, xin, x1, da, ctx, dq, q, k, v, dctx, h, dh, total;
};
__host__ __device__ inline BwdSmem bwd_smem(int D, int F, bool mma) {
  BwdSmem s;
  s.w = 0;
  s.par = al128(tc_img(D, F).total);
  s.gpar = al128(s.par + (uint32_t)(9 * D + F) * 4u);
  s.xin = al128(s.gpar + (uint32_t)(9 * D + F) * 4u);
  s.x1 = s.xin + 128u * D * 2u;
  s.da = s.x1 + 128u * (D + TC_KAUG) * 2u;            // (x1 image + the TC_KAUG ones / zero columns of the H recompute)
  s.ctx = s.da + 128u * D * 2u;
  s.dq = s.ctx + 128u * D * 2u;                       // dqkv image [128 x 3D]; M-padded reads run into the union below
  uint32_t u = al128(s.dq + 128u * 3u * D * 2u);
  s.q = u;                                            // attention phase: fp32 q rows, compact k / v, fp32 dctx rows
  uint32_t end_attn;
  if (mma) {                                          // ... or four bf16 row-major token images (tc_attn32.cuh)
    s.k = s.q + A32_IMG;
    s.v = s.k + A32_IMG;
    s.dctx = s.v + A32_IMG;
    end_attn = al128(s.dctx + A32_IMG);
  } else {
    s.k = al128(s.q + 128u * (D + 4u) * 4u);
    s.v = s.k + 128u * D * 4u;
    s.dctx = al128(s.v + 128u * D * 4u);
    end_attn = al128(s.dctx + 128u * (D + 4u) * 4u);
  }
  s.h = u;                                            // FFN phase (aliases the attention scratch): H and dH images
  s.dh = s.h + 128u * 128u * 2u;
  uint32_t end_ffn = s.dh + 128u * 128u * 2u;
  // row-statistics exchange of the two LayerNorm-backward phases (two ping-pong buffers of [128 rows][4 parts] float4 = 16 KB):
  // aliases the first half of the EVEN hidden image, which is dead in both phases — B0: start of a tile; B2: every FFN MMA of
  // the tile has retired, the last generic reads of that image (the dH epilogue of an even chunk) lie before a block-wide
  // barrier, and the attention images that share this range are written after B2's closing barrier.  (It must NOT alias the
  // dH image: the tensor-core column sum of the last chunk's dH still reads it when the first warps enter B2 — found by
  // compute-sanitizer racecheck, profiles/r02/r02_sanitizer_racecheck_bf16_before_fix.txt.)
  s.stat = s.h;
  s.total = end_attn > end_ffn ? end_attn : end_ffn;
  return s;
}

// per-lane vector of 32 values -> lane l receives the sum over the warp's lanes of v[l]  (31 shuffles)
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float send = up ? v[i] : v[i + s];
      const float keep = up ? v[i + s] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}
// LayerNorm-backward column sums of one warp (lane = token row): w[0..7] -> dgamma, w[8..15] -> dbeta, w[16..23] -> bias of the
// sub-layer, each for the thread's 8 feature columns.  On the tensor cores (umma.cuh: warp_colsum16_packed; inputs rounded to
// bf16 like every other column sum here): 12 conversions + 7 mma.sync instead of the 31-shuffle butterfly of warp_colsum32.
__device__ __forceinline__ void ln_bwd_colsums(const float (&w)[32], int lane, float *gg, float *gbe, float *gb) {
  const ColsumSel sel = colsum_sel(lane);
  uint32_t pa[8], pb[4];
#pragma unroll
  for (int j = 0; j < 8; ++j) pa[j] = pack_bf16(w[2 * j], w[2 * j + 1]);
#pragma unroll
  for (int j = 0; j < 4; ++j) pb[j] = pack_bf16(w[16 + 2 * j], w[16 + 2 * j + 1]);
  float s0, s1;
  warp_colsum16_packed(pa, sel, s0, s1);
  const float s2 = warp_colsum8_packed(pb, sel);
  const int g = lane >> 2, t = lane & 3;
  if (t == 0) atomicAdd(gg + g, s0);
  else if (t == 1) atomicAdd(gbe + g, s1);
  else if (t == 2) atomicAdd(gb + g, s2);
}
__device__ __forceinline__ float warp_sum_all(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// MN-major views of the row-major images (contract over the image's ROW index)
__device__ __forceinline__ uint64_t desc_mn(uint32_t base, int rows, int k16) {      // image [rows(k) x cols(mn)]
  return make_desc(base + (uint32_t)k16 * 256u, 128u, (uint32_t)(rows >> 3) * 128u);
}

constexpr int BWD_THREADS = 512;
constexpr int BWD_PARTS = BWD_THREADS / 128;

template <int D, int DH, int MODE, bool DEVSTEP = false>
__global__ void __launch_bounds__(BWD_THREADS, 1) tc_layer_bwd_kernel(const TcLayerArgs a_in) {
  static_assert(D == 32, "the register-tile column sums assume d_model == 32");
  const ArgsView<DEVSTEP> view(a_in);
  const TcLayerArgs &a = view.a;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_w, bar_mma, bar_h;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = tid & 127, part = tid >> 7;
  const int F = a.F, FC = a.FC, H = DH > 0 ? D / DH : a.H, dh = DH > 0 ? DH : a.dh, nchunk = F / FC;
  constexpr bool MMA = attn_mma(DH);
  constexpr bool ATTN_ONLY = MODE >= TC_MODE_ATTN_CAUSAL, CAUSAL = MODE == TC_MODE_ATTN_CAUSAL, CROSS = MODE == TC_MODE_ATTN_CROSS;
  static_assert(!ATTN_ONLY || MMA, "attention-only blocks use the mma.sync attention path");
  const BwdSmem sp = bwd_smem(D, F, MMA);
  const TcImg io = tc_img(D, F);
  uint8_t *sW = smem + sp.w;
  float *sPar = reinterpret_cast<float *>(smem + sp.par);
  float *p_bqkv = sPar, *p_bo = sPar + 3 * D, *p_b1 = p_bo + D, *p_b2 = p_b1 + F, *p_g1 = p_b2 + D, *p_be1 = p_g1 + D,
        *p_g2 = p_be1 + D;
  float *sG = reinterpret_cast<float *>(smem + sp.gpar);       // gradient partials, same order as sPar
  float *g_bqkv = sG, *g_bo = sG + 3 * D, *g_b1 = g_bo + D, *g_b2 = g_b1 + F, *g_g1 = g_b2 + D, *g_be1 = g_g1 + D,
        *g_g2 = g_be1 + D, *g_be2 = g_g2 + D;
  float4 *sEx = reinterpret_cast<float4 *>(smem + sp.stat);     // [2][128 rows][4 parts]: row-statistics exchange (ping-pong)
  int xphase = 0;
  // sum of a float4 over the four column parts of this token row: only the 4 warps that share these 32 rows (warp & 3 equal)
  // exchange, behind a 128-thread named barrier instead of a block-wide one (the ping-pong buffers make a second barrier unnecessary)
  auto xsum = [&](const float4 v) -> float4 {
    float4 *buf = sEx + xphase * 512;
    xphase ^= 1;
    buf[row * 4 + part] = v;
    named_bar_sync(1 + (warp & 3), 128);
    const float4 p0 = buf[row * 4], p1 = buf[row * 4 + 1], p2 = buf[row * 4 + 2], p3 = buf[row * 4 + 3];
    return make_float4((p0.x + p1.x) + (p2.x + p3.x), (p0.y + p1.y) + (p2.y + p3.y), (p0.z + p1.z) + (p2.z + p3.z), (p0.w + p1.w) + (p2.w + p3.w));
  };
  uint8_t *sXin = smem + sp.xin, *sX1 = smem + sp.x1, *sDA = smem + sp.da, *sCtx = smem + sp.ctx, *sDQ = smem + sp.dq;
  float *sQ = reinterpret_cast<float *>(smem + sp.q);
  float *sK = reinterpret_cast<float *>(smem + sp.k);
  float *sV = reinterpret_cast<float *>(smem + sp.v);
  float *sDC = reinterpret_cast<float *>(smem + sp.dctx);
  uint8_t *sH = smem + sp.h, *sDH = smem + sp.dh;
  constexpr int LQ = D + 4;

  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) { mbar_init(&bar_w, 1); mbar_init(&bar_mma, 1); mbar_init(&bar_h, 1); fence_mbar_init(); }
  for (int i = tid; i < 9 * D + F; i += BWD_THREADS) sG[i] = 0.f;
  if constexpr (MODE != TC_MODE_FFN)
    for (int i = tid; i < 3 * D; i += BWD_THREADS) p_bqkv[i] = a.bqkv[i];
  if constexpr (!ATTN_ONLY) write_kaug_columns(sX1, tid, D);      // linear1's bias rides in the H recompute (tc_layers.cuh: TC_KAUG)
  (void)p_b1;
  if (tid < D) {
    if constexpr (MODE != TC_MODE_FFN) { p_g1[tid] = a.g1[tid]; p_be1[tid] = a.be1[tid]; }
    if constexpr (!ATTN_ONLY) p_g2[tid] = a.g2[tid];
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  if (tid == 0) {
    mbar_expect_tx(&bar_w, a.img_bytes);
    for (uint32_t off = 0; off < a.img_bytes; off += 32768u) {
      uint32_t n = a.img_bytes - off < 32768u ? a.img_bytes - off : 32768u;
      tma_load_1d(sW + off, a.img + off, n, &bar_w);
    }
  }
  const uint32_t tmem = tmem_slot;
  const uint32_t t_big = tmem, t_sa = tmem + 128, t_sb = tmem + 160, t_dw1 = tmem + 192, t_dw2 = tmem + 320,
                 t_dwqkv = tmem + 448, t_dwo = tmem + 480;
  const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t aW = smem_u32(sW), aXin = smem_u32(sXin), aX1 = smem_u32(sX1), aDA = smem_u32(sDA), aCtx = smem_u32(sCtx),
                 aDQ = smem_u32(sDQ), aH = smem_u32(sH), aDH = smem_u32(sDH);
  mbar_wait(&bar_w, 0);
  uint32_t ph = 0, phh = 0;                          // parities of bar_mma / bar_h
  const float inv_sqrt_dh = rsqrtf((float)dh);
  const float attn_scale = inv_sqrt_dh * 1.4426950408889634f;
  const int nblk = FC / 16;
  // instruction descriptors
  const uint32_t id_kk_fc = make_idesc_bf16(128, FC, 0, 0);      // H chunk          A K-major, B K-major
  const uint32_t id_kmn_fc = make_idesc_bf16(128, FC, 0, 1);     // dH chunk         A K-major, B MN-major
  const uint32_t id_kmn_d = make_idesc_bf16(128, D, 0, 1);       // dx / dctx        A K-major, B MN-major
  const uint32_t id_mnmn_d = make_idesc_bf16(128, D, 1, 1);      // weight gradients A MN-major, B MN-major
  const uint32_t id_kk_3d = make_idesc_bf16(128, 3 * D, 0, 0);   // q|k|v recompute

  // everything above touched only this layer's parameters; dy / the saved activations are earlier kernels' output (common.cuh: PDL)
  griddep_wait();
  griddep_launch();
  int iter = 0;
  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++iter) {
    const int64_t grow = (int64_t)tile * TC_TILE + row;
    const bool valid = grow < a.M;
    const uint32_t acc0 = iter > 0 ? 1u : 0u;
    // Every row-wise phase is split four ways: thread (row, part) owns columns [8 part, 8 part + 8) of its token row; row
    // statistics are combined through xsum() (one __syncthreads per exchange).  All 16 warps work in the LayerNorm phases
    // instead of the 4 warps of part 0, and the three column sums of a phase share ONE 31-shuffle transpose.
    const int c0 = part * 8;
    float du[8];                       // gradient w.r.t. the LayerNorm input currently being processed (this thread's 8 columns)
    float mean1 = 0.f, rstd1 = 1.f;    // LayerNorm1 statistics of this row
    float u1k[8];                      // whole layer: this thread's u1 columns, loaded in B0 and kept for B2 (no second round trip)
    // ---- B0: LN2 backward ; x1 = LN1(u1) ; stage x_in (and the memory tile / the FFN input) ----
    {
      float dyv[8], u2v[8], u1v[8];
      auto load8 = [&](const float *src, float (&o)[8]) {
        const float4 t0 = valid ? *reinterpret_cast<const float4 *>(src + grow * D + c0) : make_float4(0, 0, 0, 0);
        const float4 t1 = valid ? *reinterpret_cast<const float4 *>(src + grow * D + c0 + 4) : make_float4(0, 0, 0, 0);
        o[0] = t0.x; o[1] = t0.y; o[2] = t0.z; o[3] = t0.w; o[4] = t1.x; o[5] = t1.y; o[6] = t1.z; o[7] = t1.w;
      };
      load8(a.dy, dyv);
      {
        // the row-wise phases start with plain loads of rows nobody has touched yet: pull the NEXT tile's rows into L2 now,
        // a whole tile time ahead (part p fetches array p: one 128-byte line per row)
        const int64_t nrow = grow + (int64_t)gridDim.x * TC_TILE;
        if (nrow < a.M) {
          const float *pf = part == 0 ? a.dy : part == 1 ? (ATTN_ONLY ? a.u1_in : a.u2_in) : part == 2 ? a.x_in : (MODE == TC_MODE_LAYER ? a.u1_in : (CROSS ? a.mem : a.dy));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + nrow * D));
        }
      }
      if constexpr (MODE != TC_MODE_FFN) {           // x_in tile -> bf16 A operand of the q|k|v recompute
        uint4 v = make_uint4(0, 0, 0, 0);
        if (valid) v = pack8(a.x_in + grow * D + c0);
        *reinterpret_cast<uint4 *>(sXin + kmajor_off(row, c0, 128)) = v;
      }
      if constexpr (MODE == TC_MODE_FFN || CROSS) {   // FFN block: its input IS x1 ; cross-attention: the encoder-memory tile
        const float *src = MODE == TC_MODE_FFN ? a.x_in : a.mem;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (valid) v = pack8(src + grow * D + c0);
        *reinterpret_cast<uint4 *>(sX1 + kmajor_off(row, c0, 128)) = v;
      }
      if constexpr (ATTN_ONLY) {
        // attention block alone: the incoming gradient is already dL/d(LayerNorm1 output)
#pragma unroll
        for (int j = 0; j < 8; ++j) du[j] = dyv[j];
      } else {
        load8(a.u2_in, u2v);
        float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 8; ++j) { e.x += u2v[j]; e.y = fmaf(u2v[j], u2v[j], e.y); }
        if constexpr (MODE == TC_MODE_LAYER) {
          load8(a.u1_in, u1v);
#pragma unroll
          for (int j = 0; j < 8; ++j) { e.z += u1v[j]; e.w = fmaf(u1v[j], u1v[j], e.w); }
        }
        e = xsum(e);
        const float mean2 = e.x * (1.f / D), rstd2 = rsqrtf(fmaxf(e.y * (1.f / D) - mean2 * mean2, 0.f) + LN_EPS);
        if constexpr (MODE == TC_MODE_LAYER) {
          mean1 = e.z * (1.f / D);
          rstd1 = rsqrtf(fmaxf(e.w * (1.f / D) - mean1 * mean1, 0.f) + LN_EPS);
          float y[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) { u1k[j] = u1v[j]; y[j] = (u1v[j] - mean1) * rstd1 * p_g1[c0 + j] + p_be1[c0 + j]; }
          *reinterpret_cast<uint4 *>(sX1 + kmajor_off(row, c0, 128)) =
              make_uint4(pack_bf16(y[0], y[1]), pack_bf16(y[2], y[3]), pack_bf16(y[4], y[5]), pack_bf16(y[6], y[7]));
        }
        float w[32];
        float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          u2v[j] = (u2v[j] - mean2) * rstd2;           // x-hat
          w[j] = dyv[j] * u2v[j];                      // d gamma contribution
          w[8 + j] = dyv[j];                           // d beta contribution
          const float g = dyv[j] * p_g2[c0 + j];
          m.x += g; m.y = fmaf(g, u2v[j], m.y);
        }
        m = xsum(m);
        const float m1 = m.x * (1.f / D), m2 = m.y * (1.f / D);
#pragma unroll
        for (int j = 0; j < 8; ++j) du[j] = rstd2 * (dyv[j] * p_g2[c0 + j] - m1 - u2v[j] * m2);
        const uint64_t e0 = (uint64_t)((a.seq0 * 32 + grow) * D + c0);
        float k0, k1, k2, k3;
        drop4(a.d2, e0, k0, k1, k2, k3);
        w[16] = du[0] * k0; w[17] = du[1] * k1; w[18] = du[2] * k2; w[19] = du[3] * k3;     // da2 = grad wrt the FFN2 output (+bias)
        drop4(a.d2, e0 + 4, k0, k1, k2, k3);
        w[20] = du[4] * k0; w[21] = du[5] * k1; w[22] = du[6] * k2; w[23] = du[7] * k3;
#pragma unroll
        for (int j = 24; j < 32; ++j) w[j] = 0.f;
        // the image carries the hidden dropout's 1 / (1 - p) (forward: f = fs (H_kept W2^T) + b2): both its consumers need it
        // — dH = (fs da2) W2 and dW2 = (fs da2)^T H_kept — while the bias gradient below sums the unscaled da2
        const float fs = a.d_ffn.scale;
        *reinterpret_cast<uint4 *>(sDA + kmajor_off(row, c0, 128)) =
            make_uint4(pack_bf16(w[16] * fs, w[17] * fs), pack_bf16(w[18] * fs, w[19] * fs), pack_bf16(w[20] * fs, w[21] * fs),
                       pack_bf16(w[22] * fs, w[23] * fs));
        ln_bwd_colsums(w, lane, g_g2 + c0, g_be2 + c0, g_b2 + c0);
      }
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    if constexpr (!ATTN_ONLY) {
    // ---- B1: FFN backward, chunk by chunk ----
    const uint32_t w1_chunk = tc_w1_chunk_bytes(D, FC);
    if (tid == 0) {
      fence_after_sync();
#pragma unroll
      for (int k = 0; k < (D + TC_KAUG) / 16; ++k) mma_bf16_ss(t_big, desc_a128(aX1, k), desc_b(aW + io.w1, FC, k), id_kk_fc, k > 0);
      mma_commit(&bar_h);
    }
    // The H recompute of chunk c + 1 is issued BEFORE the dx1 / dW1 / dW2 MMAs of chunk c and signals its own barrier, so the
    // H epilogue of chunk c + 1 runs on the CUDA cores while those MMAs are still in the tensor pipe.  The H image is double
    // buffered for that (odd chunks use the ctx | dqkv image region, which is dead until the attention phase); every other
    // buffer is protected by the in-order completion the dH commit of the next chunk waits for.
    for (int c = 0; c < nchunk; ++c) {
      uint8_t *sHc = (c & 1) ? sCtx : sH;
      const uint32_t aHc = (c & 1) ? aCtx : aH;
      mbar_wait(&bar_h, phh); phh ^= 1;               // H(c) accumulator ready
      fence_after_sync();
      {
        const uint64_t w0 = (uint64_t)((a.seq0 * 32 + grow) * F + c * FC) >> 2;
        for (int b = part; b < nblk; b += BWD_PARTS) {
          const int cb = b * 16;
          const uint64_t wb = w0 + (uint64_t)(cb >> 2);
          ffn_hidden_block(t_big + lane_off + (uint32_t)cb, sHc + kmajor_off(row, cb, 128), a.d_ffn, (uint32_t)wb,
                           (uint32_t)(wb >> 32) * 0x85EBCA6Bu);
        }
      }
      fence_async_smem();
      fence_before_sync();
      __syncthreads();
      if (tid == 0) {                                 // dH(c) = da2 . W2[:, chunk]   (B: MN-major view of the W2 chunk image [D x FC])
        fence_after_sync();
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          mma_bf16_ss(t_big, desc_a128(aDA, k), desc_mn(aW + io.w2 + (uint32_t)c * D * FC * 2u, D, k), id_kmn_fc, k > 0);
        mma_commit(&bar_mma);
      }
      mbar_wait(&bar_mma, ph); ph ^= 1;
      fence_after_sync();
      // dH = dH_acc where the hidden unit was kept and positive, i.e. where this thread's own H pair (still in the image it
      // wrote above) is non-zero: pack, one packed compare and one AND per pair; the 1 / (1 - p) came in with the da2 image
      for (int b = part; b < nblk; b += BWD_PARTS) {
        const int cb = b * 16;
        float w[16];
        tmem_ld16(t_big + lane_off + (uint32_t)cb, w);
        const uint4 h0 = *reinterpret_cast<const uint4 *>(sHc + kmajor_off(row, cb, 128));
        const uint4 h1 = *reinterpret_cast<const uint4 *>(sHc + kmajor_off(row, cb, 128) + 2048);
        tmem_ld_wait();
        *reinterpret_cast<uint4 *>(sDH + kmajor_off(row, cb, 128)) =
            make_uint4(pack_bf16(w[0], w[1]) & nonzero2(h0.x), pack_bf16(w[2], w[3]) & nonzero2(h0.y),
                       pack_bf16(w[4], w[5]) & nonzero2(h0.z), pack_bf16(w[6], w[7]) & nonzero2(h0.w));
        *reinterpret_cast<uint4 *>(sDH + kmajor_off(row, cb, 128) + 2048) =
            make_uint4(pack_bf16(w[8], w[9]) & nonzero2(h1.x), pack_bf16(w[10], w[11]) & nonzero2(h1.y),
                       pack_bf16(w[12], w[13]) & nonzero2(h1.z), pack_bf16(w[14], w[15]) & nonzero2(h1.w));
      }
      fence_async_smem();
      fence_before_sync();
      __syncthreads();
      if (tid == 0) {
        fence_after_sync();
        if (c + 1 < nchunk) {                         // H(c + 1) first, with its own barrier (t_big: the dH(c) accumulator is drained)
#pragma unroll
          for (int k = 0; k < (D + TC_KAUG) / 16; ++k)
            mma_bf16_ss(t_big, desc_a128(aX1, k), desc_b(aW + io.w1 + (uint32_t)(c + 1) * w1_chunk, FC, k), id_kk_fc, k > 0);
          mma_commit(&bar_h);
        }
        const uint32_t w1c = aW + io.w1 + (uint32_t)c * w1_chunk;
        for (int k = 0; k < FC / 16; ++k)             // dx1 += dH . W1[chunk, :]     (B: MN-major view of the W1 chunk image [FC x D])
          mma_bf16_ss(t_sa, desc_a128(aDH, k), desc_mn(w1c, FC, k), id_kmn_d, (c | k) > 0);
#pragma unroll
        for (int k = 0; k < 8; ++k) {                 // contract over the tile's 128 tokens
          mma_bf16_ss(t_dw1 + 32u * c, desc_mn(aDH, 128, k), desc_mn(aX1, 128, k), id_mnmn_d, k > 0 ? 1u : acc0);   // dW1[chunk]   += dH^T x1
          mma_bf16_ss(t_dw2 + 32u * c, desc_mn(aHc, 128, k), desc_mn(aDA, 128, k), id_mnmn_d, k > 0 ? 1u : acc0);   // dW2^T[chunk] += H^T da2
        }
        if (c + 1 == nchunk) mma_commit(&bar_mma);    // earlier chunks: covered by the dH commit of the next chunk
      }
      // linear1 bias gradient: column sums of the dH image — AFTER the issue, so that it overlaps the round trip of H(c + 1) instead
      // of delaying it (the image is rewritten only behind the __syncthreads that follows the next chunk's H epilogue)
      colsum_image(sDH, FC >> 3, g_b1 + c * FC, warp, lane, BWD_THREADS / 32);
    }
    mbar_wait(&bar_mma, ph); ph ^= 1;                 // dx1 complete, all weight-gradient MMAs of this tile retired
    fence_after_sync();
    if constexpr (MODE != 0) {
      // FFN block alone: dx = du2 (residual path) + dx1 (FFN path); nothing else to do for this tile
      {
        float acc[8];
        tmem_ld8(t_sa + lane_off + (uint32_t)c0, acc);
        tmem_ld_wait();
        if (valid) {
          *reinterpret_cast<float4 *>(a.dx + grow * D + c0) = make_float4(du[0] + acc[0], du[1] + acc[1], du[2] + acc[2], du[3] + acc[3]);
          *reinterpret_cast<float4 *>(a.dx + grow * D + c0 + 4) = make_float4(du[4] + acc[4], du[5] + acc[5], du[6] + acc[6], du[7] + acc[7]);
        }
      }
      fence_before_sync();
      __syncthreads();
      continue;
    }
    }  // !ATTN_ONLY
    // ---- B2: LN1 backward (all parts, 8 columns each) ----
    {
      float u1v[8], xh[8];
      if constexpr (MODE == TC_MODE_LAYER) {
#pragma unroll
        for (int j = 0; j < 8; ++j) u1v[j] = u1k[j];
      } else {
        const float4 t0 = valid ? *reinterpret_cast<const float4 *>(a.u1_in + grow * D + c0) : make_float4(0, 0, 0, 0);
        const float4 t1 = valid ? *reinterpret_cast<const float4 *>(a.u1_in + grow * D + c0 + 4) : make_float4(0, 0, 0, 0);
        u1v[0] = t0.x; u1v[1] = t0.y; u1v[2] = t0.z; u1v[3] = t0.w; u1v[4] = t1.x; u1v[5] = t1.y; u1v[6] = t1.z; u1v[7] = t1.w;
      }
      if constexpr (ATTN_ONLY) {                     // no forward recompute ran in B0: LayerNorm1 statistics from u1 here
        float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 8; ++j) { e.x += u1v[j]; e.y = fmaf(u1v[j], u1v[j], e.y); }
        e = xsum(e);
        mean1 = e.x * (1.f / D);
        rstd1 = rsqrtf(fmaxf(e.y * (1.f / D) - mean1 * mean1, 0.f) + LN_EPS);
      } else {
        float acc[8];
        tmem_ld8(t_sa + lane_off + (uint32_t)c0, acc);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) du[j] += acc[j];  // grad wrt x1 = residual path + FFN path
      }
      float w[32];
      float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        xh[j] = (u1v[j] - mean1) * rstd1;
        w[j] = du[j] * xh[j];
        w[8 + j] = du[j];
        const float g = du[j] * p_g1[c0 + j];
        m.x += g; m.y = fmaf(g, xh[j], m.y);
      }
      m = xsum(m);
      const float m1 = m.x * (1.f / D), m2 = m.y * (1.f / D);
#pragma unroll
      for (int j = 0; j < 8; ++j) du[j] = rstd1 * (du[j] * p_g1[c0 + j] - m1 - xh[j] * m2);
      const uint64_t e0 = (uint64_t)((a.seq0 * 32 + grow) * D + c0);
      float k0, k1, k2, k3;
      drop4(a.d1, e0, k0, k1, k2, k3);
      w[16] = du[0] * k0; w[17] = du[1] * k1; w[18] = du[2] * k2; w[19] = du[3] * k3;       // da1 = grad wrt the out-proj output (+bias)
      drop4(a.d1, e0 + 4, k0, k1, k2, k3);
      w[20] = du[4] * k0; w[21] = du[5] * k1; w[22] = du[6] * k2; w[23] = du[7] * k3;
#pragma unroll
      for (int j = 24; j < 32; ++j) w[j] = 0.f;
      *reinterpret_cast<uint4 *>(sDA + kmajor_off(row, c0, 128)) =
          make_uint4(pack_bf16(w[16], w[17]), pack_bf16(w[18], w[19]), pack_bf16(w[20], w[21]), pack_bf16(w[22], w[23]));
      if (valid) {                                    // park du1 (residual path into dx) in the output buffer
        *reinterpret_cast<float4 *>(a.dx + grow * D + c0) = make_float4(du[0], du[1], du[2], du[3]);
        *reinterpret_cast<float4 *>(a.dx + grow * D + c0 + 4) = make_float4(du[4], du[5], du[6], du[7]);
      }
      ln_bwd_colsums(w, lane, g_g1 + c0, g_be1 + c0, g_bo + c0);
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    // ---- B3: recompute q|k|v ; dctx = da1 . Wo ----
    if (tid == 0) {
      fence_after_sync();
      if constexpr (CROSS) {                          // q = x Wq^T ; k | v = mem Wkv^T
        const uint32_t idq = make_idesc_bf16(128, D), idkv = make_idesc_bf16(128, 2 * D);
#pragma unroll
        for (int k = 0; k < D / 16; ++k) mma_bf16_ss(t_big, desc_a128(aXin, k), desc_b_rows(aW + io.wqkv, 3 * D, 0, k), idq, k > 0);
#pragma unroll
        for (int k = 0; k < D / 16; ++k) mma_bf16_ss(t_big + (uint32_t)D, desc_a128(aX1, k), desc_b_rows(aW + io.wqkv, 3 * D, D, k), idkv, k > 0);
      } else {
#pragma unroll
        for (int k = 0; k < D / 16; ++k) mma_bf16_ss(t_big, desc_a128(aXin, k), desc_b(aW + io.wqkv, 3 * D, k), id_kk_3d, k > 0);
      }
#pragma unroll
      for (int k = 0; k < D / 16; ++k) mma_bf16_ss(t_sb, desc_a128(aDA, k), desc_mn(aW + io.wo, D, k), id_kmn_d, k > 0);
      mma_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, ph); ph ^= 1;
    fence_after_sync();
    {                                                 // part 0 -> q, 1 -> k, 2 -> v, 3 -> dctx
      float v[D];
      const uint32_t src = part < 3 ? t_big + (uint32_t)(part * D) : t_sb;
      tmem_ld16(src + lane_off, v);
      tmem_ld16(src + lane_off + 16u, v + 16);
      tmem_ld_wait();
      if (part < 3) {
#pragma unroll
        for (int c = 0; c < D; c += 4) {
          const float4 b = *reinterpret_cast<const float4 *>(p_bqkv + part * D + c);
          v[c] += b.x; v[c + 1] += b.y; v[c + 2] += b.z; v[c + 3] += b.w;
        }
      }
      if constexpr (MMA) {
        a32_store_row(smem + (part == 0 ? sp.q : part == 1 ? sp.k : part == 2 ? sp.v : sp.dctx), row, v, part == 0 ? attn_scale : 1.f);
      } else if (part == 0 || part == 3) {
        float *dst = (part == 0 ? sQ : sDC) + row * LQ;
#pragma unroll
        for (int c = 0; c < D; c += 4) *reinterpret_cast<float4 *>(dst + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
      } else {
        store_kv_row<D, DH>(part == 1 ? sK : sV, v, row, H, dh);
      }
    }
    fence_before_sync();
    __syncthreads();
    // ---- B5: attention backward, one warp per (sequence, head) ----
    if constexpr (MMA) {
      for (int p = warp; p < 4 * H; p += BWD_THREADS / 32) {
        const int s = p / H, h = p - s * H;
        const uint64_t w_pair = (uint64_t)((((a.seq0 + (int64_t)tile * 4 + s) * H + h) * 32) * 8);
        a32_attn_bwd<DH, CAUSAL>(smem + sp.q, smem + sp.k, smem + sp.v, smem + sp.dctx, sCtx, sDQ, s, h, lane, a.d_attn, w_pair);
      }
    } else
    for (int p = warp; p < 4 * H; p += BWD_THREADS / 32) {      // lane = query row; dK/dV via warp transposed sums
      const int s = p / H, h = p - s * H;
      const int r = s * 32 + lane;
      const float *q = sQ + r * LQ + h * dh;
      const float *go = sDC + r * LQ + h * dh;
      const float *kh = sK + s * 32 * D + h * 32 * dh;
      const float *vh = sV + s * 32 * D + h * 32 * dh;
      float pr[32], dp[32];
      attn_scores<DH>(q, kh, dh, attn_scale, pr);
      attn_scores<DH>(go, vh, dh, 1.f, dp);           // dP[i][j] = dO_i . v_j
      float mx = pr[0];
#pragma unroll
      for (int j = 1; j < 32; ++j) mx = fmaxf(mx, pr[j]);
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) { pr[j] = ex2_ftz(pr[j] - mx); sum += pr[j]; }
      const float inv = 1.f / sum;
      uint32_t keep = 0xFFFFFFFFu;
      if (a.d_attn.thr) {
        keep = 0;
        const uint64_t w0 = (uint64_t)((((a.seq0 + (int64_t)tile * 4 + s) * H + h) * 32 + lane) * 8);
        const uint32_t xhi = (uint32_t)(w0 >> 32) * 0x85EBCA6Bu, wlo = (uint32_t)w0;
#pragma unroll
        for (int m = 0; m < 8; ++m) {                 // quad m = keys k0, k0+1, k0+8, k0+9 (common.cuh: key_perm)
          const int k0 = (m >> 2) * 16 + (m & 3) * 2;
          uint32_t lo, hi;
          hash_quad((wlo + (uint32_t)m) ^ xhi, a.d_attn.key, lo, hi);
          keep |= ((lo & 0xFFFFu) >= a.d_attn.thr ? 1u : 0u) << k0;
          keep |= ((lo >> 16) >= a.d_attn.thr ? 1u : 0u) << (k0 + 1);
          keep |= ((hi & 0xFFFFu) >= a.d_attn.thr ? 1u : 0u) << (k0 + 8);
          keep |= ((hi >> 16) >= a.d_attn.thr ? 1u : 0u) << (k0 + 9);
        }
      }
      const float ks = a.d_attn.scale;
      float delta = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        pr[j] *= inv;                                                  // P
        dp[j] = ((keep >> j) & 1u) ? dp[j] * ks : 0.f;                 // dL/dP
        delta = fmaf(dp[j], pr[j], delta);
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) dp[j] = pr[j] * (dp[j] - delta) * inv_sqrt_dh;      // dS (1/sqrt(dh) folded in)
      for (int c = 0; c < dh; ++c) {
        const float qc = q[c], gc = go[c];
        float o = 0.f, dq = 0.f;
        float w[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float pd = ((keep >> j) & 1u) ? pr[j] * ks : 0.f;      // dropped probabilities
          o = fmaf(pd, vh[j * dh + c], o);
          dq = fmaf(dp[j], kh[j * dh + c], dq);
          w[j] = pd * gc;
        }
        const float dv = warp_colsum32(w, lane);      // lane = key row
#pragma unroll
        for (int j = 0; j < 32; ++j) w[j] = dp[j] * qc;
        const float dk = warp_colsum32(w, lane);
        *reinterpret_cast<__nv_bfloat16 *>(sCtx + kmajor_off(r, h * dh + c, 128)) = __float2bfloat16_rn(o);
        *reinterpret_cast<__nv_bfloat16 *>(sDQ + kmajor_off(r, h * dh + c, 128)) = __float2bfloat16_rn(dq);
        *reinterpret_cast<__nv_bfloat16 *>(sDQ + kmajor_off(r, D + h * dh + c, 128)) = __float2bfloat16_rn(dk);
        *reinterpret_cast<__nv_bfloat16 *>(sDQ + kmajor_off(r, 2 * D + h * dh + c, 128)) = __float2bfloat16_rn(dv);
        const float sq = warp_sum_all(dq), sk = warp_sum_all(dk), sv = warp_sum_all(dv);
        if (lane == 0) {
          atomicAdd(&g_bqkv[h * dh + c], sq);
          atomicAdd(&g_bqkv[D + h * dh + c], sk);
          atomicAdd(&g_bqkv[2 * D + h * dh + c], sv);
        }
      }
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    if constexpr (MMA) colsum_image(sDQ, 3 * D / 8, g_bqkv, warp, lane, BWD_THREADS / 32);      // in-projection bias gradient: column sums of the dq | dk | dv image
    // ---- B6: dx_in = dqkv . Wqkv ; dWqkv += dqkv^T x_in ; dWo += da1^T ctx ----
    if (tid == 0) {
      fence_after_sync();
      if constexpr (CROSS) {
        // dx = dq Wq (dqkv columns [0, D)) ; dmem = dk|dv Wkv (columns [D, 3D)) -> t_sb (dctx is consumed)
#pragma unroll
        for (int k = 0; k < D / 16; ++k) mma_bf16_ss(t_sa, desc_a128(aDQ, k), desc_mn(aW + io.wqkv, 3 * D, k), id_kmn_d, k > 0);
#pragma unroll
        for (int k = D / 16; k < 3 * D / 16; ++k) mma_bf16_ss(t_sb, desc_a128(aDQ, k), desc_mn(aW + io.wqkv, 3 * D, k), id_kmn_d, k > D / 16);
      } else {
#pragma unroll
        for (int k = 0; k < 3 * D / 16; ++k) mma_bf16_ss(t_sa, desc_a128(aDQ, k), desc_mn(aW + io.wqkv, 3 * D, k), id_kmn_d, k > 0);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        mma_bf16_ss(t_dwqkv, desc_mn(aDQ, 128, k), desc_mn(aXin, 128, k), id_mnmn_d, k > 0 ? 1u : acc0);
        // cross: rows [D, 3D) of dWqkv contract dk | dv with the MEMORY tile: a second accumulator (the unused dW1 columns)
        if constexpr (CROSS) mma_bf16_ss(t_dw1, desc_mn(aDQ, 128, k), desc_mn(aX1, 128, k), id_mnmn_d, k > 0 ? 1u : acc0);
        mma_bf16_ss(t_dwo, desc_mn(aDA, 128, k), desc_mn(aCtx, 128, k), id_mnmn_d, k > 0 ? 1u : acc0);
      }
      mma_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, ph); ph ^= 1;
    fence_after_sync();
    if constexpr (CROSS) {
      if (part == 1) {                                // dmem += dk|dv Wkv (this CTA owns the tile's rows; decoder layers run in stream order)
        float acc[D];
#pragma unroll
        for (int cb = 0; cb < D; cb += 16) tmem_ld16(t_sb + lane_off + (uint32_t)cb, acc + cb);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int c = 0; c < D; c += 4) {
            float4 t = *reinterpret_cast<const float4 *>(a.dmem + grow * D + c);
            *reinterpret_cast<float4 *>(a.dmem + grow * D + c) = make_float4(t.x + acc[c], t.y + acc[c + 1], t.z + acc[c + 2], t.w + acc[c + 3]);
          }
        }
      }
    }
    {                                                 // dx = du1 (parked by this same thread) + dx_in
      float acc[8];
      tmem_ld8(t_sa + lane_off + (uint32_t)c0, acc);
      tmem_ld_wait();
      if (valid) {
        const float4 t0 = *reinterpret_cast<const float4 *>(a.dx + grow * D + c0), t1 = *reinterpret_cast<const float4 *>(a.dx + grow * D + c0 + 4);
        *reinterpret_cast<float4 *>(a.dx + grow * D + c0) = make_float4(t0.x + acc[0], t0.y + acc[1], t0.z + acc[2], t0.w + acc[3]);
        *reinterpret_cast<float4 *>(a.dx + grow * D + c0 + 4) = make_float4(t1.x + acc[4], t1.y + acc[5], t1.z + acc[6], t1.w + acc[7]);
      }
    }
    fence_before_sync();
    __syncthreads();
  }
  // ---- flush: TMEM-resident weight gradients and shared-memory bias / LayerNorm partials -> global (fp32 atomics) ----
  fence_after_sync();
  {
    const int m = row;                                // TMEM lane = output row of the weight-gradient blocks
    // jobs: 0..nchunk-1 dW1 chunks, nchunk..2nchunk-1 dW2^T chunks, then dWqkv, dWo — dealt round-robin to the parts
    for (int job = ATTN_ONLY ? 2 * nchunk + part : part; job < 2 * nchunk + (MODE == TC_MODE_FFN ? 0 : 2); job += BWD_PARTS) {
      float v[D];
      uint32_t t;
      if (job < nchunk) t = t_dw1 + 32u * job;
      else if (job < 2 * nchunk) t = t_dw2 + 32u * (job - nchunk);
      else t = job == 2 * nchunk ? ((CROSS && m >= D) ? t_dw1 : t_dwqkv) : t_dwo;
#pragma unroll
      for (int cb = 0; cb < D; cb += 16) tmem_ld16(t + lane_off + (uint32_t)cb, v + cb);
      tmem_ld_wait();
      if (job >= nchunk && job < 2 * nchunk) {          // dW2^T: lane = hidden unit = the contiguous index of gw2 — coalesced as it is
        if (m < FC) {
#pragma unroll
          for (int j = 0; j < D; ++j) atomicAdd(a.gw2 + (int64_t)j * F + (job - nchunk) * FC + m, v[j]);
        }
      } else {
        // dW1 / dWqkv / dWo: lane = output ROW, the 32 values of a thread are contiguous in memory — as scalar atomics every warp
        // instruction touched 32 different 128-byte lines.  The warp's 32 x 32 block is turned through (dead) shared memory so
        // that one instruction adds one whole row: 32 x fewer L2 requests (the flush is 160 KB per CTA onto the same 40 K words).
        float *gdst;
        int rows;
        if (job < nchunk) { gdst = a.gw1 + (int64_t)job * FC * D; rows = FC; }
        else if (job == 2 * nchunk) { gdst = a.gwqkv; rows = 3 * D; }
        else { gdst = a.gwo; rows = D; }
        float *stg = reinterpret_cast<float *>(smem + sp.xin) + warp * (32 * 33);
#pragma unroll
        for (int i = 0; i < D; ++i) stg[lane * 33 + i] = v[i];
        __syncwarp();
        const int m0 = (warp & 3) * 32;
#pragma unroll 8
        for (int r = 0; r < 32; ++r)
          if (m0 + r < rows) atomicAdd(gdst + (int64_t)(m0 + r) * D + lane, stg[r * 33 + lane]);
        __syncwarp();
      }
    }
  }
  if constexpr (MODE != TC_MODE_FFN)
    for (int i = tid; i < 3 * D; i += BWD_THREADS) atomicAdd(a.gbqkv + i, g_bqkv[i]);
  if constexpr (!ATTN_ONLY)
    for (int i = tid; i < F; i += BWD_THREADS) atomicAdd(a.gb1 + i, g_b1[i]);
  if (tid < D) {
    if constexpr (MODE != TC_MODE_FFN) {
      atomicAdd(a.gbo + tid, g_bo[tid]);
      atomicAdd(a.gg1 + tid, g_g1[tid]); atomicAdd(a.gbe1 + tid, g_be1[tid]);
    }
    if constexpr (!ATTN_ONLY) {
      atomicAdd(a.gb2 + tid, g_b2[tid]);
      atomicAdd(a.gg2 + tid, g_g2[tid]); atomicAdd(a.gbe2 + tid, g_be2[tid]);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int DH, int MODE = 0>
static int launch_bwd(const TcLayerArgs &a, uint32_t smem, int grid, cudaStream_t st) {
  if (args_devstep(a)) {
    GT_CUDA(cudaFuncSetAttribute(tc_layer_bwd_kernel<32, DH, MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    { LaunchScope _ls(KC_TC_LAYER_BWD, st);
      GT_CUDA(launch_pdl(tc_layer_bwd_kernel<32, DH, MODE, true>, dim3(grid), dim3(BWD_THREADS), smem, st, a)); }
    GT_CUDA(cudaGetLastError());
    return 0;
  }
  GT_CUDA(cudaFuncSetAttribute(tc_layer_bwd_kernel<32, DH, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  { LaunchScope _ls(KC_TC_LAYER_BWD, st);
    GT_CUDA(launch_pdl(tc_layer_bwd_kernel<32, DH, MODE>, dim3(grid), dim3(BWD_THREADS), smem, st, a)); }
  GT_CUDA(cudaGetLastError());
  return 0;
}

int tc_layer_bwd(int D, const TcLayerArgs &a, cudaStream_t st) {
  GT_CHECK(D == 32, "tc_layer_bwd: d_model not instantiated");
  GT_CHECK(a.F / a.FC <= 4, "tc_layer_bwd: more than 4 FFN chunks do not fit the TMEM gradient accumulators");
  const BwdSmem sp = bwd_smem(D, a.F, attn_mma(a.dh));
  GT_CHECK(sp.total <= 227 * 1024, "tc_layer_bwd: shared memory budget exceeded");
  int grid = a.n_tiles < num_sms() ? a.n_tiles : num_sms();
  if (a.mode == TC_MODE_FFN) return launch_bwd<2, TC_MODE_FFN>(a, bwd_smem(D, a.F, true).total, grid, st);
  if (a.mode == TC_MODE_ATTN_CAUSAL || a.mode == TC_MODE_ATTN_CROSS) {
    const bool causal = a.mode == TC_MODE_ATTN_CAUSAL;
    switch (a.dh) {
      case 2: return causal ? launch_bwd<2, TC_MODE_ATTN_CAUSAL>(a, sp.total, grid, st) : launch_bwd<2, TC_MODE_ATTN_CROSS>(a, sp.total, grid, st);
      case 4: return causal ? launch_bwd<4, TC_MODE_ATTN_CAUSAL>(a, sp.total, grid, st) : launch_bwd<4, TC_MODE_ATTN_CROSS>(a, sp.total, grid, st);
      case 8: return causal ? launch_bwd<8, TC_MODE_ATTN_CAUSAL>(a, sp.total, grid, st) : launch_bwd<8, TC_MODE_ATTN_CROSS>(a, sp.total, grid, st);
      default: GT_FAIL("tc_layer_bwd: attention-only blocks need head dim 2, 4 or 8");
    }
  }
  switch (a.dh) {
    case 2: return launch_bwd<2>(a, sp.total, grid, st);
    case 4: return launch_bwd<4>(a, sp.total, grid, st);
    case 8: return launch_bwd<8>(a, sp.total, grid, st);
    default: return launch_bwd<0>(a, sp.total, grid, st);
  }
}

}  // namespace gt
