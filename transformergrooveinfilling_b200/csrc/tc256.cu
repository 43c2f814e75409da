// tc256.cu — fused encoder-layer kernels for d_model = 256.  See tc256.cuh for the design.
#include "tc256.cuh"
#include "umma.cuh"

namespace gt {
using namespace umma;

bool t256_shape_supported(const gt_config &c, std::string *why) {
  auto no = [&](const char *m) { if (why) *why = m; return false; };
  if (c.n_dec != 0) return no("encoder-decoder models run in precision=fp32 (the fused tcgen05 layer kernels cover the encoder stack)");
  if (c.d_model != 256) return no("t256 kernels are instantiated for d_model=256");
  if (c.n_enc > TC_MAX_LAYERS) return no("more than 16 layers");
  const int dh = c.d_model / c.nhead;
  if (dh != 16 && dh != 32) return no("d_model=256 tensor-core kernels need head_dim 16 or 32 (nhead 16 or 8)");
  if (c.dim_ff % 64 != 0 || c.dim_ff < 64 || c.dim_ff > 512) return no("d_model=256 tensor-core kernels need dim_feedforward in {64,128,...,512}");
  return true;
}

__device__ __forceinline__ uint4 t256_pack8(const float *s) {
  float4 a = *reinterpret_cast<const float4 *>(s), b = *reinterpret_cast<const float4 *>(s + 4);
  return make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
}

// =============================================================================================
// weight prep: fp32 master parameters -> the forward stage stream (bf16 canonical K-major images)
// grid = (stage blocks, layers)
// =============================================================================================
__global__ void t256_prep_kernel(TcPrepArgs a) {
  const int l = blockIdx.y;
  const int F = a.F;
  uint8_t *img = a.img + (size_t)l * a.img_stride;
  const int nst = t256_fwd_stages(F);
  for (int st = blockIdx.x; st < nst; st += gridDim.x) {
    const T256Stage s = t256_fwd_stage(st);
    uint8_t *dst = img + (size_t)st * T256_STAGE;
    const int kbn = s.K / 8, total = s.N * kbn;
    for (int id = threadIdx.x; id < total; id += blockDim.x) {
      const int n = id / kbn, kb = id % kbn;
      const float *src;
      if (s.type == 0) {             // Wqkv [768, 256]: chunk row n -> (q|k|v part, 32 feature columns of group a)
        const int part = n >> 5, within = n & 31;
        src = a.params + a.w_in[l] + (int64_t)(part * 256 + s.a * 32 + within) * 256 + s.b * 64 + kb * 8;
      } else if (s.type == 1) {      // Wo [256, 256]: K = ctx features of group a
        src = a.params + a.w_out[l] + (int64_t)n * 256 + s.a * 32 + kb * 8;
      } else if (s.type == 2) {      // W1 [F, 256]
        src = a.params + a.w1[l] + (int64_t)(s.a * 64 + n) * 256 + s.b * 128 + kb * 8;
      } else {                       // W2 [256, F]
        src = a.params + a.w2[l] + (int64_t)n * F + s.a * 64 + s.b * 32 + kb * 8;
      }
      *reinterpret_cast<uint4 *>(dst + kmajor_off(n, kb * 8, s.N)) = t256_pack8(src);
    }
  }
}

int t256_prep_weights(const TcPrepArgs &a, cudaStream_t st) {
  dim3 grid(32, a.n_layers);
  { LaunchScope _ls(KC_TC_PREP, st);
    t256_prep_kernel<<<grid, 256, 0, st>>>(a); }
  GT_CUDA(cudaGetLastError());
  return 0;
}

// =============================================================================================
// device helpers
// =============================================================================================
__device__ __forceinline__ uint64_t dA(uint32_t base, int k16) { return make_desc(base + (uint32_t)k16 * 4096u, 2048u, 128u); }   // A image, 128 rows
__device__ __forceinline__ uint64_t dB(uint32_t base, int N, int k16) {
  const uint32_t kstride = (uint32_t)(N >> 3) * 128u;
  return make_desc(base + (uint32_t)k16 * 2u * kstride, kstride, 128u);
}
// dropout multipliers of two consecutive elements (idx even): w = idx >> 1 given as (wlo + j, xhi)
__device__ __forceinline__ uint32_t drop_hash(uint32_t wlo_j, uint32_t xhi, uint32_t key) { return mix32(((wlo_j ^ xhi) * 0x9E3779B1u) ^ key); }

struct T256FwdSmem {
  static constexpr uint32_t x = 0, ring = 65536, qkv = 131072, ctx = 155648, par = 172032, stat = 183296, total = 187392;
};

// ---- attention for one half pair: 16 query rows of (sequence s, head hl of the group) against the 32 keys -------
// sQKV: canonical K-major image [128 rows x 96 cols] = [q (32) | k (32) | v (32)], q already scaled by log2(e)/sqrt(dh)
template <int DH>
__device__ __forceinline__ void t256_attn_fwd(const uint8_t *sQKV, uint8_t *sCtxBuf, int s, int hl, int half, int lane, const Drop &dr,
                                              uint64_t w_pair /* ((seq*H + h)*32)*16 : idx>>1 of (query 0, key 0) */) {
  const int g = lane >> 2, t = lane & 3;
  const int r0 = s * 32 + half * 16 + g;            // query rows r0 and r0 + 8
  const int qc = hl * DH, kc = 32 + hl * DH, vc = 64 + hl * DH;
  float sacc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { sacc[i][0] = 0.f; sacc[i][1] = 0.f; sacc[i][2] = 0.f; sacc[i][3] = 0.f; }
#pragma unroll
  for (int kt = 0; kt < DH / 16; ++kt) {
    const uint32_t a0 = lds32(sQKV + kmajor_off(r0, qc + 16 * kt + 2 * t, 128));
    const uint32_t a1 = lds32(sQKV + kmajor_off(r0 + 8, qc + 16 * kt + 2 * t, 128));
    const uint32_t a2 = lds32(sQKV + kmajor_off(r0, qc + 16 * kt + 8 + 2 * t, 128));
    const uint32_t a3 = lds32(sQKV + kmajor_off(r0 + 8, qc + 16 * kt + 8 + 2 * t, 128));
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int key = s * 32 + 8 * nt + g;
      const uint32_t b0 = lds32(sQKV + kmajor_off(key, kc + 16 * kt + 2 * t, 128));
      const uint32_t b1 = lds32(sQKV + kmajor_off(key, kc + 16 * kt + 8 + 2 * t, 128));
      mma16816(sacc[nt], a0, a1, a2, a3, b0, b1);
    }
  }
  // softmax over the 32 keys of rows r0 (c0, c1) and r0 + 8 (c2, c3)
  float m0 = sacc[0][0], m1 = sacc[0][2];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    m0 = fmaxf(m0, fmaxf(sacc[nt][0], sacc[nt][1]));
    m1 = fmaxf(m1, fmaxf(sacc[nt][2], sacc[nt][3]));
  }
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    sacc[nt][0] = exp2f(sacc[nt][0] - m0); sacc[nt][1] = exp2f(sacc[nt][1] - m0);
    sacc[nt][2] = exp2f(sacc[nt][2] - m1); sacc[nt][3] = exp2f(sacc[nt][3] - m1);
    s0 += sacc[nt][0] + sacc[nt][1]; s1 += sacc[nt][2] + sacc[nt][3];
  }
  s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
  s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
  const float i0 = dr.scale / s0, i1 = dr.scale / s1;
  if (dr.thr) {
    const int q0 = half * 16 + g;
    const uint64_t wa = w_pair + (uint64_t)q0 * 16u, wb = wa + 128u;           // rows q0 and q0 + 8
    const uint32_t alo = (uint32_t)wa, ahi = (uint32_t)(wa >> 32) * 0x85EBCA6Bu;
    const uint32_t blo = (uint32_t)wb, bhi = (uint32_t)(wb >> 32) * 0x85EBCA6Bu;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const uint32_t ha = drop_hash(alo + (uint32_t)(4 * nt + t), ahi, dr.key);
      const uint32_t hb = drop_hash(blo + (uint32_t)(4 * nt + t), bhi, dr.key);
      sacc[nt][0] = ((ha & 0xFFFFu) >= dr.thr) ? sacc[nt][0] * i0 : 0.f;
      sacc[nt][1] = ((ha >> 16) >= dr.thr) ? sacc[nt][1] * i0 : 0.f;
      sacc[nt][2] = ((hb & 0xFFFFu) >= dr.thr) ? sacc[nt][2] * i1 : 0.f;
      sacc[nt][3] = ((hb >> 16) >= dr.thr) ? sacc[nt][3] * i1 : 0.f;
    }
  } else {
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) { sacc[nt][0] *= i0; sacc[nt][1] *= i0; sacc[nt][2] *= i1; sacc[nt][3] *= i1; }
  }
  // O = P V
  float oacc[DH / 8][4];
#pragma unroll
  for (int i = 0; i < DH / 8; ++i) { oacc[i][0] = 0.f; oacc[i][1] = 0.f; oacc[i][2] = 0.f; oacc[i][3] = 0.f; }
#pragma unroll
  for (int kt = 0; kt < 2; ++kt) {                  // keys 16 kt .. 16 kt + 15
    const uint32_t p0 = pack_bf16(sacc[2 * kt][0], sacc[2 * kt][1]), p1 = pack_bf16(sacc[2 * kt][2], sacc[2 * kt][3]);
    const uint32_t p2 = pack_bf16(sacc[2 * kt + 1][0], sacc[2 * kt + 1][1]), p3 = pack_bf16(sacc[2 * kt + 1][2], sacc[2 * kt + 1][3]);
#pragma unroll
    for (int np = 0; np < DH / 16; ++np) {          // feature columns 16 np .. 16 np + 15
      const int mi = lane >> 3, rr = lane & 7;
      const int key = s * 32 + 16 * kt + (mi & 1) * 8 + rr;
      uint32_t b[4];
      ldmatrix_x4_trans(b, sQKV + kmajor_off(key, vc + 16 * np + (mi >> 1) * 8, 128));
      mma16816(oacc[2 * np], p0, p1, p2, p3, b[0], b[1]);
      mma16816(oacc[2 * np + 1], p0, p1, p2, p3, b[2], b[3]);
    }
  }
#pragma unroll
  for (int nt = 0; nt < DH / 8; ++nt) {
    *reinterpret_cast<uint32_t *>(sCtxBuf + kmajor_off(r0, hl * DH + 8 * nt + 2 * t, 128)) = pack_bf16(oacc[nt][0], oacc[nt][1]);
    *reinterpret_cast<uint32_t *>(sCtxBuf + kmajor_off(r0 + 8, hl * DH + 8 * nt + 2 * t, 128)) = pack_bf16(oacc[nt][2], oacc[nt][3]);
  }
}

// =============================================================================================
// forward
// =============================================================================================
template <int DH>
__global__ void __launch_bounds__(T256_THREADS, 1) t256_layer_fwd_kernel(const TcLayerArgs a) {
  constexpr int D = 256, G = T256_G, GH = 32 / DH, NS = T256_NS;
  using S = T256FwdSmem;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_full[NS], bar_empty[NS], bar_xready, bar_qkvfull, bar_qkvfree, bar_ctxready[2], bar_ctxfree[2],
      bar_outfull, bar_x1ready, bar_hfull, bar_hready, bar_out2full;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int F = a.F, NCH = F / 64, H = a.H;
  uint8_t *sX = smem + S::x, *sRing = smem + S::ring, *sQKV = smem + S::qkv, *sH = smem + S::qkv, *sCtx = smem + S::ctx;
  float *sPar = reinterpret_cast<float *>(smem + S::par);
  float *p_bqkv = sPar, *p_bo = sPar + 768, *p_b2 = p_bo + 256, *p_g1 = p_b2 + 256, *p_be1 = p_g1 + 256, *p_g2 = p_be1 + 256,
        *p_be2 = p_g2 + 256, *p_b1 = p_be2 + 256;
  float *sStatA = reinterpret_cast<float *>(smem + S::stat), *sStatB = sStatA + 512;

  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 1); }
    mbar_init(&bar_xready, 1); mbar_init(&bar_qkvfull, 1); mbar_init(&bar_qkvfree, 1);
    mbar_init(&bar_ctxready[0], 1); mbar_init(&bar_ctxready[1], 1); mbar_init(&bar_ctxfree[0], 1); mbar_init(&bar_ctxfree[1], 1);
    mbar_init(&bar_outfull, 1); mbar_init(&bar_x1ready, 1); mbar_init(&bar_hfull, 1); mbar_init(&bar_hready, 1); mbar_init(&bar_out2full, 1);
    fence_mbar_init();
  }
  for (int i = tid; i < 768; i += T256_THREADS) p_bqkv[i] = a.bqkv[i];
  for (int i = tid; i < F; i += T256_THREADS) p_b1[i] = a.b1[i];
  if (tid < 256) {
    p_bo[tid] = a.bo[tid]; p_b2[tid] = a.b2[tid]; p_g1[tid] = a.g1[tid]; p_be1[tid] = a.be1[tid]; p_g2[tid] = a.g2[tid]; p_be2[tid] = a.be2[tid];
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const uint32_t t_out = tmem, t_qkv = tmem + 256, t_h = tmem + 384;
  const uint32_t aX = smem_u32(sX), aRing = smem_u32(sRing), aH = smem_u32(sH), aCtx = smem_u32(sCtx);

  if (warp == 16) {
    // ======================= TMA producer: walk the stage stream of every tile =======================
    if (lane == 0) {
      const int nst = t256_fwd_stages(F);
      uint32_t cnt = 0;
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        for (int st = 0; st < nst; ++st, ++cnt) {
          const uint32_t slot = cnt % NS, use = cnt / NS;
          mbar_wait(&bar_empty[slot], (use & 1u) ^ 1u);
          const uint32_t bytes = t256_fwd_stage(st).bytes;
          mbar_expect_tx(&bar_full[slot], bytes);
          tma_load_1d(sRing + slot * T256_STAGE, a.img + (size_t)st * T256_STAGE, bytes, &bar_full[slot]);
        }
      }
    }
  } else if (warp == 17) {
    // ======================= MMA issuer =======================
    if (lane == 0) {
      const uint32_t id_qkv = make_idesc_bf16(128, 96), id_256 = make_idesc_bf16(128, 256), id_64 = make_idesc_bf16(128, 64);
      uint32_t cnt = 0, slot = 0;
      auto take = [&]() -> uint32_t {
        slot = cnt % NS;
        mbar_wait(&bar_full[slot], (cnt / NS) & 1u);
        fence_after_sync();
        return aRing + slot * T256_STAGE;
      };
      auto release = [&]() { mma_commit(&bar_empty[slot]); ++cnt; };
      int it = 0;
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
        auto wo = [&](int gg) {                      // out += ctx[:, group gg] Wo[:, group gg]^T
          const int b = gg & 1;
          const uint32_t nb = (uint32_t)(it * (G / 2) + (gg >> 1));
          mbar_wait(&bar_ctxready[b], nb & 1u);
          fence_after_sync();
          const uint32_t base = take();
#pragma unroll
          for (int k = 0; k < 2; ++k) mma_bf16_ss(t_out, dA(aCtx + (uint32_t)b * 8192u, k), dB(base, 256, k), id_256, (gg | k) > 0);
          release();
          mma_commit(&bar_ctxfree[b]);
        };
        mbar_wait(&bar_xready, (uint32_t)it & 1u);
        fence_after_sync();
        for (int g = 0; g < G; ++g) {
          const uint32_t nq = (uint32_t)(it * G + g);
          mbar_wait(&bar_qkvfree, (nq & 1u) ^ 1u);
          fence_after_sync();
          for (int kc = 0; kc < 4; ++kc) {
            const uint32_t base = take();
#pragma unroll
            for (int k = 0; k < 4; ++k) mma_bf16_ss(t_qkv, dA(aX, kc * 4 + k), dB(base, 96, k), id_qkv, (kc | k) > 0);
            release();
          }
          mma_commit(&bar_qkvfull);
          if (g >= 1) wo(g - 1);
        }
        wo(G - 1);
        mma_commit(&bar_outfull);
        // ---- FFN ----
        mbar_wait(&bar_x1ready, (uint32_t)it & 1u);
        fence_after_sync();
        auto ffn1 = [&]() {
          for (int hf = 0; hf < 2; ++hf) {
            const uint32_t base = take();
#pragma unroll
            for (int k = 0; k < 8; ++k) mma_bf16_ss(t_h, dA(aX, hf * 8 + k), dB(base, 64, k), id_64, (hf | k) > 0);
            release();
          }
          mma_commit(&bar_hfull);
        };
        ffn1();
        for (int c = 0; c < NCH; ++c) {
          const uint32_t nh = (uint32_t)(it * NCH + c);
          mbar_wait(&bar_hready, nh & 1u);
          fence_after_sync();
          for (int hf = 0; hf < 2; ++hf) {
            const uint32_t base = take();
#pragma unroll
            for (int k = 0; k < 2; ++k) mma_bf16_ss(t_out, dA(aH, hf * 2 + k), dB(base, 256, k), id_256, (c | hf | k) > 0);
            release();
          }
          if (c + 1 < NCH) ffn1();
        }
        mma_commit(&bar_out2full);
      }
    }
  } else {
    // ======================= compute warps =======================
    const int q4 = warp & 3, part = warp >> 2;
    const int row = q4 * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q4 * 32) << 16;
    const float attn_scale = rsqrtf((float)DH) * 1.4426950408889634f;
    int it = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
      const int64_t grow = (int64_t)tile * TC_TILE + row;
      const bool valid = grow < a.M;
      // ---- P0: x_in tile -> bf16 A image ----
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int idx = i * T256_CTHREADS + tid;
        const int r = ((idx >> 5) & 15) * 8 + (idx & 7), kb = (idx >> 9) * 4 + ((idx >> 3) & 3);
        const int64_t gr = (int64_t)tile * TC_TILE + r;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (gr < a.M) v = t256_pack8(a.x_in + gr * D + kb * 8);
        *reinterpret_cast<uint4 *>(sX + kmajor_off(r, kb * 8, 128)) = v;
      }
      fence_async_smem();
      named_bar_sync(1, T256_CTHREADS);
      if (tid == 0) mbar_arrive(&bar_xready);
      // ---- P1: head groups ----
      for (int g = 0; g < G; ++g) {
        const uint32_t nq = (uint32_t)(it * G + g);
        mbar_wait(&bar_qkvfull, nq & 1u);
        fence_after_sync();
        {
          float v[24];
#pragma unroll
          for (int i = 0; i < 3; ++i) tmem_ld8(t_qkv + lane_off + (uint32_t)(part * 24 + i * 8), v + i * 8);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const int n0 = part * 24 + i * 8, pq = n0 >> 5;
            const float *bias = p_bqkv + pq * 256 + g * 32 + (n0 & 31);
            const float sc = pq == 0 ? attn_scale : 1.f;
            float w[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) w[j] = (v[i * 8 + j] + bias[j]) * sc;
            *reinterpret_cast<uint4 *>(sQKV + kmajor_off(row, n0, 128)) =
                make_uint4(pack_bf16(w[0], w[1]), pack_bf16(w[2], w[3]), pack_bf16(w[4], w[5]), pack_bf16(w[6], w[7]));
          }
        }
        fence_before_sync();
        named_bar_sync(1, T256_CTHREADS);
        if (tid == 0) mbar_arrive(&bar_qkvfree);
        const int b = g & 1;
        {
          const uint32_t nb = (uint32_t)(it * (G / 2) + (g >> 1));
          mbar_wait(&bar_ctxfree[b], (nb & 1u) ^ 1u);          // the out-projection that read this ctx buffer two groups ago retired
        }
        for (int hp = warp; hp < 8 * GH; hp += T256_CTHREADS / 32) {
          const int pair = hp >> 1, half = hp & 1, s = pair / GH, hl = pair % GH;
          const int64_t seq = a.seq0 + (int64_t)tile * 4 + s;
          const uint64_t w_pair = (uint64_t)((seq * H + (g * GH + hl)) * 32) * 16u;
          t256_attn_fwd<DH>(sQKV, sCtx + b * 8192, s, hl, half, lane, a.d_attn, w_pair);
        }
        fence_async_smem();
        named_bar_sync(1, T256_CTHREADS);
        if (tid == 0) mbar_arrive(&bar_ctxready[b]);
      }
      // ---- P2: + bias, dropout, + residual, LayerNorm1 -> x1 (registers, fp32) and its bf16 A image ----
      float u[64];
      mbar_wait(&bar_outfull, (uint32_t)it & 1u);
      fence_after_sync();
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) tmem_ld16(t_out + lane_off + (uint32_t)(part * 64 + cb * 16), u + cb * 16);
      tmem_ld_wait();
      {
        const uint64_t w0 = (uint64_t)((a.seq0 * 32 + grow) * D + part * 64) >> 1;
        const uint32_t wlo = (uint32_t)w0, xhi = (uint32_t)(w0 >> 32) * 0x85EBCA6Bu;
        const float *xr = a.x_in + grow * D + part * 64;
        const float *bo = p_bo + part * 64;
        float s1 = 0.f;
#pragma unroll
        for (int c = 0; c < 64; c += 4) {
          const float4 xv = valid ? *reinterpret_cast<const float4 *>(xr + c) : make_float4(0, 0, 0, 0);
          float m0 = 1.f, m1 = 1.f, m2 = 1.f, m3 = 1.f;
          if (a.d1.thr) {
            const uint32_t h0 = drop_hash(wlo + (uint32_t)(c >> 1), xhi, a.d1.key), h1 = drop_hash(wlo + (uint32_t)(c >> 1) + 1u, xhi, a.d1.key);
            m0 = ((h0 & 0xFFFFu) >= a.d1.thr) ? a.d1.scale : 0.f; m1 = ((h0 >> 16) >= a.d1.thr) ? a.d1.scale : 0.f;
            m2 = ((h1 & 0xFFFFu) >= a.d1.thr) ? a.d1.scale : 0.f; m3 = ((h1 >> 16) >= a.d1.thr) ? a.d1.scale : 0.f;
          }
          u[c] = xv.x + (u[c] + bo[c]) * m0; u[c + 1] = xv.y + (u[c + 1] + bo[c + 1]) * m1;
          u[c + 2] = xv.z + (u[c + 2] + bo[c + 2]) * m2; u[c + 3] = xv.w + (u[c + 3] + bo[c + 3]) * m3;
          s1 += (u[c] + u[c + 1]) + (u[c + 2] + u[c + 3]);
          if (a.u1 && valid) *reinterpret_cast<float4 *>(a.u1 + grow * D + part * 64 + c) = make_float4(u[c], u[c + 1], u[c + 2], u[c + 3]);
        }
        sStatA[row * 4 + part] = s1;
        named_bar_sync(1, T256_CTHREADS);
        const float4 sa = *reinterpret_cast<const float4 *>(sStatA + row * 4);
        const float mu = ((sa.x + sa.y) + (sa.z + sa.w)) * (1.f / D);
        float qq = 0.f;
#pragma unroll
        for (int c = 0; c < 64; ++c) { u[c] -= mu; qq = fmaf(u[c], u[c], qq); }
        sStatB[row * 4 + part] = qq;
        named_bar_sync(1, T256_CTHREADS);
        const float4 sb = *reinterpret_cast<const float4 *>(sStatB + row * 4);
        const float rs = rsqrtf(((sb.x + sb.y) + (sb.z + sb.w)) * (1.f / D) + LN_EPS);
        const float *g1 = p_g1 + part * 64, *be1 = p_be1 + part * 64;
#pragma unroll
        for (int c = 0; c < 64; c += 8) {
#pragma unroll
          for (int j = 0; j < 8; ++j) u[c + j] = u[c + j] * rs * g1[c + j] + be1[c + j];
          *reinterpret_cast<uint4 *>(sX + kmajor_off(row, part * 64 + c, 128)) =
              make_uint4(pack_bf16(u[c], u[c + 1]), pack_bf16(u[c + 2], u[c + 3]), pack_bf16(u[c + 4], u[c + 5]), pack_bf16(u[c + 6], u[c + 7]));
        }
      }
      fence_async_smem();
      fence_before_sync();
      named_bar_sync(1, T256_CTHREADS);
      if (tid == 0) mbar_arrive(&bar_x1ready);
      // ---- P3: FFN hidden chunks: + bias, ReLU, dropout -> bf16 H image ----
      for (int c = 0; c < NCH; ++c) {
        const uint32_t nh = (uint32_t)(it * NCH + c);
        mbar_wait(&bar_hfull, nh & 1u);
        fence_after_sync();
        float v[16];
        tmem_ld16(t_h + lane_off + (uint32_t)(part * 16), v);
        tmem_ld_wait();
        const uint64_t w0 = (uint64_t)((a.seq0 * 32 + grow) * F + c * 64 + part * 16) >> 1;
        const uint32_t wlo = (uint32_t)w0, xhi = (uint32_t)(w0 >> 32) * 0x85EBCA6Bu;
        const float *b1 = p_b1 + c * 64 + part * 16;
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          float h0 = fmaxf(v[j] + b1[j], 0.f), h1 = fmaxf(v[j + 1] + b1[j + 1], 0.f);
          if (a.d_ffn.thr) {
            const uint32_t hs = drop_hash(wlo + (uint32_t)(j >> 1), xhi, a.d_ffn.key);
            h0 = ((hs & 0xFFFFu) >= a.d_ffn.thr) ? h0 * a.d_ffn.scale : 0.f;
            h1 = ((hs >> 16) >= a.d_ffn.thr) ? h1 * a.d_ffn.scale : 0.f;
          }
          pk[j >> 1] = pack_bf16(h0, h1);
        }
        *reinterpret_cast<uint4 *>(sH + kmajor_off(row, part * 16, 128)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4 *>(sH + kmajor_off(row, part * 16 + 8, 128)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        fence_async_smem();
        fence_before_sync();
        named_bar_sync(1, T256_CTHREADS);
        if (tid == 0) mbar_arrive(&bar_hready);
      }
      // ---- P4: + bias, dropout, + residual (x1), LayerNorm2 -> x_out ----
      mbar_wait(&bar_out2full, (uint32_t)it & 1u);
      fence_after_sync();
      {
        const uint64_t w0 = (uint64_t)((a.seq0 * 32 + grow) * D + part * 64) >> 1;
        const uint32_t wlo = (uint32_t)w0, xhi = (uint32_t)(w0 >> 32) * 0x85EBCA6Bu;
        const float *b2 = p_b2 + part * 64;
        float s1 = 0.f;
#pragma unroll
        for (int cb = 0; cb < 64; cb += 16) {
          float f[16];
          tmem_ld16(t_out + lane_off + (uint32_t)(part * 64 + cb), f);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            float m0 = 1.f, m1 = 1.f;
            if (a.d2.thr) {
              const uint32_t hs = drop_hash(wlo + (uint32_t)((cb + j) >> 1), xhi, a.d2.key);
              m0 = ((hs & 0xFFFFu) >= a.d2.thr) ? a.d2.scale : 0.f; m1 = ((hs >> 16) >= a.d2.thr) ? a.d2.scale : 0.f;
            }
            u[cb + j] += (f[j] + b2[cb + j]) * m0;
            u[cb + j + 1] += (f[j + 1] + b2[cb + j + 1]) * m1;
            s1 += u[cb + j] + u[cb + j + 1];
          }
          if (a.u2 && valid) {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              *reinterpret_cast<float4 *>(a.u2 + grow * D + part * 64 + cb + j) = make_float4(u[cb + j], u[cb + j + 1], u[cb + j + 2], u[cb + j + 3]);
          }
        }
        sStatA[row * 4 + part] = s1;
        named_bar_sync(1, T256_CTHREADS);
        const float4 sa = *reinterpret_cast<const float4 *>(sStatA + row * 4);
        const float mu = ((sa.x + sa.y) + (sa.z + sa.w)) * (1.f / D);
        float qq = 0.f;
#pragma unroll
        for (int c = 0; c < 64; ++c) { u[c] -= mu; qq = fmaf(u[c], u[c], qq); }
        sStatB[row * 4 + part] = qq;
        named_bar_sync(1, T256_CTHREADS);
        const float4 sb = *reinterpret_cast<const float4 *>(sStatB + row * 4);
        const float rs = rsqrtf(((sb.x + sb.y) + (sb.z + sb.w)) * (1.f / D) + LN_EPS);
        const float *g2 = p_g2 + part * 64, *be2 = p_be2 + part * 64;
        if (valid) {
#pragma unroll
          for (int c = 0; c < 64; c += 4)
            *reinterpret_cast<float4 *>(a.x_out + grow * D + part * 64 + c) =
                make_float4(u[c] * rs * g2[c] + be2[c], u[c + 1] * rs * g2[c + 1] + be2[c + 1], u[c + 2] * rs * g2[c + 2] + be2[c + 2],
                            u[c + 3] * rs * g2[c + 3] + be2[c + 3]);
        }
      }
      fence_before_sync();
      named_bar_sync(1, T256_CTHREADS);
    }
  }
  __syncwarp();
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

static int t256_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int DH>
static int t256_launch_fwd(const TcLayerArgs &a, int grid, cudaStream_t st) {
  GT_CUDA(cudaFuncSetAttribute(t256_layer_fwd_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T256FwdSmem::total));
  { LaunchScope _ls(KC_TC_LAYER_FWD, st);
    t256_layer_fwd_kernel<DH><<<grid, T256_THREADS, T256FwdSmem::total, st>>>(a); }
  GT_CUDA(cudaGetLastError());
  return 0;
}

int t256_layer_fwd(const TcLayerArgs &a, cudaStream_t st) {
  GT_CHECK(a.F % 64 == 0 && a.F >= 64 && a.F <= 512, "t256_layer_fwd: dim_feedforward not supported");
  const int grid = a.n_tiles < t256_num_sms() ? a.n_tiles : t256_num_sms();
  switch (a.dh) {
    case 16: return t256_launch_fwd<16>(a, grid, st);
    case 32: return t256_launch_fwd<32>(a, grid, st);
    default: GT_FAIL("t256_layer_fwd: head dim not instantiated");
  }
}

int t256_layer_bwd(const TcLayerArgs &, cudaStream_t) { GT_FAIL("t256_layer_bwd: not built yet"); }

}  // namespace gt
