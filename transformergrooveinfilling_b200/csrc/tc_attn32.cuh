// tc_attn32.cuh — attention forward / backward of the d_model = 32 fused layer kernels on warp-level
// mma.sync (bf16 operands, fp32 accumulate) for head dims 2, 4 and 8.
//
// One warp owns one (sequence, head) pair: 32 queries x 32 keys.  q (pre-scaled by log2(e)/sqrt(dh)), k, v
// (and dO in backward) live in shared memory as bf16 ROW-MAJOR token images [128 rows][32 columns + 8 pad]
// (80-byte rows: ldmatrix rows of one 8x8 tile fall into eight distinct 16-byte bank groups).  Heads narrower
// than 8 columns share an 8-column block; the A operand (q rows / dO rows) is zeroed outside the head's own
// columns, so the other heads' columns of the block contribute nothing to a contraction over head features,
// and only the lanes that own the head's columns write results.
//   S = Q K^T , dP = dO V^T       m16n8k8   (A: masked row words, B: ldmatrix of the K / V block)
//   O = P V , dQ = dS K           m16n8k16  (A: packed accumulator fragments, B: ldmatrix.trans of V / K)
//   dK = dS^T Q , dV = P^T dO     m16n8k16  (A: movmatrix-transposed fragments, B: ldmatrix.trans of Q / dO)
// Dropout masks: common.cuh hash_quad / key_perm — the four probabilities a thread owns for one query row in
// n-tiles {0,1} (or {2,3}) are one hash quad, exactly like the d_model = 256 kernels and the oracle.
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace gt {
using namespace umma;

constexpr uint32_t A32_ROWB = 80;                 // bytes per token row of a bf16 attention image
constexpr uint32_t A32_IMG = 128 * A32_ROWB;      // one image: 10240 bytes

// D(16x8, f32) += A(16x8 bf16, row) * B(8x8 bf16, col)
__device__ __forceinline__ void mma1688(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5}, {%6}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(b0));
}
// (ldmatrix_x4 — four 8x8 b16 matrices, not transposed: thread (g, t) receives word t of row g of each matrix — lives in umma.cuh)

// write 32 fp32 values of one token row as bf16 into an attention image
__device__ __forceinline__ void a32_store_row(uint8_t *img, int row, const float (&v)[32], float scale) {
  uint4 *dst = reinterpret_cast<uint4 *>(img + (uint32_t)row * A32_ROWB);
#pragma unroll
  for (int c = 0; c < 32; c += 8)
    dst[c >> 3] = make_uint4(pack_bf16(v[c] * scale, v[c + 1] * scale), pack_bf16(v[c + 2] * scale, v[c + 3] * scale),
                             pack_bf16(v[c + 4] * scale, v[c + 5] * scale), pack_bf16(v[c + 6] * scale, v[c + 7] * scale));
}

// softmax over the 32 keys of the four query rows a thread touches (rows g, g+8 of m-tiles 0 and 1); on return
// sacc holds exp2(s - max) and inv[mt][0/1] = 1 / rowsum
__device__ __forceinline__ void a32_softmax(float (&sacc)[2][4][4], float (&inv)[2][2]) {
  float m[2][2], sm[2][2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    m[u][0] = sacc[u][0][0]; m[u][1] = sacc[u][0][2];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      m[u][0] = fmaxf(m[u][0], fmaxf(sacc[u][nt][0], sacc[u][nt][1]));
      m[u][1] = fmaxf(m[u][1], fmaxf(sacc[u][nt][2], sacc[u][nt][3]));
    }
  }
#pragma unroll
  for (int o = 1; o <= 2; o <<= 1)
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      m[u][0] = fmaxf(m[u][0], __shfl_xor_sync(0xffffffffu, m[u][0], o));
      m[u][1] = fmaxf(m[u][1], __shfl_xor_sync(0xffffffffu, m[u][1], o));
    }
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    sm[u][0] = 0.f; sm[u][1] = 0.f;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      sacc[u][nt][0] = ex2_ftz(sacc[u][nt][0] - m[u][0]); sacc[u][nt][1] = ex2_ftz(sacc[u][nt][1] - m[u][0]);
      sacc[u][nt][2] = ex2_ftz(sacc[u][nt][2] - m[u][1]); sacc[u][nt][3] = ex2_ftz(sacc[u][nt][3] - m[u][1]);
      sm[u][0] += sacc[u][nt][0] + sacc[u][nt][1]; sm[u][1] += sacc[u][nt][2] + sacc[u][nt][3];
    }
  }
#pragma unroll
  for (int o = 1; o <= 2; o <<= 1)
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      sm[u][0] += __shfl_xor_sync(0xffffffffu, sm[u][0], o);
      sm[u][1] += __shfl_xor_sync(0xffffffffu, sm[u][1], o);
    }
#pragma unroll
  for (int u = 0; u < 2; ++u) { inv[u][0] = rcp_approx(sm[u][0]); inv[u][1] = rcp_approx(sm[u][1]); }
}

// ---- forward: ctx rows of one (sequence s, head h) pair -> bf16 K-major A image sCtx --------------------------
// w_pair = quad index (element index >> 2) of (query row 0, position 0) of this pair at the attention dropout site
// CAUSAL: key j of query i is masked for j > i (BGT/models/utils.py:53-56 get_tgt_mask; decoder self-attention)
// The probabilities enter the P V contraction UNNORMALISED: E = bf16(exp2(s - max)) with the dropped keys zeroed by a packed
// mask (umma.cuh: keep2), and the row factor 1 / (rowsum (1 - p)) is applied to the DH context values instead of the 32
// probabilities:  ctx = bf16( (E_kept V) / (rowsum (1 - p)) ).
template <int DH, bool CAUSAL = false>
__device__ __forceinline__ void a32_attn_fwd(const uint8_t *sQ, const uint8_t *sK, const uint8_t *sV, uint8_t *sCtx, int s, int h, int lane,
                                             const Drop &dr, uint64_t w_pair) {
  static_assert(DH == 2 || DH == 4 || DH == 8, "mma attention path: head dim 2, 4 or 8");
  const int g = lane >> 2, t = lane & 3;
  const int blk = (h * DH) >> 3, cin = (h * DH) & 7;
  const bool own = (unsigned)(2 * t - cin) < (unsigned)DH;       // this lane's column pair (2t, 2t+1) of the block belongs to head h
  const uint32_t lrow = (uint32_t)(s * 32 + (lane & 7) + 8 * (lane >> 3)) * A32_ROWB + (uint32_t)blk * 16u;   // ldmatrix: row (l & 7) of matrix (l >> 3)
  uint32_t bk[4], bv[4];
  ldmatrix_x4(bk, sK + lrow);
  ldmatrix_x4_trans(bv, sV + lrow);
  float sacc[2][4][4];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const uint32_t ro = (uint32_t)(s * 32 + 16 * u + g) * A32_ROWB + (uint32_t)blk * 16u + (uint32_t)t * 4u;
    const uint32_t a0 = own ? lds32(sQ + ro) : 0u, a1 = own ? lds32(sQ + ro + 8u * A32_ROWB) : 0u;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      sacc[u][nt][0] = 0.f; sacc[u][nt][1] = 0.f; sacc[u][nt][2] = 0.f; sacc[u][nt][3] = 0.f;
      mma1688(sacc[u][nt], a0, a1, bk[nt]);
      if constexpr (CAUSAL) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (8 * nt + 2 * t + (c & 1) > 16 * u + g + (c >> 1) * 8) sacc[u][nt][c] = -1e30f;
      }
    }
  }
  float inv[2][2];
  a32_softmax(sacc, inv);
  const uint32_t thr2 = dr.thr | (dr.thr << 16);
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    uint32_t pa[4][2];                                  // [key n-tile][row g / g + 8]: packed pairs (keys 8 nt + 2t, + 1)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      pa[nt][0] = pack_bf16(sacc[u][nt][0], sacc[u][nt][1]);
      pa[nt][1] = pack_bf16(sacc[u][nt][2], sacc[u][nt][3]);
    }
    if (dr.thr) {
      const int q0 = 16 * u + g;
      const uint64_t wa = w_pair + (uint64_t)q0 * 8u, wb = wa + 64u;            // rows q0 and q0 + 8
      const uint32_t alo = (uint32_t)wa, ahi = (uint32_t)(wa >> 32) * 0x85EBCA6Bu;
      const uint32_t blo = (uint32_t)wb, bhi = (uint32_t)(wb >> 32) * 0x85EBCA6Bu;
#pragma unroll
      for (int np = 0; np < 2; ++np) {                  // quad 4 np + t of a row = this lane's keys of n-tiles 2 np, 2 np + 1 (common.cuh: key_perm)
        uint32_t la, ha, lb, hb;
        hash_quad((alo + (uint32_t)(4 * np + t)) ^ ahi, dr.key, la, ha);
        hash_quad((blo + (uint32_t)(4 * np + t)) ^ bhi, dr.key, lb, hb);
        pa[2 * np][0] &= keep2(la, thr2); pa[2 * np + 1][0] &= keep2(ha, thr2);
        pa[2 * np][1] &= keep2(lb, thr2); pa[2 * np + 1][1] &= keep2(hb, thr2);
      }
    }
    // O = E V : keys 16 kt .. 16 kt + 15 per k-step
    float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int kt = 0; kt < 2; ++kt) mma16816(o, pa[2 * kt][0], pa[2 * kt][1], pa[2 * kt + 1][0], pa[2 * kt + 1][1], bv[2 * kt], bv[2 * kt + 1]);
    if (own) {
      const float i0 = inv[u][0] * dr.scale, i1 = inv[u][1] * dr.scale;
      const int r0 = s * 32 + 16 * u + g, col = blk * 8 + 2 * t;
      *reinterpret_cast<uint32_t *>(sCtx + kmajor_off(r0, col, 128)) = pack_bf16(o[0] * i0, o[1] * i0);
      *reinterpret_cast<uint32_t *>(sCtx + kmajor_off(r0 + 8, col, 128)) = pack_bf16(o[2] * i1, o[3] * i1);
    }
  }
}

// ---- backward of one (sequence s, head h) pair -------------------------------------------------------------------
// sDO: dL/dctx rows.  Writes the recomputed ctx rows to sCtx (for dWo) and dq | dk | dv to the K-major image sDQ
// [128 x 96] (columns [0,32) dq wrt the UNscaled q, [32,64) dk, [64,96) dv).  The in-projection bias gradient is the
// column sum of that image, taken on the tensor cores by the caller (tc_layers.cu: colsum_image).
// Arithmetic per probability (c = 1/sqrt(dh), ks = 1/(1-p), P = E / rowsum):
//   p1 = P c ;  pd = keep ? p1 ks : 0  (= c x the dropped probability) ;  delta = sum_j pd dP / c  (= sum_j P dP_dropped)
//   dS c = pd dP - p1 delta            -> bf16 -> dq = (dS c) K ,  dk = (dS c)^T Qs ln2 / c  (Qs = q log2(e) c, as staged)
//   ctx  = (pd V) / c ,  dv = (pd^T dO) / c                        (the 1 / c lands on the DH-wide results, not on 32 probabilities)
template <int DH, bool CAUSAL = false>
__device__ __forceinline__ void a32_attn_bwd(const uint8_t *sQ, const uint8_t *sK, const uint8_t *sV, const uint8_t *sDO, uint8_t *sCtx, uint8_t *sDQ,
                                             int s, int h, int lane, const Drop &dr, uint64_t w_pair) {
  static_assert(DH == 2 || DH == 4 || DH == 8, "mma attention path: head dim 2, 4 or 8");
  const int g = lane >> 2, t = lane & 3;
  const int blk = (h * DH) >> 3, cin = (h * DH) & 7;
  const bool own = (unsigned)(2 * t - cin) < (unsigned)DH;
  const float c1 = rsqrtf((float)DH), rc1 = sqrtf((float)DH), ks = dr.scale;
  const float dk_scale = 0.6931471805599453f * rc1;
  const uint32_t lrow = (uint32_t)(s * 32 + (lane & 7) + 8 * (lane >> 3)) * A32_ROWB + (uint32_t)blk * 16u;
  uint32_t bk[4], bv[4], bkt[4], bvt[4], bqt[4], bot[4];
  ldmatrix_x4(bk, sK + lrow);
  ldmatrix_x4(bv, sV + lrow);
  ldmatrix_x4_trans(bkt, sK + lrow);
  ldmatrix_x4_trans(bvt, sV + lrow);
  ldmatrix_x4_trans(bqt, sQ + lrow);
  ldmatrix_x4_trans(bot, sDO + lrow);
  float dk[2][4], dv[2][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) { dk[a][c] = 0.f; dv[a][c] = 0.f; }
  const int col = blk * 8 + 2 * t;

#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    const int r0 = s * 32 + 16 * mt + g;
    const uint32_t ro = (uint32_t)r0 * A32_ROWB + (uint32_t)blk * 16u + (uint32_t)t * 4u;
    const uint32_t a0 = own ? lds32(sQ + ro) : 0u, a1 = own ? lds32(sQ + ro + 8u * A32_ROWB) : 0u;
    const uint32_t o0 = own ? lds32(sDO + ro) : 0u, o1 = own ? lds32(sDO + ro + 8u * A32_ROWB) : 0u;
    float p[4][4], dp[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
      for (int c = 0; c < 4; ++c) { p[nt][c] = 0.f; dp[nt][c] = 0.f; }
      mma1688(p[nt], a0, a1, bk[nt]);
      mma1688(dp[nt], o0, o1, bv[nt]);
      if constexpr (CAUSAL) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (8 * nt + 2 * t + (c & 1) > 16 * mt + g + (c >> 1) * 8) p[nt][c] = -1e30f;
      }
    }
    float m0 = p[0][0], m1 = p[0][2];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) { m0 = fmaxf(m0, fmaxf(p[nt][0], p[nt][1])); m1 = fmaxf(m1, fmaxf(p[nt][2], p[nt][3])); }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      p[nt][0] = ex2_ftz(p[nt][0] - m0); p[nt][1] = ex2_ftz(p[nt][1] - m0);
      p[nt][2] = ex2_ftz(p[nt][2] - m1); p[nt][3] = ex2_ftz(p[nt][3] - m1);
      s0 += p[nt][0] + p[nt][1]; s1 += p[nt][2] + p[nt][3];
    }
    s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
    const float i0 = rcp_approx(s0) * c1, i1 = rcp_approx(s1) * c1;
    // pd = keep ? p1 ks : 0 in place of dp's partner: pdm[nt][c]
    float pdm[4][4];
    if (dr.thr) {
      const int q0 = 16 * mt + g;
      const uint64_t wa = w_pair + (uint64_t)q0 * 8u, wb = wa + 64u;
      const uint32_t alo = (uint32_t)wa, ahi = (uint32_t)(wa >> 32) * 0x85EBCA6Bu;
      const uint32_t blo = (uint32_t)wb, bhi = (uint32_t)(wb >> 32) * 0x85EBCA6Bu;
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t la, ha, lb, hb;
        hash_quad((alo + (uint32_t)(4 * np + t)) ^ ahi, dr.key, la, ha);
        hash_quad((blo + (uint32_t)(4 * np + t)) ^ bhi, dr.key, lb, hb);
        const float k0 = i0 * ks, k1 = i1 * ks;
        pdm[2 * np][0] = ((la & 0xFFFFu) >= dr.thr) ? p[2 * np][0] * k0 : 0.f;
        pdm[2 * np][1] = ((la >> 16) >= dr.thr) ? p[2 * np][1] * k0 : 0.f;
        pdm[2 * np + 1][0] = ((ha & 0xFFFFu) >= dr.thr) ? p[2 * np + 1][0] * k0 : 0.f;
        pdm[2 * np + 1][1] = ((ha >> 16) >= dr.thr) ? p[2 * np + 1][1] * k0 : 0.f;
        pdm[2 * np][2] = ((lb & 0xFFFFu) >= dr.thr) ? p[2 * np][2] * k1 : 0.f;
        pdm[2 * np][3] = ((lb >> 16) >= dr.thr) ? p[2 * np][3] * k1 : 0.f;
        pdm[2 * np + 1][2] = ((hb & 0xFFFFu) >= dr.thr) ? p[2 * np + 1][2] * k1 : 0.f;
        pdm[2 * np + 1][3] = ((hb >> 16) >= dr.thr) ? p[2 * np + 1][3] * k1 : 0.f;
      }
    } else {
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) { pdm[nt][0] = p[nt][0] * i0; pdm[nt][1] = p[nt][1] * i0; pdm[nt][2] = p[nt][2] * i1; pdm[nt][3] = p[nt][3] * i1; }
    }
    float d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      dp[nt][0] *= pdm[nt][0]; dp[nt][1] *= pdm[nt][1]; dp[nt][2] *= pdm[nt][2]; dp[nt][3] *= pdm[nt][3];      // pd dP
      d0 += dp[nt][0] + dp[nt][1];
      d1 += dp[nt][2] + dp[nt][3];
    }
    d0 += __shfl_xor_sync(0xffffffffu, d0, 1); d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 1); d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
    const float e0 = -(d0 * rc1) * i0, e1 = -(d1 * rc1) * i1;          // - delta x (p1 / E)
    uint32_t pdp[4][2], dsq[4][2];                      // c x dropped P ; c x dS
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      pdp[nt][0] = pack_bf16(pdm[nt][0], pdm[nt][1]); pdp[nt][1] = pack_bf16(pdm[nt][2], pdm[nt][3]);
      dsq[nt][0] = pack_bf16(fmaf(p[nt][0], e0, dp[nt][0]), fmaf(p[nt][1], e0, dp[nt][1]));
      dsq[nt][1] = pack_bf16(fmaf(p[nt][2], e1, dp[nt][2]), fmaf(p[nt][3], e1, dp[nt][3]));
    }
    float o[4] = {0.f, 0.f, 0.f, 0.f}, dq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int kt = 0; kt < 2; ++kt) {                   // contraction over keys 16 kt .. 16 kt + 15
      mma16816(o, pdp[2 * kt][0], pdp[2 * kt][1], pdp[2 * kt + 1][0], pdp[2 * kt + 1][1], bvt[2 * kt], bvt[2 * kt + 1]);
      mma16816(dq, dsq[2 * kt][0], dsq[2 * kt][1], dsq[2 * kt + 1][0], dsq[2 * kt + 1][1], bkt[2 * kt], bkt[2 * kt + 1]);
    }
#pragma unroll
    for (int kmt = 0; kmt < 2; ++kmt) {                // dk / dv rows = keys 16 kmt .. ; contraction over this m-tile's 16 queries
      const uint32_t s0t = movmatrix_trans(dsq[2 * kmt][0]), s1t = movmatrix_trans(dsq[2 * kmt + 1][0]);
      const uint32_t s2t = movmatrix_trans(dsq[2 * kmt][1]), s3t = movmatrix_trans(dsq[2 * kmt + 1][1]);
      mma16816(dk[kmt], s0t, s1t, s2t, s3t, bqt[2 * mt], bqt[2 * mt + 1]);
      const uint32_t p0t = movmatrix_trans(pdp[2 * kmt][0]), p1t = movmatrix_trans(pdp[2 * kmt + 1][0]);
      const uint32_t p2t = movmatrix_trans(pdp[2 * kmt][1]), p3t = movmatrix_trans(pdp[2 * kmt + 1][1]);
      mma16816(dv[kmt], p0t, p1t, p2t, p3t, bot[2 * mt], bot[2 * mt + 1]);
    }
    if (own) {
      *reinterpret_cast<uint32_t *>(sCtx + kmajor_off(r0, col, 128)) = pack_bf16(o[0] * rc1, o[1] * rc1);
      *reinterpret_cast<uint32_t *>(sCtx + kmajor_off(r0 + 8, col, 128)) = pack_bf16(o[2] * rc1, o[3] * rc1);
      *reinterpret_cast<uint32_t *>(sDQ + kmajor_off(r0, col, 128)) = pack_bf16(dq[0], dq[1]);
      *reinterpret_cast<uint32_t *>(sDQ + kmajor_off(r0 + 8, col, 128)) = pack_bf16(dq[2], dq[3]);
    }
  }
  if (own) {
#pragma unroll
    for (int kmt = 0; kmt < 2; ++kmt) {
      const int kr = s * 32 + 16 * kmt + g;
      *reinterpret_cast<uint32_t *>(sDQ + kmajor_off(kr, 32 + col, 128)) = pack_bf16(dk[kmt][0] * dk_scale, dk[kmt][1] * dk_scale);
      *reinterpret_cast<uint32_t *>(sDQ + kmajor_off(kr + 8, 32 + col, 128)) = pack_bf16(dk[kmt][2] * dk_scale, dk[kmt][3] * dk_scale);
      *reinterpret_cast<uint32_t *>(sDQ + kmajor_off(kr, 64 + col, 128)) = pack_bf16(dv[kmt][0] * rc1, dv[kmt][1] * rc1);
      *reinterpret_cast<uint32_t *>(sDQ + kmajor_off(kr + 8, 64 + col, 128)) = pack_bf16(dv[kmt][2] * rc1, dv[kmt][3] * rc1);
    }
  }
}

}  // namespace gt
