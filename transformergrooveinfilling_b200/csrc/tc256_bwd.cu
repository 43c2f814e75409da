// tc256_bwd.cu — backward of the d_model = 256 fused encoder layer: a data-gradient kernel (same
// weight-streaming structure as the forward: producer lane, MMA-issuer lane, 16 compute warps) and a
// weight-gradient kernel that contracts the saved bf16 operand images over all tokens.
//
// data-gradient kernel, per tile of 128 tokens (torch/nn/modules/transformer.py:951-956 backwards):
//   B0  LayerNorm2 backward (dy, u2)            -> du2 (tcgen05.st into the dx accumulator), da2 = du2 * mask2 (bf16 image)
//   B1  dH(c)  = da2 W2[:, chunk c]             UMMA 128x64x256      B2  relu/dropout mask from the saved H image
//   B3  dx1   += dH(c) W1[chunk c, :]           UMMA 128x256x64, accumulating onto du2
//   B4  LayerNorm1 backward (du2 + dx1, u1)     -> du1 (stored back into the accumulator), da1 = du1 * mask1 (bf16 image)
//   B5  dctx   = da1 Wo                         UMMA 128x256x256  -> bf16 -> per-CTA scratch (L2)
//   B6  per head group: recompute q|k|v (UMMA 128x192x256), attention backward on mma.sync fragments,
//       dx_in += dqkv_g Wqkv[group rows, :]     UMMA 128x256x192
//   B7  dx = du1 + dx_in  (what the accumulator holds: the dx_in UMMAs accumulate onto du1)
// Every bf16 image a weight gradient needs (da2, dH, da1, dqkv here; x, x1, ctx, H from the forward) is
// left in HBM in the canonical UMMA layout, so the weight-gradient kernel stages them with plain bulk
// copies and feeds them to the tensor core as MN-major operands (contraction over the token rows).
#include <stdlib.h>

#include "tc256_dev.cuh"

namespace gt {

struct T256BwdSmem {
  // r1: da2 image -> da1 image -> attention scratch [128 x 256] = q | k | v | dO(group)
  // r2: FFN phase: two slots of (H chunk image +0, dH chunk image +16384), 32 KB each, used alternately ; attention phase: x image
  static constexpr uint32_t r1 = 0, r2 = 65536, ring = 131072, par = 196608, gpar = par + 1280 * 4, stat = gpar + 2816 * 4,
                            total = stat + 16384;          // stat: LayerNorm row statistics (4 KB) ; head_dim 128: fragment exchange (16 KB)
};
static_assert(T256BwdSmem::total <= 227 * 1024 - 1024, "backward shared memory budget");

// ---- attention backward for one (sequence s, head hl of the group) pair; one warp ---------------------------------
// sS: scratch image [128 x 256]: cols [0,64) q (scaled by log2e/sqrt(dh)), [64,128) k, [128,192) v, [192,256) dO.
// dq / dk / dv overwrite q / k / v in place (this warp is the only reader of those rows x columns).
// g_b: shared-memory partial sums of the in-projection bias gradient for this group: [3][64].
// Arithmetic of tc_attn32.cuh: with c = 1/sqrt(dh), E = exp2(s - max) and r = rowsum(E), the two bf16 fragments are
// c Pd = keep ? E (c ks / r) : 0 and c dS = (c Pd) dO.V - E (c / r) rowsum((c Pd) dO.V) sqrt(dh) — one multiply and one FMA per
// probability after the exponential; dq = (c dS) K needs no scale, dk and dv take theirs (ln2 sqrt(dh), sqrt(dh)) when stored.
// Operand fragments of the score contractions come from ldmatrix (a core matrix of the image = 8 rows x 16 bytes).
template <int DH>
__device__ __forceinline__ void t256_attn_bwd(uint8_t *sS, int s, int hl, int lane, const Drop &dr, uint64_t w_pair, float *g_b) {
  const int g = lane >> 2, t = lane & 3;
  const int qc = hl * DH, kc = 64 + hl * DH, vc = 128 + hl * DH, oc = 192 + hl * DH;
  const float c1 = rsqrtf((float)DH), rc1 = sqrtf((float)DH), ks = dr.scale;
  const float dk_scale = 0.6931471805599453f * rc1;
  const int mi = lane >> 3, rr = lane & 7;
  // lane address of the B-operand tiles (keys x 16 features): matrix (l >> 3) = keys 8 (mi >> 1).., features 8 (mi & 1)..
  const uint32_t offB = kmajor_off(s * 32 + (mi >> 1) * 8 + rr, kc + (mi & 1) * 8, 128);
  float dk[2][DH / 8][4], dv[2][DH / 8][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < DH / 8; ++b)
#pragma unroll
      for (int c = 0; c < 4; ++c) { dk[a][b][c] = 0.f; dv[a][b][c] = 0.f; }
  float sq[DH / 8][2];                               // column sums of dq over this lane's rows
#pragma unroll
  for (int b = 0; b < DH / 8; ++b) { sq[b][0] = 0.f; sq[b][1] = 0.f; }

#pragma unroll 1
  for (int mt = 0; mt < 2; ++mt) {
    const int r0 = s * 32 + 16 * mt + g;             // query rows r0, r0 + 8
    // A-operand tiles (queries x 16 features): matrix (l >> 3) = rows 8 (mi & 1).., features 8 (mi >> 1)..
    const uint32_t offA = kmajor_off(s * 32 + 16 * mt + (mi & 1) * 8 + rr, qc + (mi >> 1) * 8, 128);
    float p[4][4], dp[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int c = 0; c < 4; ++c) { p[i][c] = 0.f; dp[i][c] = 0.f; }
#pragma unroll
    for (int kt = 0; kt < DH / 16; ++kt) {
      const uint32_t cstep = (uint32_t)kt * 4096u;    // 16 feature columns = two 8-column slabs of 2048 B
      uint32_t aq[4], ao[4], bk0[4], bk1[4], bv0[4], bv1[4];
      ldmatrix_x4(aq, sS + offA + cstep);
      ldmatrix_x4(ao, sS + offA + 192u * 256u + cstep);
      ldmatrix_x4(bk0, sS + offB + cstep);
      ldmatrix_x4(bk1, sS + offB + 256u + cstep);
      ldmatrix_x4(bv0, sS + offB + 64u * 256u + cstep);
      ldmatrix_x4(bv1, sS + offB + 64u * 256u + 256u + cstep);
      mma16816(p[0], aq[0], aq[1], aq[2], aq[3], bk0[0], bk0[1]);
      mma16816(p[1], aq[0], aq[1], aq[2], aq[3], bk0[2], bk0[3]);
      mma16816(p[2], aq[0], aq[1], aq[2], aq[3], bk1[0], bk1[1]);
      mma16816(p[3], aq[0], aq[1], aq[2], aq[3], bk1[2], bk1[3]);
      mma16816(dp[0], ao[0], ao[1], ao[2], ao[3], bv0[0], bv0[1]);
      mma16816(dp[1], ao[0], ao[1], ao[2], ao[3], bv0[2], bv0[3]);
      mma16816(dp[2], ao[0], ao[1], ao[2], ao[3], bv1[0], bv1[1]);
      mma16816(dp[3], ao[0], ao[1], ao[2], ao[3], bv1[2], bv1[3]);
    }
    // softmax (rows r0: c0,c1 ; r0+8: c2,c3)
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
    const float i0 = c1 / s0, i1 = c1 / s1;
    // pd = keep ? E (c ks / rowsum) : 0   (= c x the dropped probability)
    float pdm[4][4];
    if (dr.thr) {
      const int q0 = 16 * mt + g;
      const uint64_t wa = w_pair + (uint64_t)q0 * 8u, wb = wa + 64u;      // quad index of position 0 of rows q0 and q0 + 8
      const uint32_t alo = (uint32_t)wa, ahi = (uint32_t)(wa >> 32) * 0x85EBCA6Bu;
      const uint32_t blo = (uint32_t)wb, bhi = (uint32_t)(wb >> 32) * 0x85EBCA6Bu;
      const float k0 = i0 * ks, k1 = i1 * ks;
#pragma unroll
      for (int np = 0; np < 2; ++np) {            // quad 4 np + t = this lane's keys of nt = 2 np, 2 np + 1 (common.cuh: key_perm)
        uint32_t la, ha, lb, hb;
        hash_quad((alo + (uint32_t)(4 * np + t)) ^ ahi, dr.key, la, ha);
        hash_quad((blo + (uint32_t)(4 * np + t)) ^ bhi, dr.key, lb, hb);
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
      dp[nt][0] *= pdm[nt][0]; dp[nt][1] *= pdm[nt][1]; dp[nt][2] *= pdm[nt][2]; dp[nt][3] *= pdm[nt][3];
      d0 += dp[nt][0] + dp[nt][1];
      d1 += dp[nt][2] + dp[nt][3];
    }
    d0 += __shfl_xor_sync(0xffffffffu, d0, 1); d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 1); d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
    const float e0 = -(d0 * rc1) * i0, e1 = -(d1 * rc1) * i1;
    uint32_t pdp[4][2], dsq[4][2];                  // c x dropped probabilities, c x dS
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      pdp[nt][0] = pack_bf16(pdm[nt][0], pdm[nt][1]); pdp[nt][1] = pack_bf16(pdm[nt][2], pdm[nt][3]);
      dsq[nt][0] = pack_bf16(fmaf(p[nt][0], e0, dp[nt][0]), fmaf(p[nt][1], e0, dp[nt][1]));
      dsq[nt][1] = pack_bf16(fmaf(p[nt][2], e1, dp[nt][2]), fmaf(p[nt][3], e1, dp[nt][3]));
    }
    // dk += (c dS)^T Q ; dv += (c Pd)^T dO   (contraction over this m-tile's 16 queries; A = transposed fragments)
#pragma unroll
    for (int np = 0; np < DH / 16; ++np) {
      const int qrow = s * 32 + 16 * mt + (mi & 1) * 8 + rr;
      uint32_t bq[4], bo[4];
      ldmatrix_x4_trans(bq, sS + kmajor_off(qrow, qc + 16 * np + (mi >> 1) * 8, 128));
      ldmatrix_x4_trans(bo, sS + kmajor_off(qrow, oc + 16 * np + (mi >> 1) * 8, 128));
#pragma unroll
      for (int kmt = 0; kmt < 2; ++kmt) {
        const uint32_t s0t = movmatrix_trans(dsq[2 * kmt][0]), s1t = movmatrix_trans(dsq[2 * kmt + 1][0]);
        const uint32_t s2t = movmatrix_trans(dsq[2 * kmt][1]), s3t = movmatrix_trans(dsq[2 * kmt + 1][1]);
        mma16816(dk[kmt][2 * np], s0t, s1t, s2t, s3t, bq[0], bq[1]);
        mma16816(dk[kmt][2 * np + 1], s0t, s1t, s2t, s3t, bq[2], bq[3]);
        const uint32_t p0t = movmatrix_trans(pdp[2 * kmt][0]), p1t = movmatrix_trans(pdp[2 * kmt + 1][0]);
        const uint32_t p2t = movmatrix_trans(pdp[2 * kmt][1]), p3t = movmatrix_trans(pdp[2 * kmt + 1][1]);
        mma16816(dv[kmt][2 * np], p0t, p1t, p2t, p3t, bo[0], bo[1]);
        mma16816(dv[kmt][2 * np + 1], p0t, p1t, p2t, p3t, bo[2], bo[3]);
      }
    }
    // dq = (c dS) K
    float dq[DH / 8][4];
#pragma unroll
    for (int b = 0; b < DH / 8; ++b) { dq[b][0] = 0.f; dq[b][1] = 0.f; dq[b][2] = 0.f; dq[b][3] = 0.f; }
#pragma unroll
    for (int kt = 0; kt < 2; ++kt) {                  // keys 16 kt .. 16 kt + 15
#pragma unroll
      for (int np = 0; np < DH / 16; ++np) {
        const int key = s * 32 + 16 * kt + (mi & 1) * 8 + rr;
        uint32_t bk[4];
        ldmatrix_x4_trans(bk, sS + kmajor_off(key, kc + 16 * np + (mi >> 1) * 8, 128));
        mma16816(dq[2 * np], dsq[2 * kt][0], dsq[2 * kt][1], dsq[2 * kt + 1][0], dsq[2 * kt + 1][1], bk[0], bk[1]);
        mma16816(dq[2 * np + 1], dsq[2 * kt][0], dsq[2 * kt][1], dsq[2 * kt + 1][0], dsq[2 * kt + 1][1], bk[2], bk[3]);
      }
    }
    __syncwarp();                                     // every lane has finished reading this m-tile's q rows
#pragma unroll
    for (int nt = 0; nt < DH / 8; ++nt) {
      *reinterpret_cast<uint32_t *>(sS + kmajor_off(r0, qc + 8 * nt + 2 * t, 128)) = pack_bf16(dq[nt][0], dq[nt][1]);
      *reinterpret_cast<uint32_t *>(sS + kmajor_off(r0 + 8, qc + 8 * nt + 2 * t, 128)) = pack_bf16(dq[nt][2], dq[nt][3]);
      sq[nt][0] += dq[nt][0] + dq[nt][2]; sq[nt][1] += dq[nt][1] + dq[nt][3];
    }
  }
  __syncwarp();                                       // all reads of k / v / dO by every lane are done
#pragma unroll
  for (int kmt = 0; kmt < 2; ++kmt)
#pragma unroll
    for (int nt = 0; nt < DH / 8; ++nt) {
#pragma unroll
      for (int c = 0; c < 4; ++c) { dk[kmt][nt][c] *= dk_scale; dv[kmt][nt][c] *= rc1; }
      const int kr = s * 32 + 16 * kmt + g;
      *reinterpret_cast<uint32_t *>(sS + kmajor_off(kr, kc + 8 * nt + 2 * t, 128)) = pack_bf16(dk[kmt][nt][0], dk[kmt][nt][1]);
      *reinterpret_cast<uint32_t *>(sS + kmajor_off(kr + 8, kc + 8 * nt + 2 * t, 128)) = pack_bf16(dk[kmt][nt][2], dk[kmt][nt][3]);
      *reinterpret_cast<uint32_t *>(sS + kmajor_off(kr, vc + 8 * nt + 2 * t, 128)) = pack_bf16(dv[kmt][nt][0], dv[kmt][nt][1]);
      *reinterpret_cast<uint32_t *>(sS + kmajor_off(kr + 8, vc + 8 * nt + 2 * t, 128)) = pack_bf16(dv[kmt][nt][2], dv[kmt][nt][3]);
    }
  // in-projection bias gradient: column sums over the pair's 32 rows
#pragma unroll
  for (int nt = 0; nt < DH / 8; ++nt) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      float a = sq[nt][j];
      float b = dk[0][nt][j] + dk[0][nt][j + 2] + dk[1][nt][j] + dk[1][nt][j + 2];
      float c = dv[0][nt][j] + dv[0][nt][j + 2] + dv[1][nt][j] + dv[1][nt][j + 2];
#pragma unroll
      for (int o = 4; o <= 16; o <<= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
      }
      if (g == 0) {
        const int col = hl * DH + 8 * nt + 2 * t + j;
        atomicAdd(g_b + col, a); atomicAdd(g_b + 64 + col, b); atomicAdd(g_b + 128 + col, c);
      }
    }
  }
}

// ---- attention backward for one (sequence s, head) pair at head_dim 128; one warp ------------------------------------------
// The head's 128 features live in TWO scratch images (the even / odd 64-column group): img[half] = [128 x 256] with cols
// [0,64) q (scaled by log2e/sqrt(dh)), [64,128) k, [128,192) v, [192,256) dO of that group.  Phase A contracts the scores and
// dP = dO V^T over all 128 features and turns them into the bf16 fragments of c dS and c dropped-P (c = 1/sqrt(dh), the
// arithmetic of tc_attn32.cuh); phase B walks the features in blocks of 16 columns: dq | dk | dv overwrite q | k | v in place
// (this warp is the only reader of those rows x columns).  g_b: bias-gradient partials of the two groups, [2][3][64].
// FOUR warps share a pair (wq = 0..3).  Phase A is split by query m-tile: warps wq = 0, 1 turn the 16 queries of m-tile wq into
// the fragments and leave them in the exchange buffer xch ([m-tile][16 registers][32 lanes] words of this sequence); after one
// block-wide barrier (phase B overwrites what phase A reads) every warp loads both m-tiles' fragments and takes the feature
// blocks wq and wq + 4 of phase B.  Operand fragments come from ldmatrix (a core matrix of the image = 8 rows x 16 bytes).
__device__ __forceinline__ void t256_attn_bwd128(uint8_t *imgE, uint8_t *imgO, uint32_t *xch, int s, int wq, int lane, const Drop &dr, uint64_t w_pair,
                                                 float *g_b) {
  constexpr int DH = 128;
  const int g = lane >> 2, t = lane & 3;
  const int mi = lane >> 3, rr = lane & 7;
  const float c1 = rsqrtf((float)DH), rc1 = sqrtf((float)DH), ks = dr.scale;
  const float dk_scale = 0.6931471805599453f * rc1;
  if (wq < 2) {
    const int mt = wq;
    float p[4][4], dp[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int c = 0; c < 4; ++c) { p[i][c] = 0.f; dp[i][c] = 0.f; }
    // lane addresses: A-operand tiles (queries x 16 features): matrix (l >> 3) = rows 8 (mi & 1).., features 8 (mi >> 1)..
    //                 B-operand tiles (keys x 16 features):    matrix (l >> 3) = keys 8 (mi >> 1).., features 8 (mi & 1)..
    const uint32_t offA = kmajor_off(s * 32 + 16 * mt + (mi & 1) * 8 + rr, (mi >> 1) * 8, 128);
    const uint32_t offB = kmajor_off(s * 32 + (mi >> 1) * 8 + rr, (mi & 1) * 8, 128);
#pragma unroll 1
    for (int hf = 0; hf < 2; ++hf) {
      const uint8_t *im = hf ? imgO : imgE;
#pragma unroll
      for (int kt = 0; kt < 4; ++kt) {
        const uint32_t cstep = (uint32_t)kt * 4096u;    // 16 feature columns = two 8-column slabs of 2048 B
        uint32_t aq[4], ao[4], bk0[4], bk1[4], bv0[4], bv1[4];
        ldmatrix_x4(aq, im + offA + cstep);
        ldmatrix_x4(ao, im + offA + 192u * 256u + cstep);
        ldmatrix_x4(bk0, im + offB + 64u * 256u + cstep);
        ldmatrix_x4(bk1, im + offB + 64u * 256u + 256u + cstep);
        ldmatrix_x4(bv0, im + offB + 128u * 256u + cstep);
        ldmatrix_x4(bv1, im + offB + 128u * 256u + 256u + cstep);
        mma16816(p[0], aq[0], aq[1], aq[2], aq[3], bk0[0], bk0[1]);
        mma16816(p[1], aq[0], aq[1], aq[2], aq[3], bk0[2], bk0[3]);
        mma16816(p[2], aq[0], aq[1], aq[2], aq[3], bk1[0], bk1[1]);
        mma16816(p[3], aq[0], aq[1], aq[2], aq[3], bk1[2], bk1[3]);
        mma16816(dp[0], ao[0], ao[1], ao[2], ao[3], bv0[0], bv0[1]);
        mma16816(dp[1], ao[0], ao[1], ao[2], ao[3], bv0[2], bv0[3]);
        mma16816(dp[2], ao[0], ao[1], ao[2], ao[3], bv1[0], bv1[1]);
        mma16816(dp[3], ao[0], ao[1], ao[2], ao[3], bv1[2], bv1[3]);
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
    const float i0 = c1 / s0, i1 = c1 / s1;
    // pd = keep ? E (c ks / rowsum) : 0   (= c x the dropped probability)
    float pdm[4][4];
    if (dr.thr) {
      const int q0 = 16 * mt + g;
      const uint64_t wa = w_pair + (uint64_t)q0 * 8u, wb = wa + 64u;
      const uint32_t alo = (uint32_t)wa, ahi = (uint32_t)(wa >> 32) * 0x85EBCA6Bu;
      const uint32_t blo = (uint32_t)wb, bhi = (uint32_t)(wb >> 32) * 0x85EBCA6Bu;
      const float k0 = i0 * ks, k1 = i1 * ks;
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t la, ha, lb, hb;
        hash_quad((alo + (uint32_t)(4 * np + t)) ^ ahi, dr.key, la, ha);
        hash_quad((blo + (uint32_t)(4 * np + t)) ^ bhi, dr.key, lb, hb);
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
      dp[nt][0] *= pdm[nt][0]; dp[nt][1] *= pdm[nt][1]; dp[nt][2] *= pdm[nt][2]; dp[nt][3] *= pdm[nt][3];
      d0 += dp[nt][0] + dp[nt][1];
      d1 += dp[nt][2] + dp[nt][3];
    }
    d0 += __shfl_xor_sync(0xffffffffu, d0, 1); d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 1); d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
    const float e0 = -(d0 * rc1) * i0, e1 = -(d1 * rc1) * i1;
    // exchange layout: word ((mt * 4 + nt) * 4 + j) * 32 + lane, j = 0, 1: dropped-P rows g / g + 8 ; j = 2, 3: dS rows g / g + 8
    uint4 *xo = reinterpret_cast<uint4 *>(xch) + (mt * 4) * 32 + lane;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
      xo[nt * 32] = make_uint4(pack_bf16(pdm[nt][0], pdm[nt][1]), pack_bf16(pdm[nt][2], pdm[nt][3]),
                               pack_bf16(fmaf(p[nt][0], e0, dp[nt][0]), fmaf(p[nt][1], e0, dp[nt][1])),
                               pack_bf16(fmaf(p[nt][2], e1, dp[nt][2]), fmaf(p[nt][3], e1, dp[nt][3])));
  }
  named_bar_sync(1, T256_CTHREADS);                     // every warp has finished phase A: the images may be overwritten block by block
  uint32_t pdp[2][4][2], dsq[2][4][2];                  // [query m-tile][key n-tile][rows g / g + 8]
  {
    const uint4 *xi = reinterpret_cast<const uint4 *>(xch) + lane;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const uint4 v = xi[(mt * 4 + nt) * 32];
        pdp[mt][nt][0] = v.x; pdp[mt][nt][1] = v.y; dsq[mt][nt][0] = v.z; dsq[mt][nt][1] = v.w;
      }
  }
  // ---- phase B: 16 feature columns at a time ----
#pragma unroll 1
  for (int bi = 0; bi < 2; ++bi) {
    const int blk = wq + 4 * bi;
    uint8_t *im = blk >= 4 ? imgO : imgE;
    const int cq = (blk & 3) * 16;                      // column of the block inside its group: q at cq, k at 64 + cq, v at 128 + cq, dO at 192 + cq
    float dq[2][2][4], dk[2][2][4], dv[2][2][4];        // [m-tile][n-tile of the block][4]
#pragma unroll
    for (int a2 = 0; a2 < 2; ++a2)
#pragma unroll
      for (int b2 = 0; b2 < 2; ++b2)
#pragma unroll
        for (int c = 0; c < 4; ++c) { dq[a2][b2][c] = 0.f; dk[a2][b2][c] = 0.f; dv[a2][b2][c] = 0.f; }
#pragma unroll
    for (int kt = 0; kt < 2; ++kt) {                    // contraction index 16 kt .. 16 kt + 15: keys for dq, queries for dk / dv
      const int rowk = s * 32 + 16 * kt + (mi & 1) * 8 + rr;
      uint32_t bk[4], bq[4], bo[4];
      ldmatrix_x4_trans(bk, im + kmajor_off(rowk, 64 + cq + (mi >> 1) * 8, 128));
      ldmatrix_x4_trans(bq, im + kmajor_off(rowk, cq + (mi >> 1) * 8, 128));
      ldmatrix_x4_trans(bo, im + kmajor_off(rowk, 192 + cq + (mi >> 1) * 8, 128));
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {                  // dq rows = queries of m-tile mt, contraction over keys 16 kt ..
        mma16816(dq[mt][0], dsq[mt][2 * kt][0], dsq[mt][2 * kt][1], dsq[mt][2 * kt + 1][0], dsq[mt][2 * kt + 1][1], bk[0], bk[1]);
        mma16816(dq[mt][1], dsq[mt][2 * kt][0], dsq[mt][2 * kt][1], dsq[mt][2 * kt + 1][0], dsq[mt][2 * kt + 1][1], bk[2], bk[3]);
      }
#pragma unroll
      for (int kmt = 0; kmt < 2; ++kmt) {               // dk / dv rows = keys of m-tile kmt, contraction over the queries of m-tile kt
        const uint32_t s0t = movmatrix_trans(dsq[kt][2 * kmt][0]), s1t = movmatrix_trans(dsq[kt][2 * kmt + 1][0]);
        const uint32_t s2t = movmatrix_trans(dsq[kt][2 * kmt][1]), s3t = movmatrix_trans(dsq[kt][2 * kmt + 1][1]);
        mma16816(dk[kmt][0], s0t, s1t, s2t, s3t, bq[0], bq[1]);
        mma16816(dk[kmt][1], s0t, s1t, s2t, s3t, bq[2], bq[3]);
        const uint32_t p0t = movmatrix_trans(pdp[kt][2 * kmt][0]), p1t = movmatrix_trans(pdp[kt][2 * kmt + 1][0]);
        const uint32_t p2t = movmatrix_trans(pdp[kt][2 * kmt][1]), p3t = movmatrix_trans(pdp[kt][2 * kmt + 1][1]);
        mma16816(dv[kmt][0], p0t, p1t, p2t, p3t, bo[0], bo[1]);
        mma16816(dv[kmt][1], p0t, p1t, p2t, p3t, bo[2], bo[3]);
      }
    }
    __syncwarp();                                       // all lanes have read this block's q / k / dO columns
    float *gb = g_b + (blk >= 4 ? 192 : 0);
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      float sq0 = 0.f, sq1 = 0.f, sk0 = 0.f, sk1 = 0.f, sv0 = 0.f, sv1 = 0.f;
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int r0 = s * 32 + 16 * mt + g, col = cq + 8 * nt + 2 * t;
        const float k0 = dk[mt][nt][0] * dk_scale, k1 = dk[mt][nt][1] * dk_scale, k2 = dk[mt][nt][2] * dk_scale, k3 = dk[mt][nt][3] * dk_scale;
        const float v0 = dv[mt][nt][0] * rc1, v1 = dv[mt][nt][1] * rc1, v2 = dv[mt][nt][2] * rc1, v3 = dv[mt][nt][3] * rc1;
        *reinterpret_cast<uint32_t *>(im + kmajor_off(r0, col, 128)) = pack_bf16(dq[mt][nt][0], dq[mt][nt][1]);
        *reinterpret_cast<uint32_t *>(im + kmajor_off(r0 + 8, col, 128)) = pack_bf16(dq[mt][nt][2], dq[mt][nt][3]);
        *reinterpret_cast<uint32_t *>(im + kmajor_off(r0, 64 + col, 128)) = pack_bf16(k0, k1);
        *reinterpret_cast<uint32_t *>(im + kmajor_off(r0 + 8, 64 + col, 128)) = pack_bf16(k2, k3);
        *reinterpret_cast<uint32_t *>(im + kmajor_off(r0, 128 + col, 128)) = pack_bf16(v0, v1);
        *reinterpret_cast<uint32_t *>(im + kmajor_off(r0 + 8, 128 + col, 128)) = pack_bf16(v2, v3);
        sq0 += dq[mt][nt][0] + dq[mt][nt][2]; sq1 += dq[mt][nt][1] + dq[mt][nt][3];
        sk0 += k0 + k2; sk1 += k1 + k3; sv0 += v0 + v2; sv1 += v1 + v3;
      }
      // in-projection bias gradient: column sums over the pair's 32 rows (lanes with equal t hold the same columns)
#pragma unroll
      for (int o = 4; o <= 16; o <<= 1) {
        sq0 += __shfl_xor_sync(0xffffffffu, sq0, o); sq1 += __shfl_xor_sync(0xffffffffu, sq1, o);
        sk0 += __shfl_xor_sync(0xffffffffu, sk0, o); sk1 += __shfl_xor_sync(0xffffffffu, sk1, o);
        sv0 += __shfl_xor_sync(0xffffffffu, sv0, o); sv1 += __shfl_xor_sync(0xffffffffu, sv1, o);
      }
      if (g == 0) {
        const int col = cq + 8 * nt + 2 * t;
        atomicAdd(gb + col, sq0); atomicAdd(gb + col + 1, sq1);
        atomicAdd(gb + 64 + col, sk0); atomicAdd(gb + 64 + col + 1, sk1);
        atomicAdd(gb + 128 + col, sv0); atomicAdd(gb + 128 + col + 1, sv1);
      }
    }
  }
}

// ---- LayerNorm backward over a [128 x 256] tile; thread = (row, 64-column part) ----------------------------------
// dyv(cb, out16): loads 16 values of the incoming gradient for columns part*64 + cb..  (re-readable)
// Writes du (fp32 -> TMEM accumulator t_du) and da = du * dropmask (bf16 -> sImg and gImg); accumulates dgamma / dbeta / dbias partials.
// The row's mean / rstd come from the forward (`stat`: float2 per token row), so there is no statistics pass over u; both passes
// walk the thread's 64 columns in chunks of 16 with the NEXT chunk's loads issued before the current chunk's arithmetic (the phase
// is bound by the latency of its global loads — 4 warps per scheduler, every warp in the same phase — not by their bytes).
template <class DyLoad>
__device__ __forceinline__ void t256_ln_bwd(DyLoad dyv, const uint8_t *u_img /*global tile image*/, const float2 *stat /*this row's (mean, rstd)*/,
                                            const float *gamma, uint32_t t_du /*TMEM: this thread's row, column 0*/,
                                            uint32_t t_dy /*TMEM stash for dy (256 columns), or 0xFFFFFFFF: dyv is cheap to call again*/,
                                            uint32_t t_u /*TMEM stash for the packed u words: 8 columns at the head of every 16-column chunk*/,
                                            uint8_t *sImg, uint8_t *gImg, const Drop &dr, uint64_t e_row /* element index of (row, col 0) */,
                                            float *g_gamma, float *g_beta, float *g_bias, float *sStatA, float *sStatB, int row, int part, int lane,
                                            unsigned long long *dbg, int &ndbg) {
#define LN_STAMP() do { if (dbg != nullptr && ndbg < 60) dbg[ndbg++] = clock64(); } while (0)
  const float2 mr = __ldg(stat);
  const float rs = mr.y, nmr = -mr.x * mr.y;
  const uint8_t *u_row = u_img + kmajor_off(row, part * 64, 128);      // 8 columns further = + 2048 bytes
  // pass B: m1 = mean(dy g), m2 = mean(dy g xhat) ; dgamma / dbeta column sums
  float m1 = 0.f, m2 = 0.f;
  const ColsumSel csel = colsum_sel(lane);
  const int cs_col = (lane >> 2) + 8 * (lane & 1);    // lanes with (lane & 3) < 2 flush column g (t = 0) / g + 8 (t = 1) of a 16-column sum
  const bool cs_on = (lane & 3) < 2;
  float dyn[16];
  uint4 un0, un1;
  dyv(0, dyn);
  un0 = *reinterpret_cast<const uint4 *>(u_row); un1 = *reinterpret_cast<const uint4 *>(u_row + 2048);
#pragma unroll 1
  for (int cb = 0; cb < 64; cb += 16) {
    float dy[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) dy[j] = dyn[j];
    const uint4 v0 = un0, v1 = un1;
    if (cb + 16 < 64) {
      dyv(cb + 16, dyn);
      un0 = *reinterpret_cast<const uint4 *>(u_row + (cb + 16) * 256); un1 = *reinterpret_cast<const uint4 *>(u_row + (cb + 16) * 256 + 2048);
    }
    {
      // pass C takes this chunk's dy and u back from tensor memory instead of from L2 (both phases are bound by global-load
      // latency and SM <-> L2 bytes; a TMEM round trip costs neither)
      const uint32_t uw[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
      tmem_st8(t_u + (uint32_t)(part * 64 + cb), uw);
      if (t_dy != 0xFFFFFFFFu) tmem_st16(t_dy + (uint32_t)(part * 64 + cb), dy);
    }
    const float uu[16] = {bf16lo(v0.x), bf16hi(v0.x), bf16lo(v0.y), bf16hi(v0.y), bf16lo(v0.z), bf16hi(v0.z), bf16lo(v0.w), bf16hi(v0.w),
                          bf16lo(v1.x), bf16hi(v1.x), bf16lo(v1.y), bf16hi(v1.y), bf16lo(v1.z), bf16hi(v1.z), bf16lo(v1.w), bf16hi(v1.w)};
    uint32_t pg[8], pb[8];                            // bf16 pairs of dy * xhat (-> dgamma) and dy (-> dbeta)
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
      const float4 gm = *reinterpret_cast<const float4 *>(gamma + part * 64 + cb + (j & ~3));
      const float ga = (j & 2) ? gm.z : gm.x, gb = (j & 2) ? gm.w : gm.y;
      const float xa = fmaf(uu[j], rs, nmr), xb = fmaf(uu[j + 1], rs, nmr);
      const float gda = dy[j] * ga, gdb = dy[j + 1] * gb;
      m1 += gda + gdb; m2 = fmaf(gda, xa, m2); m2 = fmaf(gdb, xb, m2);
      pg[j >> 1] = pack_bf16(dy[j] * xa, dy[j + 1] * xb);
      pb[j >> 1] = pack_bf16(dy[j], dy[j + 1]);
    }
    float sg0, sg1, sb0, sb1;
    warp_colsum16_packed(pg, csel, sg0, sg1);
    warp_colsum16_packed(pb, csel, sb0, sb1);
    if (cs_on) {
      atomicAdd(g_gamma + part * 64 + cb + cs_col, (lane & 1) ? sg1 : sg0);
      atomicAdd(g_beta + part * 64 + cb + cs_col, (lane & 1) ? sb1 : sb0);
    }
  }
  LN_STAMP();
  sStatA[row * 4 + part] = m1; sStatB[row * 4 + part] = m2;
  tmem_st_wait();                                    // the stashes of pass B are in tensor memory
  named_bar_sync(1, T256_CTHREADS);
  LN_STAMP();
  {
    const float4 sa = *reinterpret_cast<const float4 *>(sStatA + row * 4), sb = *reinterpret_cast<const float4 *>(sStatB + row * 4);
    m1 = ((sa.x + sa.y) + (sa.z + sa.w)) * (1.f / 256);
    m2 = ((sb.x + sb.y) + (sb.z + sb.w)) * (1.f / 256);
  }
  // pass C: du = rstd (dy g - m1 - xhat m2) ; da = du * dropmask
  const uint64_t w0 = (e_row + (uint64_t)(part * 64)) >> 2;
  const uint32_t wlo = (uint32_t)w0, xhi = (uint32_t)(w0 >> 32) * 0x85EBCA6Bu;
  const float ks = dr.thr ? dr.scale : 1.f;
  const uint32_t thr2 = dr.thr | (dr.thr << 16);
#pragma unroll 1
  for (int cb = 0; cb < 64; cb += 16) {
    float w[16];
    {
      float dy[16];
      uint32_t uw[8];
      tmem_ld8(t_u + (uint32_t)(part * 64 + cb), reinterpret_cast<float *>(uw));
      if (t_dy != 0xFFFFFFFFu) { tmem_ld16(t_dy + (uint32_t)(part * 64 + cb), dy); tmem_ld_wait(); }
      else dyv(cb, dy);                               // (waits for its own TMEM load, which also completes the one above)
      const float uu[16] = {bf16lo(uw[0]), bf16hi(uw[0]), bf16lo(uw[1]), bf16hi(uw[1]), bf16lo(uw[2]), bf16hi(uw[2]), bf16lo(uw[3]), bf16hi(uw[3]),
                            bf16lo(uw[4]), bf16hi(uw[4]), bf16lo(uw[5]), bf16hi(uw[5]), bf16lo(uw[6]), bf16hi(uw[6]), bf16lo(uw[7]), bf16hi(uw[7])};
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const float4 gm = *reinterpret_cast<const float4 *>(gamma + part * 64 + cb + j);
        const float gj[4] = {gm.x, gm.y, gm.z, gm.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float xh = fmaf(uu[j + e], rs, nmr);
          w[j + e] = rs * (dy[j + e] * gj[e] - m1 - xh * m2);
        }
      }
    }
    // du (the gradient that by-passes the sub-layer through the residual) goes straight into the TMEM accumulator the next UMMAs
    // add their dx onto — no fp32 tile parked in L2 and re-read (it was 640 KB of the ~3 MB a tile moved between SM and L2)
    tmem_st16(t_du + (uint32_t)(part * 64 + cb), w);
    // da = dropout(du): scale in fp32, round to bf16 pairs, AND the packed keep masks (two decisions per half-precision compare)
    uint32_t pk[8];
#pragma unroll
    for (int j = 0; j < 16; j += 2) pk[j >> 1] = pack_bf16(w[j] * ks, w[j + 1] * ks);
    if (dr.thr) {
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        uint32_t lo, hi;
        hash_quad((wlo + (uint32_t)((cb + j) >> 2)) ^ xhi, dr.key, lo, hi);
        pk[j >> 1] &= keep2(lo, thr2);
        pk[(j >> 1) + 1] &= keep2(hi, thr2);
      }
    }
#pragma unroll
    for (int j = 0; j < 16; j += 8) {
      const uint4 pv = make_uint4(pk[j >> 1], pk[(j >> 1) + 1], pk[(j >> 1) + 2], pk[(j >> 1) + 3]);
      const uint32_t off = kmajor_off(row, part * 64 + cb + j, 128);
      *reinterpret_cast<uint4 *>(sImg + off) = pv;
      *reinterpret_cast<uint4 *>(gImg + off) = pv;
    }
    float s0, s1;
    warp_colsum16_packed(pk, csel, s0, s1);
    if (cs_on) atomicAdd(g_bias + part * 64 + cb + cs_col, (lane & 1) ? s1 : s0);
  }
}

// =============================================================================================
// data-gradient kernel
// =============================================================================================
template <int DH, bool DEVSTEP = false>
__global__ void __launch_bounds__(T256_THREADS, 1) t256_layer_bwd_kernel(const T256Args a_in) {
  const DropArgsView<T256Args, DEVSTEP> view(a_in);
  const T256Args &a = view.a;
  constexpr int G = T256_G, GH = DH >= 64 ? 1 : 64 / DH, NS = T256_NS;
  constexpr bool H128 = DH == 128;                   // a head spans two groups: q | k | v come from the forward's saved images, no recompute
  using S = T256BwdSmem;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_full[NS], bar_empty[NS], bar_da2ready, bar_hready[2], bar_hfree[2], bar_dhfull[2], bar_dhimgready[2], bar_r2a,
      bar_r2b, bar_dx1full, bar_da1ready, bar_dctxfull, bar_xready, bar_qkvfree, bar_qkvfull, bar_dqkvready, bar_dqkvfree, bar_dxinfull, bar_qkvstaged;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int F = a.F, NCH = F / 64, H = a.H;
  uint8_t *sR1 = smem + S::r1, *sR2 = smem + S::r2, *sRing = smem + S::ring;
  float *sPar = reinterpret_cast<float *>(smem + S::par);
  float *p_bqkv = sPar, *p_g1 = sPar + 768, *p_g2 = p_g1 + 256;
  float *sG = reinterpret_cast<float *>(smem + S::gpar);
  float *g_bqkv = sG, *g_bo = sG + 768, *g_b2 = g_bo + 256, *g_g1 = g_b2 + 256, *g_be1 = g_g1 + 256, *g_g2 = g_be1 + 256, *g_be2 = g_g2 + 256,
        *g_b1 = g_be2 + 256;
  float *sStatA = reinterpret_cast<float *>(smem + S::stat), *sStatB = sStatA + 512;

  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 1); }
    uint64_t *bars[] = {&bar_da2ready, &bar_hready[0], &bar_hready[1], &bar_hfree[0], &bar_hfree[1], &bar_dhfull[0], &bar_dhfull[1],
                        &bar_dhimgready[0], &bar_dhimgready[1], &bar_r2a, &bar_r2b, &bar_dx1full, &bar_da1ready,
                        &bar_dctxfull, &bar_xready, &bar_qkvfree, &bar_qkvfull, &bar_dqkvready, &bar_dqkvfree, &bar_dxinfull};
    for (uint64_t *b : bars) mbar_init(b, 1);
    mbar_init(&bar_qkvstaged, 2);                     // head_dim 128: one arrive.expect_tx per saved q | k | v group image of a head
    fence_mbar_init();
  }
  for (int i = tid; i < 768; i += T256_THREADS) p_bqkv[i] = a.bqkv[i];
  if (tid < 256) { p_g1[tid] = a.g1[tid]; p_g2[tid] = a.g2[tid]; }
  for (int i = tid; i < 2816; i += T256_THREADS) sG[i] = 0.f;
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const uint32_t t_dx = tmem, t_b = tmem + 256;
  const uint32_t aR1 = smem_u32(sR1), aR2 = smem_u32(sR2), aRing = smem_u32(sRing);
  if (a.stagger) {
    const long long until = clock64() + (long long)(blockIdx.x & 3) * (long long)a.stagger;
    while (clock64() < until) __nanosleep(200);
  }
  const int nfs = 4 * NCH;                          // FFN stages per tile
  const uint32_t uses_per_tile = H128 ? (uint32_t)(8 + NCH) : (uint32_t)(16 + NCH);   // stages per tile / NS: (64 + 4 NCH) / 4, head_dim 128: (32 + 4 NCH) / 4
  const uint8_t *wimg = a.img + (size_t)(blockIdx.x % T256_REP) * a.img_rep_stride + (size_t)t256_fwd_stages(F) * T256_STAGE;

  if (warp == 16) {
    // ======================= TMA producer =======================
    if (elect_one()) {
      uint32_t it = 0;
      const uint64_t pol_w = l2_policy_evict_last(), pol_a = l2_policy_evict_first();
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
        const uint32_t use0 = it * uses_per_tile;
        auto stage = [&](int st, uint32_t bytes) {
          const int slot = st % NS;
          mbar_wait(&bar_empty[slot], ((use0 + (uint32_t)(st / NS)) & 1u) ^ 1u);
          mbar_expect_tx(&bar_full[slot], bytes);
          tma_load_1d_hint(sRing + slot * T256_STAGE, wimg + (size_t)st * T256_STAGE, bytes, &bar_full[slot], pol_w);
        };
        mbar_wait(&bar_r2b, (it & 1u) ^ 1u);                   // the previous tile's q|k|v recompute no longer reads the x image in r2
        for (int c = 0; c < NCH; ++c) {
          // FFN chunks alternate between two (H chunk, dH image) slots of r2: chunk n uses slot n & 1 for the (n >> 1)-th time
          const uint32_t n = it * (uint32_t)NCH + (uint32_t)c, fs = n & 1u, fu = (n >> 1) & 1u;
          mbar_wait(&bar_hfree[fs], fu ^ 1u);
          mbar_expect_tx(&bar_hready[fs], 16384u);
          tma_load_1d_hint(sR2 + fs * 32768u, a.h_img + ((size_t)tile * NCH + c) * 16384, 16384u, &bar_hready[fs], pol_a);
#pragma unroll 1
          for (int j = 0; j < 4; ++j) stage(4 * c + j, 16384u);
        }
        if constexpr (H128) {
          // no x image and no q | k | v recompute stages: Wo^T x8, then WqkvT(g) x6 for the four groups
#pragma unroll 1
          for (int j = 0; j < 32; ++j) stage(nfs + j, 16384u);
          continue;
        }
        mbar_wait(&bar_r2a, it & 1u);                          // dx1 of the last FFN chunk retired: r2 is free for the x image
        mbar_expect_tx(&bar_xready, 65536u);
        tma_load_1d_hint(sR2, a.x_img_in + (size_t)tile * T256_TILE_IMG, 32768u, &bar_xready, pol_a);
        tma_load_1d_hint(sR2 + 32768, a.x_img_in + (size_t)tile * T256_TILE_IMG + 32768, 32768u, &bar_xready, pol_a);
#pragma unroll 1
        for (int j = 0; j < 8; ++j) stage(nfs + j, 16384u);
#pragma unroll 1
        for (int j = 0; j < 56; ++j) {
          // QKV(0) x8 ; g = 1..3: QKV(g) x8, WqkvT(g-1) x6 ; WqkvT(3) x6
          const bool is_qkv = j < 8 || (j < 50 && ((j - 8) % 14) < 8);
          stage(nfs + 8 + j, is_qkv ? 12288u : 16384u);
        }
      }
    }
  } else if (warp == 17) {
    // ======================= MMA issuer =======================
    if (elect_one()) {
      const uint32_t id_192 = make_idesc_bf16(128, 192), id_256 = make_idesc_bf16(128, 256), id_64 = make_idesc_bf16(128, 64);
      const uint64_t dR1 = descA128(aR1), dR2 = descA128(aR2);
      const uint64_t dB192 = descB(aRing, 192), dB256 = descB(aRing, 256), dB64 = descB(aRing, 64);
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
        const uint32_t use0 = it * uses_per_tile;
        auto full_wait = [&](int st) {
          mbar_wait(&bar_full[st % NS], (use0 + (uint32_t)(st / NS)) & 1u);
          fence_after_sync();
        };
        // ---- FFN backward ----
        mbar_wait(&bar_da2ready, it & 1u);
        fence_after_sync();
        // dH(c + 1) is issued BEFORE the issuer waits for the dH image of chunk c, into the other accumulator / image slot, so the
        // hidden-gradient epilogue of chunk c overlaps the tensor work of its neighbours.  The ring is consumed out of stream
        // order (stages 4c+4, 4c+5 before 4c+2, 4c+3) but every slot still sees its stages in order, and four slots = one chunk.
        auto issue_dh = [&](int c) {                    // dH(c) = da2 W2[:, chunk]
          const uint32_t n = it * (uint32_t)NCH + (uint32_t)c, fs = n & 1u;
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const int st = 4 * c + hf;
            full_wait(st);
            const uint64_t db = desc_adv(dB64, (uint32_t)(st % NS) * T256_STAGE);
#pragma unroll
            for (int k = 0; k < 8; ++k)
              mma_bf16_ss(t_b + fs * 64u, desc_adv(dR1, (uint32_t)(hf * 8 + k) * 4096u), desc_adv(db, (uint32_t)k * 2048u), id_64, (hf | k) > 0);
            mma_commit(&bar_empty[st % NS]);
          }
          mma_commit(&bar_dhfull[fs]);
        };
        issue_dh(0);
        for (int c = 0; c < NCH; ++c) {
          const uint32_t n = it * (uint32_t)NCH + (uint32_t)c, fs = n & 1u, fu = (n >> 1) & 1u;
          if (c + 1 < NCH) issue_dh(c + 1);
          mbar_wait(&bar_dhimgready[fs], fu);
          fence_after_sync();
          const uint64_t dDH = descA128(aR2 + fs * 32768u + 16384u);
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {              // dx1 += dH(c) W1[chunk, :]
            const int st = 4 * c + 2 + hf;
            full_wait(st);
            const uint64_t db = desc_adv(dB256, (uint32_t)(st % NS) * T256_STAGE);
#pragma unroll
            for (int k = 0; k < 2; ++k)
              mma_bf16_ss(t_dx, desc_adv(dDH, (uint32_t)(hf * 2 + k) * 4096u), desc_adv(db, (uint32_t)k * 8192u), id_256, 1);     // onto du2
            mma_commit(&bar_empty[st % NS]);
          }
        }
        mma_commit(&bar_dx1full);
        mma_commit(&bar_r2a);
        // ---- dctx = da1 Wo ----
        mbar_wait(&bar_da1ready, it & 1u);
        fence_after_sync();
#pragma unroll 2
        for (int b = 0; b < 8; ++b) {
          const int st = nfs + b;
          full_wait(st);
          const uint64_t db = desc_adv(dB256, (uint32_t)(st % NS) * T256_STAGE);
#pragma unroll
          for (int k = 0; k < 2; ++k)
            mma_bf16_ss(t_b, desc_adv(dR1, (uint32_t)(b * 2 + k) * 4096u), desc_adv(db, (uint32_t)k * 8192u), id_256, (b | k) > 0);
          mma_commit(&bar_empty[st % NS]);
        }
        mma_commit(&bar_dctxfull);
        if constexpr (H128) {
          // ---- attention backward, per HEAD: the dq | dk | dv images of its two groups sit in r1 (even) and r2 (odd) ----
          for (int h = 0; h < 2; ++h) {
            mbar_wait(&bar_dqkvready, (it * 2u + (uint32_t)h) & 1u);
            fence_after_sync();
            for (int e = 0; e < 2; ++e) {
              const int gg = 2 * h + e;
              const uint64_t dA = e ? dR2 : dR1;
#pragma unroll 2
              for (int b = 0; b < 6; ++b) {
                const int st = nfs + 8 + gg * 6 + b;
                full_wait(st);
                const uint64_t db = desc_adv(dB256, (uint32_t)(st % NS) * T256_STAGE);
#pragma unroll
                for (int k = 0; k < 2; ++k)
                  mma_bf16_ss(t_dx, desc_adv(dA, (uint32_t)(b * 2 + k) * 4096u), desc_adv(db, (uint32_t)k * 8192u), id_256, 1);     // onto du1
                mma_commit(&bar_empty[st % NS]);
              }
            }
            if (h == 0) mma_commit(&bar_dqkvfree);        // (head 1's images are released by bar_dxinfull: a commit nobody waits for is a synccheck finding)
          }
          mma_commit(&bar_dxinfull);                      // (bar_r2b is raised by the compute warps at the end of the tile: the bulk stores of dq | dk | dv read r2 too)
          continue;
        }
        // ---- attention backward, per head group ----
        mbar_wait(&bar_xready, it & 1u);
        fence_after_sync();
        const int A0 = nfs + 8;
        auto dxin = [&](int gg, int st0) {              // dx_in += dqkv[:, group gg] Wqkv[group gg rows, :]   (K = 192: six stages)
          const uint64_t dA = gg == G - 1 ? dR2 : dR1;   // the last group's dq | dk | dv image sits in r2
          const uint32_t n = it * G + (uint32_t)gg;
          mbar_wait(&bar_dqkvready, n & 1u);
          fence_after_sync();
#pragma unroll 2
          for (int b = 0; b < 6; ++b) {
            const int st = st0 + b;
            full_wait(st);
            const uint64_t db = desc_adv(dB256, (uint32_t)(st % NS) * T256_STAGE);
#pragma unroll
            for (int k = 0; k < 2; ++k)
              mma_bf16_ss(t_dx, desc_adv(dA, (uint32_t)(b * 2 + k) * 4096u), desc_adv(db, (uint32_t)k * 8192u), id_256, 1);     // onto du1
            mma_commit(&bar_empty[st % NS]);
          }
          if (gg < G - 2) mma_commit(&bar_dqkvfree);      // groups G - 2 and G - 1 release their images through bar_dxinfull (nobody waits on this barrier for them)
        };
#pragma unroll 1
        for (int g = 0; g < G; ++g) {
          const int st0 = g == 0 ? A0 : A0 + 8 + (g - 1) * 14;
          const uint32_t n = it * G + (uint32_t)g;
          mbar_wait(&bar_qkvfree, n & 1u);
          fence_after_sync();
#pragma unroll 2
          for (int kc = 0; kc < 8; ++kc) {
            const int st = st0 + kc;
            full_wait(st);
            const uint64_t db = desc_adv(dB192, (uint32_t)(st % NS) * T256_STAGE);
#pragma unroll
            for (int k = 0; k < 2; ++k)
              mma_bf16_ss(t_b, desc_adv(dR2, (uint32_t)(kc * 2 + k) * 4096u), desc_adv(db, (uint32_t)k * 6144u), id_192, (kc | k) > 0);
            mma_commit(&bar_empty[st % NS]);
          }
          mma_commit(&bar_qkvfull);
          if (g >= 1) dxin(g - 1, st0 + 8);
        }
        dxin(G - 1, A0 + 50);
        mma_commit(&bar_dxinfull);
      }
    }
  } else {
    // ======================= compute warps =======================
    const int q4 = warp & 3, part = warp >> 2;
    const int row = q4 * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q4 * 32) << 16;
    const float attn_scale = rsqrtf((float)DH) * 1.4426950408889634f;
    const ColsumSel csel = colsum_sel(lane);
    const uint64_t pol_stage = l2_policy_evict_first();
    uint8_t *scratch = a.dctx_scratch + (size_t)blockIdx.x * T256_TILE_IMG;
    // du2 / du1 (LayerNorm-input gradients, the residual by-pass) are written into the dx accumulator t_dx with tcgen05.st and the
    // following UMMAs accumulate onto them; earlier versions parked them in a per-CTA fp32 tile in L2 and re-added them
    uint32_t it = 0;
    int ndbg = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
      const bool dbg_on = a.dbg != nullptr && blockIdx.x == 0 && tid == 0 && it == 1;
      T256_STAMP();
      const int64_t grow = (int64_t)tile * TC_TILE + row;
      const float *dy_t = a.dy + (size_t)tile * T256_TILE_F32;
      float *dx_t = a.dx + (size_t)tile * T256_TILE_F32;
      const uint64_t e_row = (uint64_t)((a.seq0 * 32 + grow) * 256);
      // ---- B0: LayerNorm2 backward ----
      {
        auto dyload = [&](int cb, float (&o)[16]) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(dy_t + ((size_t)((part * 64 + cb + j) >> 2) * 128 + row) * 4));
            o[j] = v.x; o[j + 1] = v.y; o[j + 2] = v.z; o[j + 3] = v.w;
          }
        };
        t256_ln_bwd(dyload, a.u2_img + (size_t)tile * T256_TILE_IMG, a.ln2_stat + grow, p_g2, t_dx + lane_off, t_b + lane_off, t_dx + lane_off, sR1, a.da2_img + (size_t)tile * T256_TILE_IMG, a.d2, e_row,
                    g_g2, g_be2, g_b2, sStatA, sStatB, row, part, lane, dbg_on ? a.dbg : nullptr, ndbg);
      }
      tmem_st_wait();                                 // du2 sits in t_dx: dx1 accumulates onto it
      fence_async_smem();
      fence_before_sync();
      named_bar_sync(1, T256_CTHREADS);
      if (tid == 0) {
        mbar_arrive(&bar_da2ready);
        // u1 of this tile is first touched by LayerNorm1 backward, one FFN phase from now: pull it into L2 while the HBM is quiet
        prefetch_l2_bulk(a.u1_img + (size_t)tile * T256_TILE_IMG, 32768u);
        prefetch_l2_bulk(a.u1_img + (size_t)tile * T256_TILE_IMG + 32768, 32768u);
      }
      T256_STAMP();
      // ---- B2: hidden-activation gradient per FFN chunk ----
      for (int c = 0; c < NCH; ++c) {
        const uint32_t n = it * (uint32_t)NCH + (uint32_t)c, fs = n & 1u, fu = (n >> 1) & 1u;
        uint8_t *sH = sR2 + fs * 32768u;                // this chunk's slot: H image at +0, dH image at +16384
        mbar_wait(&bar_hready[fs], fu);
        mbar_wait(&bar_dhfull[fs], fu);
        fence_after_sync();
        float w[16];
        tmem_ld16(t_b + fs * 64u + lane_off + (uint32_t)(part * 16), w);
        tmem_ld_wait();
        const uint4 h0 = *reinterpret_cast<const uint4 *>(sH + kmajor_off(row, part * 16, 128));
        const uint4 h1 = *reinterpret_cast<const uint4 *>(sH + kmajor_off(row, part * 16 + 8, 128));
        const uint32_t hh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
        const float sc = a.d_ffn.scale;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          w[2 * j] = (hh[j] & 0x7FFFu) != 0u && !(hh[j] & 0x8000u) ? w[2 * j] * sc : 0.f;                  // H > 0  (bf16 low half)
          w[2 * j + 1] = (hh[j] & 0x7FFF0000u) != 0u && !(hh[j] & 0x80000000u) ? w[2 * j + 1] * sc : 0.f;  // bf16 high half
        }
        const uint4 pk0 = make_uint4(pack_bf16(w[0], w[1]), pack_bf16(w[2], w[3]), pack_bf16(w[4], w[5]), pack_bf16(w[6], w[7]));
        const uint4 pk1 = make_uint4(pack_bf16(w[8], w[9]), pack_bf16(w[10], w[11]), pack_bf16(w[12], w[13]), pack_bf16(w[14], w[15]));
        uint8_t *dhg = a.dh_img + ((size_t)tile * NCH + c) * 16384;
        const uint32_t o0 = kmajor_off(row, part * 16, 128), o1 = kmajor_off(row, part * 16 + 8, 128);
        *reinterpret_cast<uint4 *>(sH + 16384 + o0) = pk0; *reinterpret_cast<uint4 *>(sH + 16384 + o1) = pk1;
        *reinterpret_cast<uint4 *>(dhg + o0) = pk0; *reinterpret_cast<uint4 *>(dhg + o1) = pk1;
        {
          const uint32_t ph[8] = {pk0.x, pk0.y, pk0.z, pk0.w, pk1.x, pk1.y, pk1.z, pk1.w};
          float s0, s1;
          warp_colsum16_packed(ph, csel, s0, s1);       // linear1 bias gradient: column sums of the dH image rows of this warp
          if ((lane & 3) < 2) atomicAdd(g_b1 + c * 64 + part * 16 + (lane >> 2) + 8 * (lane & 1), (lane & 1) ? s1 : s0);
        }
        fence_async_smem();
        fence_before_sync();
        named_bar_sync(1, T256_CTHREADS);
        if (tid == 0) { mbar_arrive(&bar_dhimgready[fs]); mbar_arrive(&bar_hfree[fs]); }
      }
      T256_STAMP();
      // ---- B4: LayerNorm1 backward on du2 + dx1 (TMEM accumulator) ----
      mbar_wait(&bar_dx1full, it & 1u);
      fence_after_sync();
      T256_STAMP();
      // head_dim 128: the saved q | k | v group images (48 KB each) are staged by bulk TMA, one thread issuing, as early as their
      // target is free: head 0's odd group -> r2 here (the FFN is done with r2), its even group -> r1 once dctx has read the da1 image
      auto stage_qkv = [&](int gg, uint8_t *dst) {
        const uint8_t *qg = a.qkv_img + ((size_t)tile * G + gg) * T256_QKV_GROUP_IMG;
        mbar_expect_tx(&bar_qkvstaged, (uint32_t)T256_QKV_GROUP_IMG);
        tma_load_1d_hint(dst, qg, 32768u, &bar_qkvstaged, pol_stage);
        tma_load_1d_hint(dst + 32768, qg + 32768, 16384u, &bar_qkvstaged, pol_stage);
      };
      if constexpr (H128) { if (tid == 0) stage_qkv(1, sR2); }
      // dy / u2 of this CTA's NEXT tile (first touched by its LayerNorm2 backward) are pulled into L2 during the LAST attention unit of
      // this tile, when the weight stream has nothing but the final dx_in stages left (at the start of the attention phases the same
      // prefetch competes with the stream and loses: profiles/r02b)
      auto prefetch_next = [&]() {
        const int nt = tile + (int)gridDim.x;
        if (nt >= a.n_tiles) return;
        const uint8_t *dyp = reinterpret_cast<const uint8_t *>(a.dy + (size_t)nt * T256_TILE_F32);
#pragma unroll 1
        for (int q = 0; q < 4; ++q) prefetch_l2_bulk(dyp + q * 32768, 32768u);
        prefetch_l2_bulk(a.u2_img + (size_t)nt * T256_TILE_IMG, 32768u);
        prefetch_l2_bulk(a.u2_img + (size_t)nt * T256_TILE_IMG + 32768, 32768u);
      };
      {
        auto dyload = [&](int cb, float (&o)[16]) {     // du2 + dx1: the accumulator was seeded with du2
          tmem_ld16(t_dx + lane_off + (uint32_t)(part * 64 + cb), o);
          tmem_ld_wait();
        };
        t256_ln_bwd(dyload, a.u1_img + (size_t)tile * T256_TILE_IMG, a.ln1_stat + grow, p_g1, t_dx + lane_off, 0xFFFFFFFFu, t_b + lane_off, sR1, a.da1_img + (size_t)tile * T256_TILE_IMG, a.d1, e_row,
                    g_g1, g_be1, g_bo, sStatA, sStatB, row, part, lane, dbg_on ? a.dbg : nullptr, ndbg);
      }
      tmem_st_wait();                                 // du1 sits in t_dx: dx_in accumulates onto it
      fence_async_smem();
      fence_before_sync();
      named_bar_sync(1, T256_CTHREADS);
      if (tid == 0) mbar_arrive(&bar_da1ready);
      T256_STAMP();
      // ---- B5: dctx (TMEM) -> bf16 -> this CTA's scratch image ----
      mbar_wait(&bar_dctxfull, it & 1u);
      fence_after_sync();
      T256_STAMP();
      if constexpr (H128) { if (tid == 0) stage_qkv(0, sR1); }
#pragma unroll
      for (int cb = 0; cb < 64; cb += 16) {
        float f[16];
        tmem_ld16(t_b + lane_off + (uint32_t)(part * 64 + cb), f);
        tmem_ld_wait();
        *reinterpret_cast<uint4 *>(scratch + kmajor_off(row, part * 64 + cb, 128)) =
            make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
        *reinterpret_cast<uint4 *>(scratch + kmajor_off(row, part * 64 + cb + 8, 128)) =
            make_uint4(pack_bf16(f[8], f[9]), pack_bf16(f[10], f[11]), pack_bf16(f[12], f[13]), pack_bf16(f[14], f[15]));
      }
      __threadfence_block();
      fence_before_sync();
      named_bar_sync(1, T256_CTHREADS);
      if (tid == 0 && !H128) mbar_arrive(&bar_qkvfree);
      T256_STAMP();
      if constexpr (H128) {
        // ---- B6 (head_dim 128): per head, stage q | k | v (saved by the forward) and dO (from this CTA's dctx scratch) of its two
        // groups into r1 (even group) and r2 (odd group), attention backward in place, hand both dq | dk | dv images to the issuer ----
        for (int h = 0; h < 2; ++h) {
          if (h >= 1) {
            mbar_wait(&bar_dqkvfree, it & 1u);                    // dx_in of head 0 no longer reads r1 / r2
            if (tid == 0) tma_store_wait_read();                   // ... nor do the bulk stores of its dq | dk | dv slices
            named_bar_sync(1, T256_CTHREADS);
            if (tid == 0) { stage_qkv(2, sR1); stage_qkv(3, sR2); }
          }
#pragma unroll
          for (int e = 0; e < 2; ++e) {                            // dO of the head's two groups: from this CTA's dctx scratch (generic copies)
            const int gg = 2 * h + e;
            uint8_t *dst = e ? sR2 : sR1;
            *reinterpret_cast<uint4 *>(dst + 49152 + (size_t)tid * 16) = __ldcg(reinterpret_cast<const uint4 *>(scratch + (size_t)gg * 16384 + (size_t)tid * 16));
            *reinterpret_cast<uint4 *>(dst + 49152 + 8192 + (size_t)tid * 16) = __ldcg(reinterpret_cast<const uint4 *>(scratch + (size_t)gg * 16384 + 8192 + (size_t)tid * 16));
          }
          mbar_wait(&bar_qkvstaged, (it * 2u + (uint32_t)h) & 1u);  // both q | k | v images of the head have landed
          named_bar_sync(1, T256_CTHREADS);
          if (h == 1 && tid == 0) prefetch_next();
          T256_STAMP();
          {
            const int s = warp & 3;
            const int64_t seq = a.seq0 + (int64_t)tile * 4 + s;
            const uint64_t w_pair = (uint64_t)((seq * H + h) * 32) * 8u;
            t256_attn_bwd128(sR1, sR2, reinterpret_cast<uint32_t *>(smem + S::stat) + s * 1024, s, warp >> 2, lane, a.d_attn, w_pair,
                             g_bqkv + (2 * h) * 192);
          }
          fence_async_smem();
          named_bar_sync(1, T256_CTHREADS);
          if (tid == 0) {
            mbar_arrive(&bar_dqkvready);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const uint8_t *src = e ? sR2 : sR1;
              uint8_t *dq_g = a.dqkv_img + (size_t)tile * (3 * T256_TILE_IMG) + (size_t)(2 * h + e) * 16384;
              tma_store_1d(dq_g, src, 16384u);
              tma_store_1d(dq_g + T256_TILE_IMG, src + 16384, 16384u);
              tma_store_1d(dq_g + 2 * T256_TILE_IMG, src + 32768, 16384u);
            }
            tma_store_commit();
          }
          T256_STAMP();
        }
      } else
      // ---- B6: head groups ----
      for (int g = 0; g < G; ++g) {
        const uint32_t n = it * G + (uint32_t)g;
        mbar_wait(&bar_qkvfull, n & 1u);
        fence_after_sync();
        T256_STAMP();
        float v[48];
#pragma unroll
        for (int i = 0; i < 3; ++i) tmem_ld16(t_b + lane_off + (uint32_t)(part * 48 + i * 16), v + i * 16);
        const uint4 dc0 = __ldcg(reinterpret_cast<const uint4 *>(scratch + (size_t)g * 16384 + (size_t)tid * 16));
        const uint4 dc1 = __ldcg(reinterpret_cast<const uint4 *>(scratch + (size_t)g * 16384 + 8192 + (size_t)tid * 16));
        tmem_ld_wait();
        // The LAST group takes r2 as its scratch image: the x image there is dead once its q|k|v recompute has retired (bar_qkvfull
        // above), so its staging does not wait for dx_in of group G - 2 to release r1 (4 K clocks of a 121 K tile at C4).
        uint8_t *sS = g == G - 1 ? sR2 : sR1;
        if (g >= 1 && g < G - 1) mbar_wait(&bar_dqkvfree, (it * (uint32_t)(G - 2) + (uint32_t)(g - 1)) & 1u);     // dx_in of the previous group no longer reads the dqkv image in r1
        if (tid == 0 && g < G - 1) tma_store_wait_read();                       // ... nor do the bulk stores of its dq | dk | dv slices
        fence_before_sync();
        named_bar_sync(1, T256_CTHREADS);
        if (tid == 0 && g + 1 < G) mbar_arrive(&bar_qkvfree);     // the TMEM chunk is drained: the next group's q|k|v may overwrite it
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const int n0 = part * 48 + i * 8, pq = n0 >> 6;
          const float *bias = p_bqkv + pq * 256 + g * 64 + (n0 & 63);
          const float sc = pq == 0 ? attn_scale : 1.f;
          float w[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) w[j] = (v[i * 8 + j] + bias[j]) * sc;
          *reinterpret_cast<uint4 *>(sS + kmajor_off(row, n0, 128)) =
              make_uint4(pack_bf16(w[0], w[1]), pack_bf16(w[2], w[3]), pack_bf16(w[4], w[5]), pack_bf16(w[6], w[7]));
        }
        *reinterpret_cast<uint4 *>(sS + 49152 + (size_t)tid * 16) = dc0;
        *reinterpret_cast<uint4 *>(sS + 49152 + 8192 + (size_t)tid * 16) = dc1;
        named_bar_sync(1, T256_CTHREADS);
        if (g == G - 1 && tid == 0) prefetch_next();
        T256_STAMP();
        if (warp < 4 * GH) {
          const int s = warp / GH, hl = warp % GH;
          const int64_t seq = a.seq0 + (int64_t)tile * 4 + s;
          const uint64_t w_pair = (uint64_t)((seq * H + (g * GH + hl)) * 32) * 8u;     // quad index of (row 0, position 0)
          if constexpr (!H128) t256_attn_bwd<DH>(sS, s, hl, lane, a.d_attn, w_pair, g_bqkv + g * 192);
        }
        fence_async_smem();
        named_bar_sync(1, T256_CTHREADS);
        if (tid == 0) {
          mbar_arrive(&bar_dqkvready);
          uint8_t *dq_g = a.dqkv_img + (size_t)tile * (3 * T256_TILE_IMG) + (size_t)g * 16384;
          tma_store_1d(dq_g, sS, 16384u);
          tma_store_1d(dq_g + T256_TILE_IMG, sS + 16384, 16384u);
          tma_store_1d(dq_g + 2 * T256_TILE_IMG, sS + 32768, 16384u);
          tma_store_commit();
        }
        T256_STAMP();
      }
      // ---- B7: dx = du1 + dx_in (one accumulator) ----
      mbar_wait(&bar_dxinfull, it & 1u);
      fence_after_sync();
      T256_STAMP();
#pragma unroll
      for (int cb = 0; cb < 64; cb += 16) {
        float f[16];
        tmem_ld16(t_dx + lane_off + (uint32_t)(part * 64 + cb), f);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<float4 *>(dx_t + ((size_t)((part * 64 + cb + 4 * j) >> 2) * 128 + row) * 4) =
              make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
      }
      if (tid == 0) {
        tma_store_wait_read();
        mbar_arrive(&bar_r2b);                          // dx_in has retired (bar_dxinfull) and the dq | dk | dv bulk stores no longer read r2
      }
      fence_before_sync();
      named_bar_sync(1, T256_CTHREADS);
      T256_STAMP();
    }
    if (tid == 0) tma_store_wait_all();
    // ---- flush bias / LayerNorm gradient partials ----
    named_bar_sync(1, T256_CTHREADS);
    for (int i = tid; i < 768; i += T256_CTHREADS) {
      // g_bqkv is stored per group: [g][3][64] -> in_proj_bias index part*256 + g*64 + col
      const int g = i / 192, r = i % 192, pq = r / 64, col = r % 64;
      atomicAdd(a.gbqkv + pq * 256 + g * 64 + col, g_bqkv[i]);
    }
    for (int i = tid; i < F; i += T256_CTHREADS) atomicAdd(a.gb1 + i, g_b1[i]);
    if (tid < 256) {
      atomicAdd(a.gbo + tid, g_bo[tid]); atomicAdd(a.gb2 + tid, g_b2[tid]);
      atomicAdd(a.gg1 + tid, g_g1[tid]); atomicAdd(a.gbe1 + tid, g_be1[tid]);
      atomicAdd(a.gg2 + tid, g_g2[tid]); atomicAdd(a.gbe2 + tid, g_be2[tid]);
    }
  }
  __syncwarp();
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int DH>
static int t256_launch_bwd(const T256Args &a_in, int grid, cudaStream_t st) {
  static T256Dbg dbg;
  // Persistent CTAs with equal work run in lockstep, so their HBM-heavy phases (LayerNorm2 backward pulls dy + u2 while the dx
  // rows of the previous tile drain) coincide and the HBM idles during the attention phases.  Starting CTA b (b % 4) * S clocks late
  // spreads them out (GT_T256_STAGGER=S: C4 backward 20.3 -> 19.7 ms at S = 20000 before the next-tile L2 prefetch existed);
  // with that prefetch (issued during the last attention unit) the two are equivalent (18.9 ms either way), so the default is off.
  static const uint32_t stagger = getenv("GT_T256_STAGGER") ? (uint32_t)atoi(getenv("GT_T256_STAGGER")) : 0u;
  T256Args a = a_in;
  a.stagger = a.n_tiles >= 16 * grid ? stagger : 0u;
  const bool d = dbg.arm(a, st);
  if (drop_args_devstep(a)) {
    GT_CUDA(cudaFuncSetAttribute(t256_layer_bwd_kernel<DH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T256BwdSmem::total));
    { LaunchScope _ls(KC_TC_LAYER_BWD, st);
      t256_layer_bwd_kernel<DH, true><<<grid, T256_THREADS, T256BwdSmem::total, st>>>(a); }
    GT_CUDA(cudaGetLastError());
    return 0;
  }
  GT_CUDA(cudaFuncSetAttribute(t256_layer_bwd_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T256BwdSmem::total));
  { LaunchScope _ls(KC_TC_LAYER_BWD, st);
    t256_layer_bwd_kernel<DH><<<grid, T256_THREADS, T256BwdSmem::total, st>>>(a); }
  GT_CUDA(cudaGetLastError());
  if (d) dbg.report("t256 bwd", st);
  return 0;
}

int t256_layer_bwd(const T256Args &a, cudaStream_t st) {
  GT_CHECK(a.F % 64 == 0 && a.F >= 64 && a.F <= 512, "t256_layer_bwd: dim_feedforward not supported");
  int grid = a.n_tiles < t256_num_sms() ? a.n_tiles : t256_num_sms();
  if (grid > 160) grid = 160;                        // dctx scratch is sized for 160 CTAs
  switch (a.dh) {
    case 16: return t256_launch_bwd<16>(a, grid, st);
    case 32: return t256_launch_bwd<32>(a, grid, st);
    case 128:
      GT_CHECK(a.qkv_img != nullptr, "t256_layer_bwd: head_dim 128 needs the q | k | v images saved by the forward");
      return t256_launch_bwd<128>(a, grid, st);
    default: GT_FAIL("t256_layer_bwd: head dim not instantiated");
  }
}

// =============================================================================================
// weight-gradient kernel: dW = sum over tiles of  A_tile^T B_tile  with both operands taken as MN-major views
// of the saved [128 tokens x C] images.  CTA = (job, split): job = TWO adjacent [128 x N] blocks of one weight gradient
// (256 output rows), split = a contiguous range of tiles.  The kernel is bound by the bytes it pulls from L2 (~43 B / clock /
// SM with every SM streaming), so the B image of a tile (up to 64 KB) is fetched once for both 128-row blocks: 128 KB per
// 8.4 M MACs instead of 96 KB per 4.2 M.  Six 32 KB slots form a ring of sub-stages (see the kernel); the accumulators
// (2 blocks x [128 x N] fp32) live in TMEM for the CTA's whole tile range and are added to the fp32 gradient with vector
// atomics at the end.
// =============================================================================================
struct T256WJob {
  const uint8_t *a_img, *b_img;       // per-tile images
  uint32_t a_tile_stride, a_off;      // bytes: tile stride of the A image, offset of this job's first 128-column block
  uint32_t b_tile_stride, b_bytes;    // B: whole image of the tile ([128 x N])
  int N;                              // 16..256
  int na;                             // 128-column blocks of A taken by this job: always 2 (adjacent)
  float *out;                         // out[(m) * ld_m + n * ld_n], m = 0 .. 255
  int ld_m, ld_n;
  int tile0, tile1;
};
constexpr int T256_WG_MAXJOBS = 160;
struct T256WgradLaunch {
  T256WJob jobs[T256_WG_MAXJOBS];
};
static_assert(sizeof(T256WJob) * T256_WG_MAXJOBS <= T256_WG_JOBBUF, "weight-gradient job list exceeds its device buffer");

__device__ __forceinline__ uint64_t t256_desc_mn(uint32_t base, uint32_t bytes) {      // image [128 rows (k) x cols (mn)]; k16 step = 256 B
  return make_desc(base + bytes, 128u, 2048u);
}

constexpr uint32_t T256_WG_SLOT = 32768;
constexpr int T256_WG_SLOTS = 6;

// Sub-stages of a tile, 32 KB each, in ring order: A0 (first 128-column block of A), B0 (first <= 128 columns of B), A1, B1 (the
// rest of B when N > 128).  Four groups of 8 UMMAs [128 x N/nbh x 16] contract them pairwise — (A0,B0) (A1,B0) (A0,B1) (A1,B1) —
// and each sub-stage is handed back to the producer right after the last group that reads it has been issued, so 64 - 96 KB of
// loads are in flight all the time (the kernel is bound by L2 -> SM bytes: bytes in flight / latency is its throughput).
__global__ void __launch_bounds__(192, 1) t256_wgrad_kernel(const T256WJob *__restrict__ jobs) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_full[T256_WG_SLOTS], bar_empty[T256_WG_SLOTS], bar_done;
  __shared__ uint32_t tmem_slot;
  const T256WJob job = jobs[blockIdx.x];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    for (int i = 0; i < T256_WG_SLOTS; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 1); }
    mbar_init(&bar_done, 1);
    fence_mbar_init();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const int ntl = job.tile1 - job.tile0;
  const int nbh = job.b_bytes > 32768u ? 2 : 1;              // halves of B
  const int nh = job.N / nbh;                                // columns per half
  const uint32_t bh_bytes = job.b_bytes / (uint32_t)nbh;
  const uint32_t ns = 2u + (uint32_t)nbh;                    // sub-stages per tile (job.na == 2)
  if (warp == 4) {
    if (elect_one()) {
      const uint64_t pol = l2_policy_evict_first();        // every image byte is read once per job: do not let it displace the other kernels' weights
      for (int i = 0; i < ntl; ++i) {
        const size_t tile = (size_t)(job.tile0 + i);
        const uint8_t *ag = job.a_img + tile * job.a_tile_stride + job.a_off, *bg = job.b_img + tile * job.b_tile_stride;
        for (uint32_t j = 0; j < ns; ++j) {                  // A0, B0, A1, B1
          const uint32_t q = (uint32_t)i * ns + j, slot = q % T256_WG_SLOTS, use = q / T256_WG_SLOTS;
          const bool is_a = (j & 1u) == 0u;
          const uint8_t *src = is_a ? ag + (j >> 1) * 32768u : bg + (j >> 1) * bh_bytes;
          const uint32_t bytes = is_a ? 32768u : bh_bytes;
          mbar_wait(&bar_empty[slot], (use & 1u) ^ 1u);
          mbar_expect_tx(&bar_full[slot], bytes);
          tma_load_1d_hint(smem + slot * T256_WG_SLOT, src, bytes, &bar_full[slot], pol);
        }
      }
    }
  } else if (warp == 5) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16(128, nh, 1, 1);
      const uint32_t base = smem_u32(smem);
      for (int i = 0; i < ntl; ++i) {
        uint32_t sl[4];
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) sl[j] = ((uint32_t)i * ns + j) % T256_WG_SLOTS;
        auto wait_full = [&](uint32_t j) {
          mbar_wait(&bar_full[sl[j]], ((((uint32_t)i * ns + j) / T256_WG_SLOTS) & 1u));
          fence_after_sync();
        };
        auto group = [&](uint32_t ja, uint32_t jb, uint32_t acc) {
          const uint32_t pa = base + sl[ja] * T256_WG_SLOT, pb = base + sl[jb] * T256_WG_SLOT;
#pragma unroll
          for (int k = 0; k < 8; ++k)
            mma_bf16_ss(tmem + acc, t256_desc_mn(pa, (uint32_t)k * 256u), t256_desc_mn(pb, (uint32_t)k * 256u), idesc, (i | k) > 0);
        };
        wait_full(0); wait_full(1);
        group(0, 1, 0u);                                     // A0 x B0
        wait_full(2);
        group(2, 1, 256u);                                   // A1 x B0
        mma_commit(&bar_empty[sl[1]]);
        if (nbh == 2) {
          wait_full(3);
          group(0, 3, 128u);                                 // A0 x B1
          mma_commit(&bar_empty[sl[0]]);
          group(2, 3, 256u + 128u);                          // A1 x B1
          mma_commit(&bar_empty[sl[2]]);
          mma_commit(&bar_empty[sl[3]]);
        } else {
          mma_commit(&bar_empty[sl[0]]);
          mma_commit(&bar_empty[sl[2]]);
        }
      }
      mma_commit(&bar_done);
    }
  }
  if (warp < 4) {
    if (ntl > 0) {
      mbar_wait(&bar_done, 0);
      fence_after_sync();
      const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
      for (int ab = 0; ab < 2; ++ab) {
        const int m = ab * 128 + warp * 32 + lane;
        for (int cb = 0; cb < job.N; cb += 16) {
          float f[16];
          // accumulator of (block ab, half h) at column 256 ab + 128 h; inside a half the columns are consecutive
          tmem_ld16(tmem + (uint32_t)ab * 256u + (uint32_t)(cb / nh) * 128u + (uint32_t)(cb % nh) + lane_off, f);
          tmem_ld_wait();
          if (job.ld_n == 1) {
            float *o = job.out + (size_t)m * job.ld_m + cb;
#pragma unroll
            for (int j = 0; j < 16; j += 4) atomicAdd(reinterpret_cast<float4 *>(o + j), make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) atomicAdd(job.out + (size_t)m * job.ld_m + (size_t)(cb + j) * job.ld_n, f[j]);
          }
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

static int t256_wgrad_launch(T256WgradLaunch &L, int nj, void *job_buf, cudaStream_t st) {
  GT_CUDA(cudaMemcpyAsync(job_buf, L.jobs, sizeof(T256WJob) * nj, cudaMemcpyHostToDevice, st));
  GT_CUDA(cudaFuncSetAttribute(t256_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(T256_WG_SLOTS * T256_WG_SLOT)));
  { LaunchScope _ls(KC_TC_WGRAD, st);
    t256_wgrad_kernel<<<nj, 192, T256_WG_SLOTS * T256_WG_SLOT, st>>>(reinterpret_cast<const T256WJob *>(job_buf)); }
  GT_CUDA(cudaGetLastError());
  return 0;
}

int t256_wgrad(const T256WgradArgs &a, void *job_buf, cudaStream_t st) {
  // job list (each job = 256 output rows = two 128-column blocks of the A image): dWqkv 3 (A = dqkv, B = x), dWo 1 (A = da1, B = ctx),
  // per block of <= 256 hidden units: dW1^T 1 (A = x1, B = dH), dW2 1 (A = da2, B = H)
  static thread_local T256WgradLaunch L;
  const int F = a.F, nt = a.n_tiles;
  const uint32_t hbytes = (uint32_t)(128 * F * 2);
  struct Base { const uint8_t *ai; uint32_t as, ao; const uint8_t *bi; uint32_t bs, bb; int N; float *out; int ldm, ldn; };
  Base base[16];
  int nb = 0;
  for (int p = 0; p < 3; ++p)
    base[nb++] = {a.dqkv_img, (uint32_t)(3 * T256_TILE_IMG), (uint32_t)p * 65536u, a.x_img, (uint32_t)T256_TILE_IMG, 65536u, 256,
                  a.gwqkv + (size_t)p * 256 * 256, 256, 1};
  base[nb++] = {a.da1_img, (uint32_t)T256_TILE_IMG, 0u, a.ctx_img, (uint32_t)T256_TILE_IMG, 65536u, 256, a.gwo, 256, 1};
  const int nfb = (F + 255) / 256;                   // N blocks of at most 256 hidden units
  for (int fb = 0; fb < nfb; ++fb) {
    const int Nf = F - fb * 256 < 256 ? F - fb * 256 : 256;
    // dW1[f][j] = sum_t dH[t][f] x1[t][j]  computed transposed: acc[m = j][n = f]
    base[nb++] = {a.x1_img, (uint32_t)T256_TILE_IMG, 0u, a.dh_img + (size_t)fb * 65536, hbytes, (uint32_t)(Nf * 256), Nf,
                  a.gw1 + (size_t)fb * 256 * 256, 1, 256};
    // dW2[j][f] = sum_t da2[t][j] H[t][f]
    base[nb++] = {a.da2_img, (uint32_t)T256_TILE_IMG, 0u, a.h_img + (size_t)fb * 65536, hbytes, (uint32_t)(Nf * 256), Nf,
                  a.gw2 + (size_t)fb * 256, F, 1};
  }
  // CTAs per job in proportion to the bytes a tile of the job pulls from L2 (that is what bounds the kernel)
  int64_t wsum = 0;
  for (int i = 0; i < nb; ++i) wsum += 65536 + base[i].bb;
  const int sms = t256_num_sms();
  int nj = 0;
  for (int i = 0; i < nb; ++i) {
    int splits = (int)((int64_t)sms * (65536 + base[i].bb) / wsum);
    if (splits < 1) splits = 1;
    if (splits > nt) splits = nt;
    for (int s = 0; s < splits && nj < T256_WG_MAXJOBS; ++s) {
      T256WJob &j = L.jobs[nj++];
      j.a_img = base[i].ai; j.a_tile_stride = base[i].as; j.a_off = base[i].ao;
      j.b_img = base[i].bi; j.b_tile_stride = base[i].bs; j.b_bytes = base[i].bb; j.N = base[i].N; j.na = 2;
      j.out = base[i].out; j.ld_m = base[i].ldm; j.ld_n = base[i].ldn;
      j.tile0 = (int)((int64_t)nt * s / splits); j.tile1 = (int)((int64_t)nt * (s + 1) / splits);
    }
  }
  return t256_wgrad_launch(L, nj, job_buf, st);
}

// out[256][N] (row stride N) += sum over tiles of A_tile^T B_tile: A = [128 x 256] images (tile stride 64 KB), B = [128 x N] images
// (tile stride 128 * N * 2 bytes), N in {16, 32, ..., 256}.  Used by edge256.cu for the parameter gradients of the input layer and
// of the final LayerNorm + head.
int t256_wgrad_pair(const uint8_t *a_img, const uint8_t *b_img, int N, float *out, int n_tiles, void *job_buf, cudaStream_t st) {
  static thread_local T256WgradLaunch L;
  GT_CHECK(N >= 16 && N <= 256 && N % 16 == 0, "t256_wgrad_pair: N must be a multiple of 16 in [16, 256]");
  int splits = t256_num_sms();
  if (splits > T256_WG_MAXJOBS) splits = T256_WG_MAXJOBS;
  if (splits > n_tiles) splits = n_tiles;
  if (splits < 1) splits = 1;
  int nj = 0;
  for (int s = 0; s < splits; ++s) {
    T256WJob &j = L.jobs[nj++];
    j.a_img = a_img; j.a_tile_stride = (uint32_t)T256_TILE_IMG; j.a_off = 0u;
    j.b_img = b_img; j.b_tile_stride = (uint32_t)(128 * N * 2); j.b_bytes = (uint32_t)(128 * N * 2); j.N = N; j.na = 2;
    j.out = out; j.ld_m = N; j.ld_n = 1;
    j.tile0 = (int)((int64_t)n_tiles * s / splits); j.tile1 = (int)((int64_t)n_tiles * (s + 1) / splits);
  }
  return t256_wgrad_launch(L, nj, job_buf, st);
}

}  // namespace gt
