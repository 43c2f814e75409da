// attn_mma.cu — attention forward / backward of the PER-OP bf16 path (shapes without fused layer kernels: C3's head dim 128,
// sweep draws with d_model 64 / 128 / 512) on warp-level mma.sync, head dims 16 / 32 / 64 / 128.
// torch/nn/functional.py:6682-6690 (scaled_dot_product_attention: softmax(Q K^T / sqrt(dh)) with dropout on the probabilities, P V).
//
// One warp owns one (sequence, head) pair: 32 queries x 32 keys.  q | k | v (and dO) are read straight from the row-major fp32
// activations as the mma fragments need them — every fragment load of a warp touches whole 32-byte sectors (8 rows x 32 B for the
// row-operand pattern, 4 rows x 32 B for the column-operand pattern), so nothing is staged in shared memory (the fp32 SIMT kernel
// staged 49 KB per warp at head dim 128 and ran at 4 warps per SM) — and rounded to bf16 on the way into the registers.
//   S = (Q qs) K^T, dP = dO V^T      m16n8k16, contraction over the head features in steps of 16
//   O = P V, dQ = dS K               m16n8k16, contraction over the 32 keys
//   dK = dS^T Q, dV = P^T dO         m16n8k16, A = movmatrix-transposed fragments, contraction over the 32 queries
// Dropout: common.cuh hash_quad / key_perm — the four probabilities a thread owns for one query row in n-tiles {0,1} (or {2,3})
// are one quad, exactly like the fused kernels, the fp32 kernels and the oracle.
#include "common.cuh"
#include "umma.cuh"

namespace gt {
using namespace umma;

__device__ __forceinline__ uint32_t am_pack2(const float *p, float s) {       // two consecutive floats of one row
  const float2 v = __ldg(reinterpret_cast<const float2 *>(p));
  return pack_bf16(v.x * s, v.y * s);
}
__device__ __forceinline__ uint32_t am_pack_col(const float *p, int64_t ld) {  // the same column of two consecutive rows
  return pack_bf16(__ldg(p), __ldg(p + ld));
}

// S (or dP) accumulators of both query m-tiles: acc[u][nt][4] += A[32 x DH] B[32 x DH]^T, A scaled by `as`
template <int DH>
__device__ __forceinline__ void am_scores(const float *A, int64_t lda, const float *B, int64_t ldb, float as, int g, int t, float (&acc)[2][4][4]) {
#pragma unroll 2
  for (int ks = 0; ks < DH / 16; ++ks) {
    const int c0 = 16 * ks + 2 * t;
    uint32_t a[2][4];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const float *r0 = A + (int64_t)(16 * u + g) * lda + c0, *r1 = r0 + 8 * lda;
      a[u][0] = am_pack2(r0, as); a[u][1] = am_pack2(r1, as); a[u][2] = am_pack2(r0 + 8, as); a[u][3] = am_pack2(r1 + 8, as);
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const float *kr = B + (int64_t)(8 * nt + g) * ldb + c0;
      const uint32_t b0 = am_pack2(kr, 1.f), b1 = am_pack2(kr + 8, 1.f);
      mma16816(acc[0][nt], a[0][0], a[0][1], a[0][2], a[0][3], b0, b1);
      mma16816(acc[1][nt], a[1][0], a[1][1], a[1][2], a[1][3], b0, b1);
    }
  }
}

// keep bits of the 16 probabilities this lane owns in query m-tile u: bit (4 nt + c)
__device__ __forceinline__ uint32_t am_keep_bits(const Drop &dr, uint64_t w_pair, int u, int g, int t) {
  if (!dr.thr) return 0xFFFFu;
  uint32_t keep = 0;
  const uint64_t wa = w_pair + (uint64_t)(16 * u + g) * 8u, wb = wa + 64u;     // rows q0 and q0 + 8
#pragma unroll
  for (int np = 0; np < 2; ++np) {
    uint32_t la, ha, lb, hb;
    const uint64_t qa = wa + (uint64_t)(4 * np + t), qb = wb + (uint64_t)(4 * np + t);
    hash_quad((uint32_t)qa ^ ((uint32_t)(qa >> 32) * 0x85EBCA6Bu), dr.key, la, ha);
    hash_quad((uint32_t)qb ^ ((uint32_t)(qb >> 32) * 0x85EBCA6Bu), dr.key, lb, hb);
    keep |= ((la & 0xFFFFu) >= dr.thr ? 1u : 0u) << (8 * np);
    keep |= ((la >> 16) >= dr.thr ? 1u : 0u) << (8 * np + 1);
    keep |= ((lb & 0xFFFFu) >= dr.thr ? 1u : 0u) << (8 * np + 2);
    keep |= ((lb >> 16) >= dr.thr ? 1u : 0u) << (8 * np + 3);
    keep |= ((ha & 0xFFFFu) >= dr.thr ? 1u : 0u) << (8 * np + 4);
    keep |= ((ha >> 16) >= dr.thr ? 1u : 0u) << (8 * np + 5);
    keep |= ((hb & 0xFFFFu) >= dr.thr ? 1u : 0u) << (8 * np + 6);
    keep |= ((hb >> 16) >= dr.thr ? 1u : 0u) << (8 * np + 7);
  }
  return keep;
}

// in place: s -> softmax probabilities (rows g and g + 8 of m-tile u live in c = 0,1 and c = 2,3); scores are in log2 units
__device__ __forceinline__ void am_softmax(float (&s)[4][4], bool causal, int u, int g, int t) {
  if (causal) {
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (8 * nt + 2 * t + (c & 1) > 16 * u + g + (c >> 1) * 8) s[nt][c] = -1e30f;
  }
  float m0 = s[0][0], m1 = s[0][2];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) { m0 = fmaxf(m0, fmaxf(s[nt][0], s[nt][1])); m1 = fmaxf(m1, fmaxf(s[nt][2], s[nt][3])); }
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    s[nt][0] = ex2_ftz(s[nt][0] - m0); s[nt][1] = ex2_ftz(s[nt][1] - m0);
    s[nt][2] = ex2_ftz(s[nt][2] - m1); s[nt][3] = ex2_ftz(s[nt][3] - m1);
    s0 += s[nt][0] + s[nt][1]; s1 += s[nt][2] + s[nt][3];
  }
  s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
  s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
  const float i0 = 1.f / s0, i1 = 1.f / s1;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) { s[nt][0] *= i0; s[nt][1] *= i0; s[nt][2] *= i1; s[nt][3] *= i1; }
}

constexpr int AM_WARPS = 4;

template <int DH>
__global__ void __launch_bounds__(AM_WARPS * 32) attn_mma_fwd_kernel(const AttnArgs a) {
  Drop adrop = a.drop;
  drop_resolve(adrop);                               // graph replay: key from the device step counter
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int64_t pair = (int64_t)blockIdx.x * AM_WARPS + warp;
  if (pair >= a.n_seq * a.H) return;
  const int64_t seq = pair / a.H;
  const int head = (int)(pair % a.H);
  const float *Q = a.q + seq * T * a.ldq + head * DH, *K = a.k + seq * T * a.ldk + head * DH, *V = a.v + seq * T * a.ldv + head * DH;
  float *O = a.o + seq * T * a.ldo + head * DH;
  float p[2][4][4];
#pragma unroll
  for (int u = 0; u < 2; ++u)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int c = 0; c < 4; ++c) p[u][nt][c] = 0.f;
  am_scores<DH>(Q, a.ldq, K, a.ldk, rsqrtf((float)DH) * 1.4426950408889634f, g, t, p);
  const uint64_t w_pair = (uint64_t)((((a.seq0 + seq) * a.H + head) * T) * 8);
  uint32_t pa[2][2][4];                              // dropped probabilities as A fragments: [m-tile][key k-tile][4]
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    am_softmax(p[u], a.causal != 0, u, g, t);
    const uint32_t keep = am_keep_bits(adrop, w_pair, u, g, t);
    const float ks = adrop.scale;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int c = 0; c < 4; ++c) p[u][nt][c] = ((keep >> (4 * nt + c)) & 1u) ? p[u][nt][c] * ks : 0.f;
#pragma unroll
    for (int kt = 0; kt < 2; ++kt) {
      pa[u][kt][0] = pack_bf16(p[u][2 * kt][0], p[u][2 * kt][1]); pa[u][kt][1] = pack_bf16(p[u][2 * kt][2], p[u][2 * kt][3]);
      pa[u][kt][2] = pack_bf16(p[u][2 * kt + 1][0], p[u][2 * kt + 1][1]); pa[u][kt][3] = pack_bf16(p[u][2 * kt + 1][2], p[u][2 * kt + 1][3]);
    }
  }
#pragma unroll 2
  for (int nb = 0; nb < DH / 8; ++nb) {              // 8 output features per step
    float o[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
    for (int kt = 0; kt < 2; ++kt) {
      const float *vc = V + (int64_t)(16 * kt + 2 * t) * a.ldv + 8 * nb + g;
      const uint32_t b0 = am_pack_col(vc, a.ldv), b1 = am_pack_col(vc + 8 * a.ldv, a.ldv);
      mma16816(o[0], pa[0][kt][0], pa[0][kt][1], pa[0][kt][2], pa[0][kt][3], b0, b1);
      mma16816(o[1], pa[1][kt][0], pa[1][kt][1], pa[1][kt][2], pa[1][kt][3], b0, b1);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      float *r0 = O + (int64_t)(16 * u + g) * a.ldo + 8 * nb + 2 * t;
      *reinterpret_cast<float2 *>(r0) = make_float2(o[u][0], o[u][1]);
      *reinterpret_cast<float2 *>(r0 + 8 * a.ldo) = make_float2(o[u][2], o[u][3]);
    }
  }
}

template <int DH>
__global__ void __launch_bounds__(AM_WARPS * 32) attn_mma_bwd_kernel(const AttnArgs a) {
  Drop adrop = a.drop;
  drop_resolve(adrop);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int64_t pair = (int64_t)blockIdx.x * AM_WARPS + warp;
  if (pair >= a.n_seq * a.H) return;
  const int64_t seq = pair / a.H;
  const int head = (int)(pair % a.H);
  const float *Q = a.q + seq * T * a.ldq + head * DH, *K = a.k + seq * T * a.ldk + head * DH, *V = a.v + seq * T * a.ldv + head * DH;
  const float *dO = a.d_o + seq * T * a.ld_do + head * DH;
  float *dQ = a.dq + seq * T * a.ld_dq + head * DH, *dK = a.dk + seq * T * a.ld_dk + head * DH, *dV = a.dv + seq * T * a.ld_dv + head * DH;
  const float inv_sqrt_dh = rsqrtf((float)DH);
  uint32_t pd[2][4][2], ds[2][4][2];                 // [query m-tile][key n-tile][rows g / g + 8]: dropped P ; dS / sqrt(dh)
  {
    float p[2][4][4], dp[2][4][4];
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int c = 0; c < 4; ++c) { p[u][nt][c] = 0.f; dp[u][nt][c] = 0.f; }
    am_scores<DH>(Q, a.ldq, K, a.ldk, inv_sqrt_dh * 1.4426950408889634f, g, t, p);
    am_scores<DH>(dO, a.ld_do, V, a.ldv, 1.f, g, t, dp);
    const uint64_t w_pair = (uint64_t)((((a.seq0 + seq) * a.H + head) * T) * 8);
    const float ks = adrop.scale;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      am_softmax(p[u], a.causal != 0, u, g, t);
      const uint32_t keep = am_keep_bits(adrop, w_pair, u, g, t);
      float d0 = 0.f, d1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
        for (int c = 0; c < 4; ++c) dp[u][nt][c] = ((keep >> (4 * nt + c)) & 1u) ? dp[u][nt][c] * ks : 0.f;      // dL/dP through the dropout
        d0 += dp[u][nt][0] * p[u][nt][0] + dp[u][nt][1] * p[u][nt][1];
        d1 += dp[u][nt][2] * p[u][nt][2] + dp[u][nt][3] * p[u][nt][3];
      }
      d0 += __shfl_xor_sync(0xffffffffu, d0, 1); d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
      d1 += __shfl_xor_sync(0xffffffffu, d1, 1); d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        float s4[4], q4[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          s4[c] = p[u][nt][c] * (dp[u][nt][c] - (c < 2 ? d0 : d1)) * inv_sqrt_dh;
          q4[c] = ((keep >> (4 * nt + c)) & 1u) ? p[u][nt][c] * ks : 0.f;
        }
        ds[u][nt][0] = pack_bf16(s4[0], s4[1]); ds[u][nt][1] = pack_bf16(s4[2], s4[3]);
        pd[u][nt][0] = pack_bf16(q4[0], q4[1]); pd[u][nt][1] = pack_bf16(q4[2], q4[3]);
      }
    }
  }
  // transposed fragments for dK = dS^T Q and dV = Pd^T dO: [key m-tile][query k-tile][4]
  uint32_t dst[2][2][4], pdt[2][2][4];
#pragma unroll
  for (int kmt = 0; kmt < 2; ++kmt)
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      dst[kmt][mt][0] = movmatrix_trans(ds[mt][2 * kmt][0]); dst[kmt][mt][1] = movmatrix_trans(ds[mt][2 * kmt + 1][0]);
      dst[kmt][mt][2] = movmatrix_trans(ds[mt][2 * kmt][1]); dst[kmt][mt][3] = movmatrix_trans(ds[mt][2 * kmt + 1][1]);
      pdt[kmt][mt][0] = movmatrix_trans(pd[mt][2 * kmt][0]); pdt[kmt][mt][1] = movmatrix_trans(pd[mt][2 * kmt + 1][0]);
      pdt[kmt][mt][2] = movmatrix_trans(pd[mt][2 * kmt][1]); pdt[kmt][mt][3] = movmatrix_trans(pd[mt][2 * kmt + 1][1]);
    }
#pragma unroll 1
  for (int nb = 0; nb < DH / 8; ++nb) {              // 8 features per step
    float dq[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, dk[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}},
          dv[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
    for (int kt = 0; kt < 2; ++kt) {                 // contraction index 16 kt .. 16 kt + 15 (keys for dQ, queries for dK / dV)
      const int64_t r = 16 * kt + 2 * t;
      const int col = 8 * nb + g;
      const uint32_t bk0 = am_pack_col(K + r * a.ldk + col, a.ldk), bk1 = am_pack_col(K + (r + 8) * a.ldk + col, a.ldk);
      const uint32_t bq0 = am_pack_col(Q + r * a.ldq + col, a.ldq), bq1 = am_pack_col(Q + (r + 8) * a.ldq + col, a.ldq);
      const uint32_t bo0 = am_pack_col(dO + r * a.ld_do + col, a.ld_do), bo1 = am_pack_col(dO + (r + 8) * a.ld_do + col, a.ld_do);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        mma16816(dq[u], ds[u][2 * kt][0], ds[u][2 * kt][1], ds[u][2 * kt + 1][0], ds[u][2 * kt + 1][1], bk0, bk1);
        mma16816(dk[u], dst[u][kt][0], dst[u][kt][1], dst[u][kt][2], dst[u][kt][3], bq0, bq1);
        mma16816(dv[u], pdt[u][kt][0], pdt[u][kt][1], pdt[u][kt][2], pdt[u][kt][3], bo0, bo1);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t r0 = 16 * u + g;
      const int c0 = 8 * nb + 2 * t;
      *reinterpret_cast<float2 *>(dQ + r0 * a.ld_dq + c0) = make_float2(dq[u][0], dq[u][1]);
      *reinterpret_cast<float2 *>(dQ + (r0 + 8) * a.ld_dq + c0) = make_float2(dq[u][2], dq[u][3]);
      *reinterpret_cast<float2 *>(dK + r0 * a.ld_dk + c0) = make_float2(dk[u][0], dk[u][1]);
      *reinterpret_cast<float2 *>(dK + (r0 + 8) * a.ld_dk + c0) = make_float2(dk[u][2], dk[u][3]);
      *reinterpret_cast<float2 *>(dV + r0 * a.ld_dv + c0) = make_float2(dv[u][0], dv[u][1]);
      *reinterpret_cast<float2 *>(dV + (r0 + 8) * a.ld_dv + c0) = make_float2(dv[u][2], dv[u][3]);
    }
  }
}

bool attention_tc_supported(const AttnArgs &a) {
  const bool dh_ok = a.dh == 16 || a.dh == 32 || a.dh == 64 || a.dh == 128;
  // float2 fragment loads / stores: even leading dimensions and 8-byte aligned bases (true for every activation buffer of the plan)
  auto ok = [](const void *p, int64_t ld) { return p == nullptr || (((uintptr_t)p & 7) == 0 && (ld & 1) == 0); };
  return dh_ok && ok(a.q, a.ldq) && ok(a.k, a.ldk) && ok(a.v, a.ldv) && ok(a.o, a.ldo) && ok(a.d_o, a.ld_do) && ok(a.dq, a.ld_dq) &&
         ok(a.dk, a.ld_dk) && ok(a.dv, a.ld_dv);
}

template <int DH>
static int am_launch(const AttnArgs &a, bool bwd, cudaStream_t st) {
  const int64_t pairs = a.n_seq * a.H;
  const unsigned grid = (unsigned)((pairs + AM_WARPS - 1) / AM_WARPS);
  { LaunchScope _ls(bwd ? KC_ATTN_BWD : KC_ATTN_FWD, st);
    if (bwd) attn_mma_bwd_kernel<DH><<<grid, AM_WARPS * 32, 0, st>>>(a);
    else attn_mma_fwd_kernel<DH><<<grid, AM_WARPS * 32, 0, st>>>(a); }
  GT_CUDA(cudaGetLastError());
  return 0;
}
static int am_dispatch(const AttnArgs &a, bool bwd, cudaStream_t st) {
  if (a.n_seq == 0) return 0;
  GT_CHECK(attention_tc_supported(a), "attention_tc: head dim / alignment not supported");
  switch (a.dh) {
    case 16: return am_launch<16>(a, bwd, st);
    case 32: return am_launch<32>(a, bwd, st);
    case 64: return am_launch<64>(a, bwd, st);
    default: return am_launch<128>(a, bwd, st);
  }
}
int attention_fwd_tc(const AttnArgs &a, cudaStream_t st) { return am_dispatch(a, false, st); }
int attention_bwd_tc(const AttnArgs &a, cudaStream_t st) { return am_dispatch(a, true, st); }

}  // namespace gt
