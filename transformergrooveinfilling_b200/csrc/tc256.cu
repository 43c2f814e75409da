// tc256.cu — fused encoder-layer kernels for d_model = 256.  See tc256.cuh for the design.
#include <stdlib.h>

#include "tc256_dev.cuh"

namespace gt {

bool t256_shape_supported(const gt_config &c, std::string *why) {
  auto no = [&](const char *m) { if (why) *why = m; return false; };
  if (c.n_dec != 0) return no("encoder-decoder models run in precision=fp32 (the fused tcgen05 layer kernels cover the encoder stack)");
  if (c.d_model != 256) return no("t256 kernels are instantiated for d_model=256");
  if (c.n_enc > TC_MAX_LAYERS) return no("more than 16 layers");
  const int dh = c.d_model / c.nhead;
  if (dh != 16 && dh != 32 && dh != 128) return no("d_model=256 tensor-core kernels need head_dim 16, 32 or 128 (nhead 16, 8 or 2)");
  if (c.dim_ff % 64 != 0 || c.dim_ff < 64 || c.dim_ff > 512) return no("d_model=256 tensor-core kernels need dim_feedforward in {64,128,...,512}");
  return true;
}

__device__ __forceinline__ uint4 t256_pack8(const float *s) {
  float4 a = *reinterpret_cast<const float4 *>(s), b = *reinterpret_cast<const float4 *>(s + 4);
  return make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
}

// =============================================================================================
// weight prep: fp32 master parameters -> forward + backward stage streams (bf16 canonical K-major
// B images).  grid = (stage blocks, layers)
// =============================================================================================
__global__ void t256_prep_kernel(TcPrepArgs a) {
  const int l = blockIdx.y;
  const int F = a.F;
  uint8_t *img = a.img + (size_t)l * a.img_stride + (size_t)blockIdx.z * (a.img_stride / T256_REP);
  const int nf = t256_fwd_stages(F), nst = nf + t256_bwd_stages(F, a.dh);
  const float *Wqkv = a.params + a.w_in[l], *Wo = a.params + a.w_out[l], *W1 = a.params + a.w1[l], *W2 = a.params + a.w2[l];
  for (int st = blockIdx.x; st < nst; st += gridDim.x) {
    const T256Stage s = st < nf ? t256_fwd_stage(st) : t256_bwd_stage(st - nf, F, a.dh);
    uint8_t *dst = img + (size_t)st * T256_STAGE;
    const int kbn = s.K / 8, total = s.N * kbn;
    for (int id = threadIdx.x; id < total; id += blockDim.x) {
      const int n = id / kbn, kb = id % kbn, k0 = kb * 8;
      uint4 out;
      if (s.type <= 3) {               // K runs along the contiguous dimension of the source
        const float *src;
        if (s.type == 0) {             // Wqkv [768, 256]: chunk row n -> (q|k|v part, 64 feature columns of group a)
          src = Wqkv + (int64_t)((n >> 6) * 256 + s.a * 64 + (n & 63)) * 256 + s.b * 32 + k0;
        } else if (s.type == 1) {      // Wo [256, 256]: K = ctx features of group a
          src = Wo + (int64_t)n * 256 + s.a * 64 + s.b * 32 + k0;
        } else if (s.type == 2) {      // W1 [F, 256]
          src = W1 + (int64_t)(s.a * 64 + n) * 256 + s.b * 128 + k0;
        } else {                       // W2 [256, F]
          src = W2 + (int64_t)n * F + s.a * 64 + s.b * 32 + k0;
        }
        out = t256_pack8(src);
      } else {                         // transposed operands: element (n, k) = W[krow(k)][n-dependent column]
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int k = k0 + j;
          if (s.type == 4) v[j] = W2[(int64_t)(s.b * 128 + k) * F + s.a * 64 + n];                     // (f, j) = W2[j][chunk f]
          else if (s.type == 5) v[j] = W1[(int64_t)(s.a * 64 + s.b * 32 + k) * 256 + n];               // (j, f) = W1[chunk f][j]
          else if (s.type == 6) v[j] = Wo[(int64_t)(s.b * 32 + k) * 256 + n];                          // (c, j) = Wo[j][c]
          else {                                                                                        // (j, m) = Wqkv[row(m)][j]
            const int m = s.b * 32 + k;
            v[j] = Wqkv[(int64_t)((m >> 6) * 256 + s.a * 64 + (m & 63)) * 256 + n];
          }
        }
        out = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
      }
      *reinterpret_cast<uint4 *>(dst + kmajor_off(n, k0, s.N)) = out;
    }
  }
}

int t256_prep_weights(const TcPrepArgs &a, cudaStream_t st) {
  dim3 grid(48, a.n_layers, T256_REP);
  { LaunchScope _ls(KC_TC_PREP, st);
    t256_prep_kernel<<<grid, 256, 0, st>>>(a); }
  GT_CUDA(cudaGetLastError());
  return 0;
}

// =============================================================================================
// layout conversion at the stack boundaries (row-major fp32 <-> tiled fp32 + bf16 image)
// =============================================================================================
__global__ void t256_to_image_kernel(const float *__restrict__ rm, uint8_t *__restrict__ img, int64_t M, int64_t n_chunks) {
  for (int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; id < n_chunks; id += (int64_t)gridDim.x * blockDim.x) {
    const int kb = (int)(id & 31);                 // 8-column chunk
    const int64_t grow = id >> 5;
    const int64_t tile = grow >> 7;
    const int r = (int)(grow & 127);
    float4 v0 = make_float4(0, 0, 0, 0), v1 = v0;
    if (grow < M) {
      v0 = *reinterpret_cast<const float4 *>(rm + grow * 256 + kb * 8);
      v1 = *reinterpret_cast<const float4 *>(rm + grow * 256 + kb * 8 + 4);
    }
    *reinterpret_cast<uint4 *>(img + tile * T256_TILE_IMG + kmajor_off(r, kb * 8, 128)) =
        make_uint4(pack_bf16(v0.x, v0.y), pack_bf16(v0.z, v0.w), pack_bf16(v1.x, v1.y), pack_bf16(v1.z, v1.w));
  }
}
__global__ void t256_from_image_kernel(const uint8_t *__restrict__ img, float *__restrict__ rm, int64_t M) {
  const int64_t n = M * 32;                          // 8-column chunks
  for (int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; id < n; id += (int64_t)gridDim.x * blockDim.x) {
    const int kb = (int)(id & 31);
    const int64_t grow = id >> 5;
    const int64_t tile = grow >> 7;
    const int r = (int)(grow & 127);
    const uint4 v = *reinterpret_cast<const uint4 *>(img + tile * T256_TILE_IMG + kmajor_off(r, kb * 8, 128));
    *reinterpret_cast<float4 *>(rm + grow * 256 + kb * 8) = make_float4(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xFFFF0000u), __uint_as_float(v.y << 16), __uint_as_float(v.y & 0xFFFF0000u));
    *reinterpret_cast<float4 *>(rm + grow * 256 + kb * 8 + 4) = make_float4(__uint_as_float(v.z << 16), __uint_as_float(v.z & 0xFFFF0000u), __uint_as_float(v.w << 16), __uint_as_float(v.w & 0xFFFF0000u));
  }
}
__global__ void t256_to_tiled_kernel(const float *__restrict__ rm, float *__restrict__ tiled, int64_t M, int64_t n_chunks) {
  for (int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; id < n_chunks; id += (int64_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(id & 63);
    const int64_t grow = id >> 6;
    const int64_t tile = grow >> 7;
    const int r = (int)(grow & 127);
    float4 v = make_float4(0, 0, 0, 0);
    if (grow < M) v = *reinterpret_cast<const float4 *>(rm + grow * 256 + c4 * 4);
    *reinterpret_cast<float4 *>(tiled + tile * T256_TILE_F32 + t256_tiled_off(r, c4 * 4)) = v;
  }
}
__global__ void t256_from_tiled_kernel(const float *__restrict__ tiled, float *__restrict__ rm, int64_t M) {
  const int64_t n = M * 64;                          // float4 chunks
  for (int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; id < n; id += (int64_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(id & 63);
    const int64_t grow = id >> 6;
    const int64_t tile = grow >> 7;
    const int r = (int)(grow & 127);
    *reinterpret_cast<float4 *>(rm + grow * 256 + c4 * 4) =
        *reinterpret_cast<const float4 *>(tiled + tile * T256_TILE_F32 + t256_tiled_off(r, c4 * 4));
  }
}
static int t256_ew_grid(int64_t n) { return (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16); }
int t256_to_image(const float *rowmajor, uint8_t *img, int64_t M, int n_tiles, cudaStream_t st) {
  const int64_t n = (int64_t)n_tiles * 128 * 32;
  { LaunchScope _ls(KC_ELEMWISE, st);
    t256_to_image_kernel<<<t256_ew_grid(n), 256, 0, st>>>(rowmajor, img, M, n); }
  GT_CUDA(cudaGetLastError());
  return 0;
}
int t256_from_image(const uint8_t *img, float *rowmajor, int64_t M, cudaStream_t st) {
  { LaunchScope _ls(KC_ELEMWISE, st);
    t256_from_image_kernel<<<t256_ew_grid(M * 32), 256, 0, st>>>(img, rowmajor, M); }
  GT_CUDA(cudaGetLastError());
  return 0;
}
int t256_to_tiled(const float *rowmajor, float *tiled, int64_t M, int n_tiles, cudaStream_t st) {
  const int64_t n = (int64_t)n_tiles * 128 * 64;
  { LaunchScope _ls(KC_ELEMWISE, st);
    t256_to_tiled_kernel<<<t256_ew_grid(n), 256, 0, st>>>(rowmajor, tiled, M, n); }
  GT_CUDA(cudaGetLastError());
  return 0;
}
int t256_from_tiled(const float *tiled, float *rowmajor, int64_t M, cudaStream_t st) {
  { LaunchScope _ls(KC_ELEMWISE, st);
    t256_from_tiled_kernel<<<t256_ew_grid(M * 64), 256, 0, st>>>(tiled, rowmajor, M); }
  GT_CUDA(cudaGetLastError());
  return 0;
}


struct T256FwdSmem {
  static constexpr uint32_t x = 0, ring = 65536, qkv = 131072, ctx = 180224, par = 212992, stat = 224256, total = 228352;
};

// ---- attention for NHP half pairs per warp, interleaved for ILP: 16 query rows of (sequence s, head hl of the
// group) against the 32 keys.  sQKV: canonical K-major image [128 rows x 192 cols] = [q (64) | k (64) | v (64)],
// q already scaled by log2(e)/sqrt(dh).  The context rows go to the [128 x 64] A image sCtxBuf.
template <int DH, int NHP>
__device__ __forceinline__ void t256_attn_fwd(const uint8_t *sQKV, uint8_t *sCtxBuf, const int (&s)[NHP], const int (&hl)[NHP],
                                              const int (&half)[NHP], int lane, const Drop &dr, const uint64_t (&w_pair)[NHP]) {
  const int g = lane >> 2, t = lane & 3;
  int r0[NHP];
  float sacc[NHP][4][4];
#pragma unroll
  for (int u = 0; u < NHP; ++u) {
    r0[u] = s[u] * 32 + half[u] * 16 + g;          // query rows r0 and r0 + 8
#pragma unroll
    for (int i = 0; i < 4; ++i) { sacc[u][i][0] = 0.f; sacc[u][i][1] = 0.f; sacc[u][i][2] = 0.f; sacc[u][i][3] = 0.f; }
  }
#pragma unroll
  for (int kt = 0; kt < DH / 16; ++kt) {
#pragma unroll
    for (int u = 0; u < NHP; ++u) {
      const int qc = hl[u] * DH + 16 * kt + 2 * t, kc = 64 + qc;
      const uint32_t a0 = lds32(sQKV + kmajor_off(r0[u], qc, 128));
      const uint32_t a1 = lds32(sQKV + kmajor_off(r0[u] + 8, qc, 128));
      const uint32_t a2 = lds32(sQKV + kmajor_off(r0[u], qc + 8, 128));
      const uint32_t a3 = lds32(sQKV + kmajor_off(r0[u] + 8, qc + 8, 128));
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int key = s[u] * 32 + 8 * nt + g;
        const uint32_t b0 = lds32(sQKV + kmajor_off(key, kc, 128));
        const uint32_t b1 = lds32(sQKV + kmajor_off(key, kc + 8, 128));
        mma16816(sacc[u][nt], a0, a1, a2, a3, b0, b1);
      }
    }
  }
  // softmax over the 32 keys of rows r0 (c0, c1) and r0 + 8 (c2, c3)
  float m0[NHP], m1[NHP], s0[NHP], s1[NHP];
#pragma unroll
  for (int u = 0; u < NHP; ++u) {
    m0[u] = sacc[u][0][0]; m1[u] = sacc[u][0][2];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      m0[u] = fmaxf(m0[u], fmaxf(sacc[u][nt][0], sacc[u][nt][1]));
      m1[u] = fmaxf(m1[u], fmaxf(sacc[u][nt][2], sacc[u][nt][3]));
    }
  }
#pragma unroll
  for (int o = 1; o <= 2; o <<= 1) {
#pragma unroll
    for (int u = 0; u < NHP; ++u) {
      m0[u] = fmaxf(m0[u], __shfl_xor_sync(0xffffffffu, m0[u], o));
      m1[u] = fmaxf(m1[u], __shfl_xor_sync(0xffffffffu, m1[u], o));
    }
  }
#pragma unroll
  for (int u = 0; u < NHP; ++u) {
    s0[u] = 0.f; s1[u] = 0.f;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      sacc[u][nt][0] = ex2_ftz(sacc[u][nt][0] - m0[u]); sacc[u][nt][1] = ex2_ftz(sacc[u][nt][1] - m0[u]);
      sacc[u][nt][2] = ex2_ftz(sacc[u][nt][2] - m1[u]); sacc[u][nt][3] = ex2_ftz(sacc[u][nt][3] - m1[u]);
      s0[u] += sacc[u][nt][0] + sacc[u][nt][1]; s1[u] += sacc[u][nt][2] + sacc[u][nt][3];
    }
  }
#pragma unroll
  for (int o = 1; o <= 2; o <<= 1) {
#pragma unroll
    for (int u = 0; u < NHP; ++u) {
      s0[u] += __shfl_xor_sync(0xffffffffu, s0[u], o);
      s1[u] += __shfl_xor_sync(0xffffffffu, s1[u], o);
    }
  }
#pragma unroll
  for (int u = 0; u < NHP; ++u) {
    const float i0 = dr.scale / s0[u], i1 = dr.scale / s1[u];
    if (dr.thr) {
      // quads of one query row (common.cuh: key_perm): quad t = this lane's keys of nt 0,1 ; quad 4 + t = those of nt 2,3
      const int q0 = half[u] * 16 + g;
      const uint64_t wa = w_pair[u] + (uint64_t)q0 * 8u, wb = wa + 64u;            // rows q0 and q0 + 8
      const uint32_t alo = (uint32_t)wa, ahi = (uint32_t)(wa >> 32) * 0x85EBCA6Bu;
      const uint32_t blo = (uint32_t)wb, bhi = (uint32_t)(wb >> 32) * 0x85EBCA6Bu;
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t la, ha, lb, hb;
        hash_quad((alo + (uint32_t)(4 * np + t)) ^ ahi, dr.key, la, ha);
        hash_quad((blo + (uint32_t)(4 * np + t)) ^ bhi, dr.key, lb, hb);
        sacc[u][2 * np][0] = ((la & 0xFFFFu) >= dr.thr) ? sacc[u][2 * np][0] * i0 : 0.f;
        sacc[u][2 * np][1] = ((la >> 16) >= dr.thr) ? sacc[u][2 * np][1] * i0 : 0.f;
        sacc[u][2 * np + 1][0] = ((ha & 0xFFFFu) >= dr.thr) ? sacc[u][2 * np + 1][0] * i0 : 0.f;
        sacc[u][2 * np + 1][1] = ((ha >> 16) >= dr.thr) ? sacc[u][2 * np + 1][1] * i0 : 0.f;
        sacc[u][2 * np][2] = ((lb & 0xFFFFu) >= dr.thr) ? sacc[u][2 * np][2] * i1 : 0.f;
        sacc[u][2 * np][3] = ((lb >> 16) >= dr.thr) ? sacc[u][2 * np][3] * i1 : 0.f;
        sacc[u][2 * np + 1][2] = ((hb & 0xFFFFu) >= dr.thr) ? sacc[u][2 * np + 1][2] * i1 : 0.f;
        sacc[u][2 * np + 1][3] = ((hb >> 16) >= dr.thr) ? sacc[u][2 * np + 1][3] * i1 : 0.f;
      }
    } else {
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) { sacc[u][nt][0] *= i0; sacc[u][nt][1] *= i0; sacc[u][nt][2] *= i1; sacc[u][nt][3] *= i1; }
    }
  }
  // O = P V
  float oacc[NHP][DH / 8][4];
#pragma unroll
  for (int u = 0; u < NHP; ++u)
#pragma unroll
    for (int i = 0; i < DH / 8; ++i) { oacc[u][i][0] = 0.f; oacc[u][i][1] = 0.f; oacc[u][i][2] = 0.f; oacc[u][i][3] = 0.f; }
#pragma unroll
  for (int kt = 0; kt < 2; ++kt) {                  // keys 16 kt .. 16 kt + 15
#pragma unroll
    for (int u = 0; u < NHP; ++u) {
      const uint32_t p0 = pack_bf16(sacc[u][2 * kt][0], sacc[u][2 * kt][1]), p1 = pack_bf16(sacc[u][2 * kt][2], sacc[u][2 * kt][3]);
      const uint32_t p2 = pack_bf16(sacc[u][2 * kt + 1][0], sacc[u][2 * kt + 1][1]), p3 = pack_bf16(sacc[u][2 * kt + 1][2], sacc[u][2 * kt + 1][3]);
#pragma unroll
      for (int np = 0; np < DH / 16; ++np) {        // feature columns 16 np .. 16 np + 15
        const int mi = lane >> 3, rr = lane & 7;
        const int key = s[u] * 32 + 16 * kt + (mi & 1) * 8 + rr;
        uint32_t b[4];
        ldmatrix_x4_trans(b, sQKV + kmajor_off(key, 128 + hl[u] * DH + 16 * np + (mi >> 1) * 8, 128));
        mma16816(oacc[u][2 * np], p0, p1, p2, p3, b[0], b[1]);
        mma16816(oacc[u][2 * np + 1], p0, p1, p2, p3, b[2], b[3]);
      }
    }
  }
#pragma unroll
  for (int u = 0; u < NHP; ++u)
#pragma unroll
    for (int nt = 0; nt < DH / 8; ++nt) {
      *reinterpret_cast<uint32_t *>(sCtxBuf + kmajor_off(r0[u], hl[u] * DH + 8 * nt + 2 * t, 128)) = pack_bf16(oacc[u][nt][0], oacc[u][nt][1]);
      *reinterpret_cast<uint32_t *>(sCtxBuf + kmajor_off(r0[u] + 8, hl[u] * DH + 8 * nt + 2 * t, 128)) = pack_bf16(oacc[u][nt][2], oacc[u][nt][3]);
    }
}

// ---- head_dim 128 (InfillingKicksAndSnares: 2 heads of 128): a head spans TWO 64-column groups ---------------------------
// One warp owns one unit = 16 query rows (m-tile `half`) of (sequence s, the head the group pair belongs to).  The scores
// contract over all 128 features, i.e. over both groups: the EVEN group adds its 64-feature partial into the warp's
// accumulators (kept in registers until the odd group arrives), the ODD group completes them.
__device__ __forceinline__ void t256_attn128_scores(const uint8_t *sQKV, int s, int half, int lane, float (&sacc)[4][4]) {
  const int g = lane >> 2, t = lane & 3;
  const int r0 = s * 32 + half * 16 + g;
#pragma unroll
  for (int kt = 0; kt < 4; ++kt) {
    const int qc = 16 * kt + 2 * t, kc = 64 + qc;
    const uint32_t a0 = lds32(sQKV + kmajor_off(r0, qc, 128));
    const uint32_t a1 = lds32(sQKV + kmajor_off(r0 + 8, qc, 128));
    const uint32_t a2 = lds32(sQKV + kmajor_off(r0, qc + 8, 128));
    const uint32_t a3 = lds32(sQKV + kmajor_off(r0 + 8, qc + 8, 128));
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int key = s * 32 + 8 * nt + g;
      mma16816(sacc[nt], a0, a1, a2, a3, lds32(sQKV + kmajor_off(key, kc, 128)), lds32(sQKV + kmajor_off(key, kc + 8, 128)));
    }
  }
}
// softmax + dropout of the completed scores -> packed bf16 probability fragments pa[key n-tile][row g / g + 8]
__device__ __forceinline__ void t256_attn128_probs(float (&sacc)[4][4], uint32_t (&pa)[4][2], int half, int lane, const Drop &dr, uint64_t w_pair) {
  const int g = lane >> 2, t = lane & 3;
  float m0 = sacc[0][0], m1 = sacc[0][2];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) { m0 = fmaxf(m0, fmaxf(sacc[nt][0], sacc[nt][1])); m1 = fmaxf(m1, fmaxf(sacc[nt][2], sacc[nt][3])); }
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    sacc[nt][0] = ex2_ftz(sacc[nt][0] - m0); sacc[nt][1] = ex2_ftz(sacc[nt][1] - m0);
    sacc[nt][2] = ex2_ftz(sacc[nt][2] - m1); sacc[nt][3] = ex2_ftz(sacc[nt][3] - m1);
    s0 += sacc[nt][0] + sacc[nt][1]; s1 += sacc[nt][2] + sacc[nt][3];
  }
  s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
  s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
  const float i0 = dr.scale / s0, i1 = dr.scale / s1;
  if (dr.thr) {
    const int q0 = half * 16 + g;
    const uint64_t wa = w_pair + (uint64_t)q0 * 8u, wb = wa + 64u;            // rows q0 and q0 + 8
    const uint32_t alo = (uint32_t)wa, ahi = (uint32_t)(wa >> 32) * 0x85EBCA6Bu;
    const uint32_t blo = (uint32_t)wb, bhi = (uint32_t)(wb >> 32) * 0x85EBCA6Bu;
#pragma unroll
    for (int np = 0; np < 2; ++np) {
      uint32_t la, ha, lb, hb;
      hash_quad((alo + (uint32_t)(4 * np + t)) ^ ahi, dr.key, la, ha);
      hash_quad((blo + (uint32_t)(4 * np + t)) ^ bhi, dr.key, lb, hb);
      sacc[2 * np][0] = ((la & 0xFFFFu) >= dr.thr) ? sacc[2 * np][0] * i0 : 0.f;
      sacc[2 * np][1] = ((la >> 16) >= dr.thr) ? sacc[2 * np][1] * i0 : 0.f;
      sacc[2 * np + 1][0] = ((ha & 0xFFFFu) >= dr.thr) ? sacc[2 * np + 1][0] * i0 : 0.f;
      sacc[2 * np + 1][1] = ((ha >> 16) >= dr.thr) ? sacc[2 * np + 1][1] * i0 : 0.f;
      sacc[2 * np][2] = ((lb & 0xFFFFu) >= dr.thr) ? sacc[2 * np][2] * i1 : 0.f;
      sacc[2 * np][3] = ((lb >> 16) >= dr.thr) ? sacc[2 * np][3] * i1 : 0.f;
      sacc[2 * np + 1][2] = ((hb & 0xFFFFu) >= dr.thr) ? sacc[2 * np + 1][2] * i1 : 0.f;
      sacc[2 * np + 1][3] = ((hb >> 16) >= dr.thr) ? sacc[2 * np + 1][3] * i1 : 0.f;
    }
  } else {
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) { sacc[nt][0] *= i0; sacc[nt][1] *= i0; sacc[nt][2] *= i1; sacc[nt][3] *= i1; }
  }
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) { pa[nt][0] = pack_bf16(sacc[nt][0], sacc[nt][1]); pa[nt][1] = pack_bf16(sacc[nt][2], sacc[nt][3]); }
}
// O[16 x 64] = P V for 64 value columns: V rows (keys) of sequence s start at column vcol0 of the K-major image `vimg` (R = 128)
__device__ __forceinline__ void t256_attn128_pv(const uint32_t (&pa)[4][2], const uint8_t *vimg, int vcol0, int s, int lane, float (&oacc)[8][4]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) { oacc[i][0] = 0.f; oacc[i][1] = 0.f; oacc[i][2] = 0.f; oacc[i][3] = 0.f; }
  const int mi = lane >> 3, rr = lane & 7;
#pragma unroll
  for (int kt = 0; kt < 2; ++kt) {                  // keys 16 kt .. 16 kt + 15
#pragma unroll
    for (int np = 0; np < 4; ++np) {                // value columns 16 np .. 16 np + 15
      const int key = s * 32 + 16 * kt + (mi & 1) * 8 + rr;
      uint32_t b[4];
      ldmatrix_x4_trans(b, vimg + kmajor_off(key, vcol0 + 16 * np + (mi >> 1) * 8, 128));
      mma16816(oacc[2 * np], pa[2 * kt][0], pa[2 * kt][1], pa[2 * kt + 1][0], pa[2 * kt + 1][1], b[0], b[1]);
      mma16816(oacc[2 * np + 1], pa[2 * kt][0], pa[2 * kt][1], pa[2 * kt + 1][0], pa[2 * kt + 1][1], b[2], b[3]);
    }
  }
}
__device__ __forceinline__ void t256_attn128_store_ctx(uint8_t *ctx_buf, const float (&oacc)[8][4], int s, int half, int lane) {
  const int g = lane >> 2, t = lane & 3, r0 = s * 32 + half * 16 + g;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    *reinterpret_cast<uint32_t *>(ctx_buf + kmajor_off(r0, 8 * nt + 2 * t, 128)) = pack_bf16(oacc[nt][0], oacc[nt][1]);
    *reinterpret_cast<uint32_t *>(ctx_buf + kmajor_off(r0 + 8, 8 * nt + 2 * t, 128)) = pack_bf16(oacc[nt][2], oacc[nt][3]);
  }
}

// =============================================================================================
// forward
// =============================================================================================
template <int DH, bool DEVSTEP = false>
__global__ void __launch_bounds__(T256_THREADS, 1) t256_layer_fwd_kernel(const T256Args a_in) {
  const DropArgsView<T256Args, DEVSTEP> view(a_in);
  const T256Args &a = view.a;
  constexpr int G = T256_G, GH = DH >= 64 ? 1 : 64 / DH, NS = T256_NS, NHP = DH >= 64 ? 1 : (8 * GH) / 16;
  using S = T256FwdSmem;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_full[NS], bar_empty[NS], bar_xready, bar_xfree, bar_qkvfull, bar_qkvfree, bar_ctxready[2],
      bar_ctxfree[2], bar_outfull, bar_x1ready, bar_hfull[2], bar_hready[2], bar_out2full;
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
    mbar_init(&bar_xready, 1); mbar_init(&bar_xfree, 1); mbar_init(&bar_qkvfull, 1); mbar_init(&bar_qkvfree, 1);
    mbar_init(&bar_ctxready[0], 1); mbar_init(&bar_ctxready[1], 1); mbar_init(&bar_ctxfree[0], 1); mbar_init(&bar_ctxfree[1], 1);
    mbar_init(&bar_outfull, 1); mbar_init(&bar_x1ready, 1); mbar_init(&bar_hfull[0], 1); mbar_init(&bar_hfull[1], 1);
    mbar_init(&bar_hready[0], 1); mbar_init(&bar_hready[1], 1); mbar_init(&bar_out2full, 1);
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
  // FFN chunks alternate between two (hidden accumulator, H image) slots: chunk n = it * NCH + c uses slot n & 1 — accumulator columns
  // 448.. / 256.. (the q | k | v accumulator is idle during the FFN), image at sH + slot * 16384 (sH aliases the 48 KB q | k | v image)
  const uint32_t t_out = tmem, t_qkv = tmem + 256;
  auto t_h_slot = [&](uint32_t fs) { return fs ? tmem + 256u : tmem + 448u; };
  const uint32_t aX = smem_u32(sX), aRing = smem_u32(sRing), aH = smem_u32(sH), aCtx = smem_u32(sCtx);
  if (a.stagger) {                                    // see t256_launch_bwd
    const long long until = clock64() + (long long)(blockIdx.x & 3) * (long long)a.stagger;
    while (clock64() < until) __nanosleep(200);
  }

  if (warp == 16) {
    // ======================= TMA producer: x image of the tile, then its weight stage stream =======================
    // One elected lane runs the whole program (a single-lane region: no re-convergence inside the loop).  The
    // stream has a multiple of NS stages per tile, so every stage's ring slot is a compile-time constant.
    if (elect_one()) {
      const uint8_t *wimg = a.img + (size_t)(blockIdx.x % T256_REP) * a.img_rep_stride;
      const uint64_t pol_w = l2_policy_evict_last();
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
        mbar_wait(&bar_xfree, (it & 1u) ^ 1u);                 // FFN1 of the previous tile no longer reads sX
        mbar_expect_tx(&bar_xready, 65536u);
        tma_load_1d(sX, a.x_img_in + (size_t)tile * T256_TILE_IMG, 32768u, &bar_xready);
        tma_load_1d(sX + 32768, a.x_img_in + (size_t)tile * T256_TILE_IMG + 32768, 32768u, &bar_xready);
        const uint32_t use0 = it * (uint32_t)(10 + NCH);         // ring wraps per tile: (40 + 4 NCH) / NS
#pragma unroll 1
        for (int st = 0; st < 40; ++st) {
          constexpr uint32_t kQ = 192 * 32 * 2;
          const int slot = st % NS;
          const uint32_t use = use0 + (uint32_t)(st / NS);
          // stages 0..7 QKV(0); then per g = 1..3: 8 x QKV(g), 2 x Wo(g-1); last two: Wo(3)
          const bool is_wo = st >= 38 || (st >= 8 && ((st - 8) % 10) >= 8);
          const uint32_t bytes = is_wo ? 16384u : kQ;
          mbar_wait(&bar_empty[slot], (use & 1u) ^ 1u);
          if (a.dbg && blockIdx.x == 0 && it == 2) a.dbg[192 + st] = clock64();
          mbar_expect_tx(&bar_full[slot], bytes);
          tma_load_1d_hint(sRing + slot * T256_STAGE, wimg + (size_t)st * T256_STAGE, bytes, &bar_full[slot], pol_w);
        }
        for (int c = 0; c < NCH; ++c) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t use = use0 + 10u + (uint32_t)c;
            mbar_wait(&bar_empty[j], (use & 1u) ^ 1u);
            mbar_expect_tx(&bar_full[j], 16384u);
            tma_load_1d_hint(sRing + j * T256_STAGE, wimg + (size_t)(40 + c * 4 + j) * T256_STAGE, 16384u, &bar_full[j], pol_w);
          }
        }
      }
    }
  } else if (warp == 17) {
    // ======================= MMA issuer: one elected lane runs the whole program =======================
    if (elect_one()) {
      const uint32_t id_qkv = make_idesc_bf16(128, 192), id_256 = make_idesc_bf16(128, 256), id_64 = make_idesc_bf16(128, 64);
      const uint64_t dX = descA128(aX);
      const uint64_t dH = descA128(aH);
      const uint64_t dC[2] = {descA128(aCtx), descA128(aCtx + 16384u)};
      const uint64_t dR192 = descB(aRing, 192), dR256 = descB(aRing, 256), dR64 = descB(aRing, 64);   // slot 0; slot j adds j * STAGE
      uint32_t it = 0, nstamp = 0;
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
        const uint32_t use0 = it * (uint32_t)(10 + NCH);
        // st: static index of the stage inside the tile's stream
        auto full_wait = [&](int st, uint32_t use) {
          mbar_wait(&bar_full[st % NS], use & 1u);
          fence_after_sync();
          if (a.dbg && blockIdx.x == 0 && it == 2 && nstamp < 120) a.dbg[64 + nstamp++] = clock64();
        };
        auto wo = [&](int gg, int st0) {               // out += ctx[:, group gg] Wo[:, group gg]^T   (K = 64: two stages)
          const int b = gg & 1;
          const uint32_t nb = it * (G / 2) + (uint32_t)(gg >> 1);
          mbar_wait(&bar_ctxready[b], nb & 1u);
          fence_after_sync();
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const int st = st0 + hf;
            full_wait(st, use0 + (uint32_t)(st / NS));
            const uint64_t db = desc_adv(dR256, (uint32_t)(st % NS) * T256_STAGE);
#pragma unroll
            for (int k = 0; k < 2; ++k)
              mma_bf16_ss(t_out, desc_adv(dC[b], (uint32_t)(hf * 2 + k) * 4096u), desc_adv(db, (uint32_t)k * 8192u), id_256, (gg | hf | k) > 0);
            mma_commit(&bar_empty[st % NS]);
          }
          mma_commit(&bar_ctxfree[b]);
        };
        mbar_wait(&bar_xready, it & 1u);
        fence_after_sync();
#pragma unroll 1
        for (int g = 0; g < G; ++g) {
          const int st0 = g == 0 ? 0 : 8 + (g - 1) * 10;
          const uint32_t nq = it * G + (uint32_t)g;
          mbar_wait(&bar_qkvfree, (nq & 1u) ^ 1u);
          fence_after_sync();
#pragma unroll 2
          for (int kc = 0; kc < 8; ++kc) {
            const int st = st0 + kc;
            full_wait(st, use0 + (uint32_t)(st / NS));
            const uint64_t db = desc_adv(dR192, (uint32_t)(st % NS) * T256_STAGE);
#pragma unroll
            for (int k = 0; k < 2; ++k)
              mma_bf16_ss(t_qkv, desc_adv(dX, (uint32_t)(kc * 2 + k) * 4096u), desc_adv(db, (uint32_t)k * 6144u), id_qkv, (kc | k) > 0);
            mma_commit(&bar_empty[st % NS]);
          }
          mma_commit(&bar_qkvfull);
          if (g >= 1) wo(g - 1, st0 + 8);
        }
        wo(G - 1, 38);
        mma_commit(&bar_outfull);
        // ---- FFN ----
        mbar_wait(&bar_x1ready, it & 1u);
        fence_after_sync();
        auto ffn1 = [&](int c) {
          const uint32_t fs = (it * (uint32_t)NCH + (uint32_t)c) & 1u;
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            full_wait(hf, use0 + 10u + (uint32_t)c);
            const uint64_t db = desc_adv(dR64, (uint32_t)hf * T256_STAGE);
#pragma unroll
            for (int k = 0; k < 8; ++k)
              mma_bf16_ss(t_h_slot(fs), desc_adv(dX, (uint32_t)(hf * 8 + k) * 4096u), desc_adv(db, (uint32_t)k * 2048u), id_64, (hf | k) > 0);
            mma_commit(&bar_empty[hf]);
          }
          mma_commit(&bar_hfull[fs]);
          if (c + 1 == NCH) mma_commit(&bar_xfree);        // sX may be refilled with the next tile's x image
        };
        // FFN1 of chunk c + 1 is issued BEFORE the issuer waits for the H image of chunk c (other accumulator / image slot), so the
        // hidden-activation epilogue overlaps the tensor work of its neighbours; ring slots 0, 1 (W1) run one use ahead of 2, 3 (W2)
        ffn1(0);
        for (int c = 0; c < NCH; ++c) {
          const uint32_t nh = it * (uint32_t)NCH + (uint32_t)c, fs = nh & 1u;
          if (c + 1 < NCH) ffn1(c + 1);
          mbar_wait(&bar_hready[fs], (nh >> 1) & 1u);
          fence_after_sync();
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            full_wait(2 + hf, use0 + 10u + (uint32_t)c);
            const uint64_t db = desc_adv(dR256, (uint32_t)(2 + hf) * T256_STAGE);
#pragma unroll
            for (int k = 0; k < 2; ++k)
              mma_bf16_ss(t_out, desc_adv(dH, fs * 16384u + (uint32_t)(hf * 2 + k) * 4096u), desc_adv(db, (uint32_t)k * 8192u), id_256, (c | hf | k) > 0);
            mma_commit(&bar_empty[2 + hf]);
          }
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
    uint32_t it = 0;
    int ndbg = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
      const bool dbg_on = a.dbg != nullptr && blockIdx.x == 0 && tid == 0 && it == 2;
      T256_STAMP();
      const int64_t grow = (int64_t)tile * TC_TILE + row;
      // ---- P1: head groups ----
      float s128[4][4];                                 // head_dim 128: partial scores of the even group (warps 0..7)
      for (int g = 0; g < G; ++g) {
        const uint32_t nq = it * G + (uint32_t)g;
        mbar_wait(&bar_qkvfull, nq & 1u);
        fence_after_sync();
        T256_STAMP();
        {
          float v[48];
#pragma unroll
          for (int i = 0; i < 3; ++i) tmem_ld16(t_qkv + lane_off + (uint32_t)(part * 48 + i * 16), v + i * 16);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            const int n0 = part * 48 + i * 8, pq = n0 >> 6;
            const float *bias = p_bqkv + pq * 256 + g * 64 + (n0 & 63);
            const float sc = pq == 0 ? attn_scale : 1.f;
            float w[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) w[j] = (v[i * 8 + j] + bias[j]) * sc;
            *reinterpret_cast<uint4 *>(sQKV + kmajor_off(row, n0, 128)) =
                make_uint4(pack_bf16(w[0], w[1]), pack_bf16(w[2], w[3]), pack_bf16(w[4], w[5]), pack_bf16(w[6], w[7]));
          }
        }
        fence_before_sync();
        if (tid == 0) tma_store_wait_read();                    // bulk stores that were still reading sCtx / sH / sQKV have drained
        if constexpr (DH == 128) fence_async_smem();            // the q | k | v image is bulk-stored below
        named_bar_sync(1, T256_CTHREADS);
        if (tid == 0) mbar_arrive(&bar_qkvfree);
        T256_STAMP();
        const int b = g & 1;
        {
          const uint32_t nb = it * (G / 2) + (uint32_t)(g >> 1);
          mbar_wait(&bar_ctxfree[b], (nb & 1u) ^ 1u);          // the out-projection that read this ctx buffer two groups ago retired
        }
        if constexpr (DH == 128) {
          // ---- a head of 128 = groups (2h, 2h + 1).  Train: the group's q | k | v image is saved for the backward (no recompute there)
          if (tid == 0 && a.qkv_img) {
            tma_store_1d(a.qkv_img + ((size_t)tile * G + g) * T256_QKV_GROUP_IMG, sQKV, (uint32_t)T256_QKV_GROUP_IMG);
            tma_store_commit();
          }
          const int us = warp >> 1, uhalf = warp & 1;            // unit of warps 0..7: 16 query rows of sequence us
          if ((g & 1) == 0) {
            if (warp < 8) {
#pragma unroll
              for (int i = 0; i < 4; ++i) { s128[i][0] = 0.f; s128[i][1] = 0.f; s128[i][2] = 0.f; s128[i][3] = 0.f; }
              t256_attn128_scores(sQKV, us, uhalf, lane, s128);
            }
            // park the even group's 64 value columns in ctx buffer 0 (free until this head's context is written): the q | k | v
            // image is overwritten by the odd group before the probabilities exist
            *reinterpret_cast<uint4 *>(sCtx + kmajor_off(row, part * 16, 128)) = *reinterpret_cast<const uint4 *>(sQKV + kmajor_off(row, 128 + part * 16, 128));
            *reinterpret_cast<uint4 *>(sCtx + kmajor_off(row, part * 16 + 8, 128)) = *reinterpret_cast<const uint4 *>(sQKV + kmajor_off(row, 128 + part * 16 + 8, 128));
            if (tid == 0) tma_store_wait_read();                 // the bulk store of this group's q | k | v image has read sQKV: the odd group may overwrite it
            named_bar_sync(1, T256_CTHREADS);
            T256_STAMP();
            continue;                                            // no context yet: both ctxready barriers are raised after the odd group
          }
          if (warp < 8) {
            t256_attn128_scores(sQKV, us, uhalf, lane, s128);
            const int64_t seq = a.seq0 + (int64_t)tile * 4 + us;
            const uint64_t w_pair = (uint64_t)((seq * H + (g >> 1)) * 32) * 8u;
            uint32_t pa[4][2];
            t256_attn128_probs(s128, pa, uhalf, lane, a.d_attn, w_pair);
            float oacc[8][4];
            t256_attn128_pv(pa, sQKV, 128, us, lane, oacc);                 // odd group's value columns: still in the q | k | v image
            t256_attn128_store_ctx(sCtx + 16384, oacc, us, uhalf, lane);
            t256_attn128_pv(pa, sCtx, 0, us, lane, oacc);                   // even group's value columns: parked in ctx buffer 0 ...
            named_bar_sync(2, 256);                                          // ... which every unit has now finished reading
            t256_attn128_store_ctx(sCtx, oacc, us, uhalf, lane);
          }
          if (tid == 0) tma_store_wait_read();                   // (as above: sQKV is rewritten by the next group / the FFN's hidden image)
          fence_async_smem();
          named_bar_sync(1, T256_CTHREADS);
          if (tid == 0) {
            mbar_arrive(&bar_ctxready[0]);
            mbar_arrive(&bar_ctxready[1]);
            if (a.ctx_img) {
              tma_store_1d(a.ctx_img + (size_t)tile * T256_TILE_IMG + (size_t)(g - 1) * 16384, sCtx, 16384u);
              tma_store_1d(a.ctx_img + (size_t)tile * T256_TILE_IMG + (size_t)g * 16384, sCtx + 16384, 16384u);
              tma_store_commit();
            }
          }
          T256_STAMP();
          continue;
        }
        if (warp < 8 * GH / NHP) {
          int s[NHP], hl[NHP], half[NHP];
          uint64_t w_pair[NHP];
#pragma unroll
          for (int u = 0; u < NHP; ++u) {
            const int hp = warp + 16 * u, pair = hp >> 1;
            half[u] = hp & 1; s[u] = pair / GH; hl[u] = pair % GH;
            const int64_t seq = a.seq0 + (int64_t)tile * 4 + s[u];
            w_pair[u] = (uint64_t)((seq * H + (g * GH + hl[u])) * 32) * 8u;      // quad index of (row 0, position 0) of this (sequence, head)
          }
          if constexpr (DH != 128) t256_attn_fwd<DH, NHP>(sQKV, sCtx + b * 16384, s, hl, half, lane, a.d_attn, w_pair);
        }
        fence_async_smem();
        named_bar_sync(1, T256_CTHREADS);
        if (tid == 0) {
          mbar_arrive(&bar_ctxready[b]);
          if (a.ctx_img) {
            tma_store_1d(a.ctx_img + (size_t)tile * T256_TILE_IMG + (size_t)g * 16384, sCtx + b * 16384, 16384u);
            tma_store_commit();
          }
        }
        T256_STAMP();
      }
      // ---- P2: + bias, dropout, + residual, LayerNorm1 -> x1 (registers, fp32) and its bf16 A image ----
      float u[64];
      mbar_wait(&bar_outfull, it & 1u);
      fence_after_sync();
      T256_STAMP();
#pragma unroll
      for (int c = 0; c < 64; c += 8) {               // residual = the bf16 x image still resident in sX (every q|k|v UMMA has retired)
        const uint4 xv = *reinterpret_cast<const uint4 *>(sX + kmajor_off(row, part * 64 + c, 128));
        u[c] = bf16lo(xv.x); u[c + 1] = bf16hi(xv.x); u[c + 2] = bf16lo(xv.y); u[c + 3] = bf16hi(xv.y);
        u[c + 4] = bf16lo(xv.z); u[c + 5] = bf16hi(xv.z); u[c + 6] = bf16lo(xv.w); u[c + 7] = bf16hi(xv.w);
      }
      {
        const uint64_t w0 = (uint64_t)((a.seq0 * 32 + grow) * 256 + part * 64) >> 2;
        const uint32_t wlo = (uint32_t)w0, xhi = (uint32_t)(w0 >> 32) * 0x85EBCA6Bu;
        const float *bo = p_bo + part * 64;
        uint8_t *u1g = a.u1_img ? a.u1_img + (size_t)tile * T256_TILE_IMG : nullptr;
        float s1 = 0.f;
#pragma unroll
        for (int cb = 0; cb < 64; cb += 16) {
          float f[16];
          tmem_ld16(t_out + lane_off + (uint32_t)(part * 64 + cb), f);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            float m0 = 1.f, m1 = 1.f, m2 = 1.f, m3 = 1.f;
            if (a.d1.thr) {
              uint32_t lo, hi;
              hash_quad((wlo + (uint32_t)((cb + j) >> 2)) ^ xhi, a.d1.key, lo, hi);
              m0 = ((lo & 0xFFFFu) >= a.d1.thr) ? a.d1.scale : 0.f; m1 = ((lo >> 16) >= a.d1.thr) ? a.d1.scale : 0.f;
              m2 = ((hi & 0xFFFFu) >= a.d1.thr) ? a.d1.scale : 0.f; m3 = ((hi >> 16) >= a.d1.thr) ? a.d1.scale : 0.f;
            }
            u[cb + j] += (f[j] + bo[cb + j]) * m0;
            u[cb + j + 1] += (f[j + 1] + bo[cb + j + 1]) * m1;
            u[cb + j + 2] += (f[j + 2] + bo[cb + j + 2]) * m2;
            u[cb + j + 3] += (f[j + 3] + bo[cb + j + 3]) * m3;
            s1 += (u[cb + j] + u[cb + j + 1]) + (u[cb + j + 2] + u[cb + j + 3]);
          }
          if (u1g) {
#pragma unroll
            for (int j = 0; j < 16; j += 8)
              *reinterpret_cast<uint4 *>(u1g + kmajor_off(row, part * 64 + cb + j, 128)) =
                  make_uint4(pack_bf16(u[cb + j], u[cb + j + 1]), pack_bf16(u[cb + j + 2], u[cb + j + 3]), pack_bf16(u[cb + j + 4], u[cb + j + 5]),
                             pack_bf16(u[cb + j + 6], u[cb + j + 7]));
          }
        }
        T256_STAMP();
        sStatA[row * 4 + part] = s1;
        named_bar_sync(1, T256_CTHREADS);
        const float4 sa = *reinterpret_cast<const float4 *>(sStatA + row * 4);
        const float mu = ((sa.x + sa.y) + (sa.z + sa.w)) * (1.f / 256);
        float qq = 0.f;
#pragma unroll
        for (int c = 0; c < 64; ++c) { u[c] -= mu; qq = fmaf(u[c], u[c], qq); }
        sStatB[row * 4 + part] = qq;
        named_bar_sync(1, T256_CTHREADS);
        const float4 sb = *reinterpret_cast<const float4 *>(sStatB + row * 4);
        const float rs = rsqrtf(((sb.x + sb.y) + (sb.z + sb.w)) * (1.f / 256) + LN_EPS);
        if (part == 0 && a.ln1_stat) a.ln1_stat[grow] = make_float2(mu, rs);      // the backward takes the row statistics from here
        const float *g1 = p_g1 + part * 64, *be1 = p_be1 + part * 64;
        uint8_t *x1g = a.x1_img ? a.x1_img + (size_t)tile * T256_TILE_IMG : nullptr;
#pragma unroll
        for (int c = 0; c < 64; c += 8) {
#pragma unroll
          for (int j = 0; j < 8; ++j) u[c + j] = u[c + j] * rs * g1[c + j] + be1[c + j];
          const uint4 pk = make_uint4(pack_bf16(u[c], u[c + 1]), pack_bf16(u[c + 2], u[c + 3]), pack_bf16(u[c + 4], u[c + 5]), pack_bf16(u[c + 6], u[c + 7]));
          const uint32_t off = kmajor_off(row, part * 64 + c, 128);
          *reinterpret_cast<uint4 *>(sX + off) = pk;
          if (x1g) *reinterpret_cast<uint4 *>(x1g + off) = pk;
        }
      }
      fence_async_smem();
      fence_before_sync();
      named_bar_sync(1, T256_CTHREADS);
      if (tid == 0) mbar_arrive(&bar_x1ready);
      T256_STAMP();
      // ---- P3: FFN hidden chunks: + bias, ReLU, dropout -> bf16 H image ----
      for (int c = 0; c < NCH; ++c) {
        const uint32_t nh = it * (uint32_t)NCH + (uint32_t)c, fs = nh & 1u;
        uint8_t *sHs = sH + fs * 16384u;
        mbar_wait(&bar_hfull[fs], (nh >> 1) & 1u);
        fence_after_sync();
        T256_STAMP();
        float v[16];
        tmem_ld16(t_h_slot(fs) + lane_off + (uint32_t)(part * 16), v);
        tmem_ld_wait();
        const uint64_t w0 = (uint64_t)((a.seq0 * 32 + grow) * F + c * 64 + part * 16) >> 2;
        const uint32_t wlo = (uint32_t)w0, xhi = (uint32_t)(w0 >> 32) * 0x85EBCA6Bu;
        const float *b1 = p_b1 + c * 64 + part * 16;
        // ReLU fused into the bf16 conversion, two dropout decisions per packed half-precision compare (umma.cuh: keep2);
        // relu((v + b) s) = relu(v + b) s for s > 0, one rounding either way
        const float hs = a.d_ffn.thr ? a.d_ffn.scale : 1.f;
        const uint32_t thr2 = a.d_ffn.thr | (a.d_ffn.thr << 16);
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 bb = *reinterpret_cast<const float4 *>(b1 + j);
          pk[j >> 1] = pack_bf16_relu((v[j] + bb.x) * hs, (v[j + 1] + bb.y) * hs);
          pk[(j >> 1) + 1] = pack_bf16_relu((v[j + 2] + bb.z) * hs, (v[j + 3] + bb.w) * hs);
          if (a.d_ffn.thr) {
            uint32_t lo, hi;
            hash_quad((wlo + (uint32_t)(j >> 2)) ^ xhi, a.d_ffn.key, lo, hi);
            pk[j >> 1] &= keep2(lo, thr2);
            pk[(j >> 1) + 1] &= keep2(hi, thr2);
          }
        }
        // this slot's previous H image (chunk n - 2) is no longer read: its out-projection UMMAs retired before bar_hfull of this
        // chunk was committed, and its bulk store finished reading before the closing barrier of chunk n - 1 (below)
        *reinterpret_cast<uint4 *>(sHs + kmajor_off(row, part * 16, 128)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4 *>(sHs + kmajor_off(row, part * 16 + 8, 128)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        fence_async_smem();
        fence_before_sync();
        if (tid == 0) tma_store_wait_read();           // the bulk store of the PREVIOUS chunk's image (other slot) has drained: chunk n + 1 may overwrite it
        named_bar_sync(1, T256_CTHREADS);
        if (tid == 0) {
          mbar_arrive(&bar_hready[fs]);
          if (a.h_img) {
            tma_store_1d(a.h_img + ((size_t)tile * NCH + c) * 16384, sHs, 16384u);
            tma_store_commit();
          }
        }
        T256_STAMP();
      }
      // ---- P4: + bias, dropout, + residual (x1), LayerNorm2 -> x_out ----
      mbar_wait(&bar_out2full, it & 1u);
      fence_after_sync();
      T256_STAMP();
      {
        const uint64_t w0 = (uint64_t)((a.seq0 * 32 + grow) * 256 + part * 64) >> 2;
        const uint32_t wlo = (uint32_t)w0, xhi = (uint32_t)(w0 >> 32) * 0x85EBCA6Bu;
        const float *b2 = p_b2 + part * 64;
        uint8_t *u2g = a.u2_img ? a.u2_img + (size_t)tile * T256_TILE_IMG : nullptr;
        float s1 = 0.f;
#pragma unroll
        for (int cb = 0; cb < 64; cb += 16) {
          float f[16];
          tmem_ld16(t_out + lane_off + (uint32_t)(part * 64 + cb), f);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            float m0 = 1.f, m1 = 1.f, m2 = 1.f, m3 = 1.f;
            if (a.d2.thr) {
              uint32_t lo, hi;
              hash_quad((wlo + (uint32_t)((cb + j) >> 2)) ^ xhi, a.d2.key, lo, hi);
              m0 = ((lo & 0xFFFFu) >= a.d2.thr) ? a.d2.scale : 0.f; m1 = ((lo >> 16) >= a.d2.thr) ? a.d2.scale : 0.f;
              m2 = ((hi & 0xFFFFu) >= a.d2.thr) ? a.d2.scale : 0.f; m3 = ((hi >> 16) >= a.d2.thr) ? a.d2.scale : 0.f;
            }
            u[cb + j] += (f[j] + b2[cb + j]) * m0;
            u[cb + j + 1] += (f[j + 1] + b2[cb + j + 1]) * m1;
            u[cb + j + 2] += (f[j + 2] + b2[cb + j + 2]) * m2;
            u[cb + j + 3] += (f[j + 3] + b2[cb + j + 3]) * m3;
            s1 += (u[cb + j] + u[cb + j + 1]) + (u[cb + j + 2] + u[cb + j + 3]);
          }
          if (u2g) {
#pragma unroll
            for (int j = 0; j < 16; j += 8)
              *reinterpret_cast<uint4 *>(u2g + kmajor_off(row, part * 64 + cb + j, 128)) =
                  make_uint4(pack_bf16(u[cb + j], u[cb + j + 1]), pack_bf16(u[cb + j + 2], u[cb + j + 3]), pack_bf16(u[cb + j + 4], u[cb + j + 5]),
                             pack_bf16(u[cb + j + 6], u[cb + j + 7]));
          }
        }
        T256_STAMP();
        sStatA[row * 4 + part] = s1;
        named_bar_sync(1, T256_CTHREADS);
        const float4 sa = *reinterpret_cast<const float4 *>(sStatA + row * 4);
        const float mu = ((sa.x + sa.y) + (sa.z + sa.w)) * (1.f / 256);
        float qq = 0.f;
#pragma unroll
        for (int c = 0; c < 64; ++c) { u[c] -= mu; qq = fmaf(u[c], u[c], qq); }
        sStatB[row * 4 + part] = qq;
        named_bar_sync(1, T256_CTHREADS);
        T256_STAMP();
        const float4 sb = *reinterpret_cast<const float4 *>(sStatB + row * 4);
        const float rs = rsqrtf(((sb.x + sb.y) + (sb.z + sb.w)) * (1.f / 256) + LN_EPS);
        if (part == 0 && a.ln2_stat) a.ln2_stat[grow] = make_float2(mu, rs);
        const float *g2 = p_g2 + part * 64, *be2 = p_be2 + part * 64;
        uint8_t *xog = a.x_img_out + (size_t)tile * T256_TILE_IMG;
#pragma unroll
        for (int c = 0; c < 64; c += 8) {
#pragma unroll
          for (int j = 0; j < 8; ++j) u[c + j] = u[c + j] * rs * g2[c + j] + be2[c + j];
          *reinterpret_cast<uint4 *>(xog + kmajor_off(row, part * 64 + c, 128)) =
                make_uint4(pack_bf16(u[c], u[c + 1]), pack_bf16(u[c + 2], u[c + 3]), pack_bf16(u[c + 4], u[c + 5]), pack_bf16(u[c + 6], u[c + 7]));
        }
      }
      fence_before_sync();
      if (tid == 0) tma_store_wait_read();
      named_bar_sync(1, T256_CTHREADS);
      T256_STAMP();
    }
    if (tid == 0) tma_store_wait_all();
  }
  __syncwarp();
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}


// =============================================================================================
// micro-benchmark: back-to-back UMMAs (M = 128, K = 16 each) from shared-memory operands in the
// canonical no-swizzle layout; out[0] = clocks per UMMA seen by the issuing thread (incl. completion)
// =============================================================================================
__global__ void __launch_bounds__(128, 1) t256_umma_rate_kernel(int N, int n_mma, int ksteps, int mode, float *out) {
  // mode bit 0: commit to an mbarrier after every block of `ksteps` UMMAs ; bit 1: also wait on a (pre-completed) mbarrier
  // per block ; bit 2: the issuing warp has ONE thread (blockDim.x == 65) instead of an elected lane of a full warp
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar, bar2, bar3;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (128 + 256) * 256 * 2 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u + (uint32_t)i;
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); mbar_init(&bar3, 1); fence_mbar_init(); }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const bool single = (mode & 4) != 0;
  if (warp == (single ? 2 : 1)) {
    const uint32_t idesc = make_idesc_bf16(128, N);
    const uint64_t dAb = descA128(smem_u32(smem)), dBb = descB(smem_u32(smem) + 65536u, N);
    long long t0 = clock64();
    for (int i = 0; i < n_mma; i += ksteps) {
      if (mode & 2) { mbar_wait(&bar3, 1); fence_after_sync(); }      // fresh barrier: parity-1 wait passes immediately
      if (single || elect_one()) {
        for (int k = 0; k < ksteps; ++k)
          mma_bf16_ss(tmem, desc_adv(dAb, (uint32_t)k * 4096u), desc_adv(dBb, (uint32_t)k * (uint32_t)N * 32u), idesc, 1u);
        if (mode & 1) mma_commit(&bar2);
      }
      if (!single) __syncwarp();
    }
    if (single || elect_one()) mma_commit(&bar);
    if (!single) __syncwarp();
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if ((tid & 31) == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0) / (float)n_mma;
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}
int t256_debug_umma_rate(int N, int n_mma, int ksteps, float *out, cudaStream_t st) {
  const int mode = ksteps >> 8;
  ksteps &= 255;
  GT_CUDA(cudaFuncSetAttribute(t256_umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  t256_umma_rate_kernel<<<148, (mode & 4) ? 65 : 128, 196608 + 1024, st>>>(N, n_mma, ksteps, mode, out);
  GT_CUDA(cudaGetLastError());
  return 0;
}

int t256_num_sms() {
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
static int t256_launch_fwd(const T256Args &a_in, int grid, cudaStream_t st) {
  static T256Dbg dbg;
  static const uint32_t stagger = getenv("GT_T256_STAGGER_FWD") ? (uint32_t)atoi(getenv("GT_T256_STAGGER_FWD")) : 0u;
  T256Args a = a_in;
  a.stagger = a.n_tiles >= 16 * grid ? stagger : 0u;
  const bool d = dbg.arm(a, st);
  if (drop_args_devstep(a)) {                // graph replay: dropout keys derived on the device from the step counter
    GT_CUDA(cudaFuncSetAttribute(t256_layer_fwd_kernel<DH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T256FwdSmem::total));
    { LaunchScope _ls(KC_TC_LAYER_FWD, st);
      t256_layer_fwd_kernel<DH, true><<<grid, T256_THREADS, T256FwdSmem::total, st>>>(a); }
    GT_CUDA(cudaGetLastError());
    return 0;
  }
  GT_CUDA(cudaFuncSetAttribute(t256_layer_fwd_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T256FwdSmem::total));
  { LaunchScope _ls(KC_TC_LAYER_FWD, st);
    t256_layer_fwd_kernel<DH><<<grid, T256_THREADS, T256FwdSmem::total, st>>>(a); }
  GT_CUDA(cudaGetLastError());
  if (d) dbg.report("t256 fwd", st);
  return 0;
}

int t256_layer_fwd(const T256Args &a, cudaStream_t st) {
  GT_CHECK(a.F % 64 == 0 && a.F >= 64 && a.F <= 512, "t256_layer_fwd: dim_feedforward not supported");
  const int grid = a.n_tiles < t256_num_sms() ? a.n_tiles : t256_num_sms();
  switch (a.dh) {
    case 16: return t256_launch_fwd<16>(a, grid, st);
    case 32: return t256_launch_fwd<32>(a, grid, st);
    case 128: return t256_launch_fwd<128>(a, grid, st);
    default: GT_FAIL("t256_layer_fwd: head dim not instantiated");
  }
}


}  // namespace gt
