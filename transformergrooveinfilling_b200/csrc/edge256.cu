// edge256.cu — the two ends of the d_model = 256 fused path on the tile-native layouts of tc256.cuh (bf16 K-major images
// for activations, tiled fp32 for gradients), replacing the row-major fp32 GEMM / LayerNorm / element-wise / layout-conversion
// launches that made 16 % of a C4 training step:
//
//   stem forward : src [M, E] -> x0 image = dropout(relu(src W_in^T + b) + pe)            BGT/models/io_layers.py:17-22
//   tail forward : x_L image -> LayerNorm -> 27 logits -> h | sigmoid | 0.5 tanh (+ calculate_loss, BGT/models/train.py:9-40:
//                  metric partial sums and dL/dlogits)                                    encoder.py:8-16, io_layers.py:36-48
//   tail backward: rows kernel: dz = dlogits W_out, LayerNorm backward -> dx (tiled fp32); dl' = dlogits * rstd as a bf16
//                  [128 x 32] image per tile.  ALL parameter gradients of the tail follow from ONE tensor-core contraction
//                  over the tokens, T[c][j] = sum_r x[r][c] dl'[r][j] (t256_wgrad_kernel: the saved x_L image against the dl'
//                  image), plus two 27-vectors s[j] = sum_r dl'[r][j] mu[r] and sb[j] = sum_r dlogits[r][j]:
//                    dW_out[j][c] = g[c] (T[c][j] - s[j]) + b[c] sb[j]      db_out[j] = sb[j]
//                    dgamma[c]    = sum_j W_out[j][c] (T[c][j] - s[j])      dbeta[c]  = sum_j W_out[j][c] sb[j]
//   stem backward: rows kernel: g = dx0 * dropmask * (r > 0) (r recomputed from src) as a bf16 image + the src tile as a
//                  bf16 [128 x 32] image with a column of ones; dW_in | db_in = g^T [src | 1] on the tensor core.
//
// Row kernels: one persistent CTA per SM, 512 threads, thread = (token row r = tid & 127, column part = tid >> 7: 64 columns);
// a warp's 32 rows make every image chunk access (16 B per row) and every tiled-fp32 access (float4 per row) 512 contiguous bytes.
#include "tc256_dev.cuh"

namespace gt {

constexpr int E256_THREADS = 512;

__device__ __forceinline__ float e256_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void e256_unpack8(const uint4 &v, float (&f)[8]) {
  f[0] = bf16lo(v.x); f[1] = bf16hi(v.x); f[2] = bf16lo(v.y); f[3] = bf16hi(v.y);
  f[4] = bf16lo(v.z); f[5] = bf16hi(v.z); f[6] = bf16lo(v.w); f[7] = bf16hi(v.w);
}
// keep decisions of the 8 consecutive elements starting at element index e (multiple of 8) of one dropout site
__device__ __forceinline__ void e256_drop8(const Drop &d, uint64_t e, float (&v)[8]) {
  if (d.thr == 0) return;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const uint64_t w = (e >> 2) + (uint64_t)q;
    uint32_t lo, hi;
    hash_quad((uint32_t)w ^ ((uint32_t)(w >> 32) * 0x85EBCA6Bu), d.key, lo, hi);
    v[4 * q] = ((lo & 0xFFFFu) >= d.thr) ? v[4 * q] * d.scale : 0.f;
    v[4 * q + 1] = ((lo >> 16) >= d.thr) ? v[4 * q + 1] * d.scale : 0.f;
    v[4 * q + 2] = ((hi & 0xFFFFu) >= d.thr) ? v[4 * q + 2] * d.scale : 0.f;
    v[4 * q + 3] = ((hi >> 16) >= d.thr) ? v[4 * q + 3] * d.scale : 0.f;
  }
}

// ---- stem -----------------------------------------------------------------------------------------------------------
// shared: WT [E][256] | bias [256] | peT [256][32] (forward) | sx [128][EP]   (EP odd: conflict-free per-row reads)
template <int E> struct StemCfg { static constexpr int EP = E | 1; };

template <int E, bool BWD>
__global__ void __launch_bounds__(E256_THREADS, 1) stem256_kernel(const float *__restrict__ src, const float *__restrict__ W,
                                                                  const float *__restrict__ b, const float *__restrict__ pe,
                                                                  uint8_t *__restrict__ x_img, const float *__restrict__ dx_tiled,
                                                                  uint8_t *__restrict__ g_img, uint8_t *__restrict__ src_img, int64_t M,
                                                                  int n_tiles, Drop drop, int64_t row0) {
  drop_resolve(drop);                                // graph replay: key from the device step counter
  constexpr int EP = StemCfg<E>::EP;
  extern __shared__ __align__(16) float esm[];
  float *sWT = esm, *sB = sWT + E * 256, *sPe = sB + 256, *sx = sPe + (BWD ? 0 : 256 * 32);
  const int tid = threadIdx.x, r = tid & 127, part = tid >> 7;
  for (int i = tid; i < 256 * E; i += E256_THREADS) sWT[(i % E) * 256 + i / E] = W[i];
  if (tid < 256) sB[tid] = b[tid];
  if (!BWD)
    for (int i = tid; i < 32 * 256; i += E256_THREADS) sPe[(i & 255) * 32 + (i >> 8)] = pe[i];      // pe [32][256] -> [c][t]
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    __syncthreads();
    const int64_t tok0 = (int64_t)tile * 128;
    const int64_t nvalid = min((int64_t)128, M - tok0);
    for (int i = tid; i < 128 * E; i += E256_THREADS) sx[(i / E) * EP + i % E] = i < nvalid * E ? __ldg(src + tok0 * E + i) : 0.f;
    __syncthreads();
    const bool valid = r < nvalid;
    float xr[E];
#pragma unroll
    for (int k = 0; k < E; ++k) xr[k] = sx[r * EP + k];
    const uint64_t e_row = (uint64_t)((row0 + tok0 + r) * 256);
    if (BWD && part == 0) {
      // [src | 1 | 0..] as a bf16 [128 x 32] image: the B operand of dW_in | db_in = g^T [src | 1]
      float v[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) v[k] = k < E ? xr[k] : (k == E && valid ? 1.f : 0.f);
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8)
        *reinterpret_cast<uint4 *>(src_img + (size_t)tile * 8192 + kmajor_off(r, c8 * 8, 128)) =
            make_uint4(pack_bf16(v[c8 * 8], v[c8 * 8 + 1]), pack_bf16(v[c8 * 8 + 2], v[c8 * 8 + 3]), pack_bf16(v[c8 * 8 + 4], v[c8 * 8 + 5]),
                       pack_bf16(v[c8 * 8 + 6], v[c8 * 8 + 7]));
    }
#pragma unroll 1
    for (int c8 = 0; c8 < 8; ++c8) {
      const int c = part * 64 + c8 * 8;
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = sB[c + j];
#pragma unroll
      for (int k = 0; k < E; ++k) {
        const float4 w0 = *reinterpret_cast<const float4 *>(sWT + k * 256 + c), w1 = *reinterpret_cast<const float4 *>(sWT + k * 256 + c + 4);
        acc[0] = fmaf(xr[k], w0.x, acc[0]); acc[1] = fmaf(xr[k], w0.y, acc[1]); acc[2] = fmaf(xr[k], w0.z, acc[2]); acc[3] = fmaf(xr[k], w0.w, acc[3]);
        acc[4] = fmaf(xr[k], w1.x, acc[4]); acc[5] = fmaf(xr[k], w1.y, acc[5]); acc[6] = fmaf(xr[k], w1.z, acc[6]); acc[7] = fmaf(xr[k], w1.w, acc[7]);
      }
      float v[8];
      if (!BWD) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = valid ? fmaxf(acc[j], 0.f) + sPe[(c + j) * 32 + (r & 31)] : 0.f;
        e256_drop8(drop, e_row + (uint64_t)c, v);
        *reinterpret_cast<uint4 *>(x_img + (size_t)tile * T256_TILE_IMG + kmajor_off(r, c, 128)) =
            make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
      } else {
        const float *dxt = dx_tiled + (size_t)tile * T256_TILE_F32;
        const float4 d0 = *reinterpret_cast<const float4 *>(dxt + ((size_t)(c >> 2) * 128 + r) * 4);
        const float4 d1 = *reinterpret_cast<const float4 *>(dxt + ((size_t)((c >> 2) + 1) * 128 + r) * 4);
        v[0] = d0.x; v[1] = d0.y; v[2] = d0.z; v[3] = d0.w; v[4] = d1.x; v[5] = d1.y; v[6] = d1.z; v[7] = d1.w;
        e256_drop8(drop, e_row + (uint64_t)c, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (valid && acc[j] > 0.f) ? v[j] : 0.f;
        *reinterpret_cast<uint4 *>(g_img + (size_t)tile * T256_TILE_IMG + kmajor_off(r, c, 128)) =
            make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
      }
    }
  }
}

static int e256_grid(int n_tiles) { const int s = t256_num_sms(); return n_tiles < s ? n_tiles : s; }

template <int E, bool BWD>
static int stem256_launch(const float *src, const float *W, const float *b, const float *pe, uint8_t *x_img, const float *dx_tiled,
                          uint8_t *g_img, uint8_t *src_img, int64_t M, int n_tiles, const Drop &drop, int64_t row0, cudaStream_t st) {
  const size_t smem = (size_t)(E * 256 + 256 + (BWD ? 0 : 256 * 32) + 128 * StemCfg<E>::EP) * sizeof(float);
  GT_CUDA(cudaFuncSetAttribute(stem256_kernel<E, BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  { LaunchScope _ls(KC_TC_INPUT, st);
    stem256_kernel<E, BWD><<<e256_grid(n_tiles), E256_THREADS, smem, st>>>(src, W, b, pe, x_img, dx_tiled, g_img, src_img, M, n_tiles, drop, row0); }
  GT_CUDA(cudaGetLastError());
  return 0;
}
int edge256_stem_fwd(const float *src, int E, const float *W, const float *b, const float *pe, uint8_t *x_img, int64_t M, int n_tiles,
                     const Drop &drop, int64_t row0, cudaStream_t st) {
  if (E == 16) return stem256_launch<16, false>(src, W, b, pe, x_img, nullptr, nullptr, nullptr, M, n_tiles, drop, row0, st);
  if (E == 27) return stem256_launch<27, false>(src, W, b, pe, x_img, nullptr, nullptr, nullptr, M, n_tiles, drop, row0, st);
  GT_FAIL("edge256: embedding_size_src must be 16 or 27");
}
int edge256_stem_bwd_rows(const float *dx_tiled, const float *src, int E, const float *W, const float *b, uint8_t *g_img, uint8_t *src_img,
                          int64_t M, int n_tiles, const Drop &drop, int64_t row0, cudaStream_t st) {
  if (E == 16) return stem256_launch<16, true>(src, W, b, nullptr, nullptr, dx_tiled, g_img, src_img, M, n_tiles, drop, row0, st);
  if (E == 27) return stem256_launch<27, true>(src, W, b, nullptr, nullptr, dx_tiled, g_img, src_img, M, n_tiles, drop, row0, st);
  GT_FAIL("edge256: embedding_size_src must be 16 or 27");
}

// ---- tail forward (+ optional fused loss) ---------------------------------------------------------------------------
// shared: WT [256][28] | gamma [256] | beta [256] | bias [28] + 4 | sStat [128][4][2] | sLog [3][128][28] | red [4][16]
struct TailSmem {
  static constexpr int wt = 0, gm = wt + 256 * 28, be = gm + 256, bo = be + 256, stat = bo + 32, logit = stat + 128 * 8, red = logit + 3 * 128 * 28,
                       total = red + 64;
};

__global__ void __launch_bounds__(E256_THREADS, 1) tail256_fwd_kernel(const uint8_t *__restrict__ x_img, const float *__restrict__ gamma,
                                                                      const float *__restrict__ beta, const float *__restrict__ Wout,
                                                                      const float *__restrict__ bout, float *__restrict__ hvo, float *mean,
                                                                      float *rstd, int64_t M, int n_tiles, float thres,
                                                                      const float *__restrict__ y, float penalty, float gscale,
                                                                      float *__restrict__ dlog, float *partials) {
  extern __shared__ __align__(16) float esm[];
  using S = TailSmem;
  float *sWT = esm + S::wt, *sG = esm + S::gm, *sBe = esm + S::be, *sBo = esm + S::bo, *sStat = esm + S::stat, *sLog = esm + S::logit,
        *sRed = esm + S::red;
  const int tid = threadIdx.x, r = tid & 127, part = tid >> 7, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 256 * 28; i += E256_THREADS) { const int c = i / 28, j = i % 28; sWT[i] = j < 27 ? Wout[j * 256 + c] : 0.f; }
  if (tid < 256) { sG[tid] = gamma[tid]; sBe[tid] = beta[tid]; }
  if (tid < 32) sBo[tid] = tid < 27 ? bout[tid] : 0.f;
  float a_bce = 0.f, a_mv = 0.f, a_mo = 0.f, a_ok = 0.f;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    __syncthreads();
    const int64_t grow = (int64_t)tile * 128 + r;
    const bool valid = grow < M;
    const uint8_t *xt = x_img + (size_t)tile * T256_TILE_IMG;
    uint4 ch[8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c8 = 0; c8 < 8; ++c8) {
      ch[c8] = *reinterpret_cast<const uint4 *>(xt + kmajor_off(r, part * 64 + c8 * 8, 128));
      float f[8];
      e256_unpack8(ch[c8], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) { s1 += f[j]; s2 = fmaf(f[j], f[j], s2); }
    }
    sStat[(r * 4 + part) * 2] = s1; sStat[(r * 4 + part) * 2 + 1] = s2;
    __syncthreads();
    float mu, rs;
    {
      const float4 a = *reinterpret_cast<const float4 *>(sStat + r * 8), b4 = *reinterpret_cast<const float4 *>(sStat + r * 8 + 4);
      mu = ((a.x + a.z) + (b4.x + b4.z)) * (1.f / 256);
      const float var = fmaxf(((a.y + a.w) + (b4.y + b4.w)) * (1.f / 256) - mu * mu, 0.f);
      rs = rsqrtf(var + LN_EPS);
    }
    float acc[28];
#pragma unroll
    for (int j = 0; j < 28; ++j) acc[j] = 0.f;
#pragma unroll
    for (int c8 = 0; c8 < 8; ++c8) {
      float f[8];
      e256_unpack8(ch[c8], f);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int c = part * 64 + c8 * 8 + q;
        const float z = (f[q] - mu) * rs * sG[c] + sBe[c];
#pragma unroll
        for (int j = 0; j < 28; j += 4) {
          const float4 w = *reinterpret_cast<const float4 *>(sWT + c * 28 + j);
          acc[j] = fmaf(z, w.x, acc[j]); acc[j + 1] = fmaf(z, w.y, acc[j + 1]); acc[j + 2] = fmaf(z, w.z, acc[j + 2]); acc[j + 3] = fmaf(z, w.w, acc[j + 3]);
        }
      }
    }
    if (part > 0) {
#pragma unroll
      for (int j = 0; j < 28; j += 4)
        *reinterpret_cast<float4 *>(sLog + ((part - 1) * 128 + r) * 28 + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
    }
    __syncthreads();
    if (part == 0 && valid) {
#pragma unroll
      for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int j = 0; j < 28; j += 4) {
          const float4 v = *reinterpret_cast<const float4 *>(sLog + (p * 128 + r) * 28 + j);
          acc[j] += v.x; acc[j + 1] += v.y; acc[j + 2] += v.z; acc[j + 3] += v.w;
        }
      if (mean != nullptr) { mean[grow] = mu; rstd[grow] = rs; }
      float *out = hvo + grow * 27;
      const float *yt = y != nullptr ? y + grow * 27 : nullptr;
      float *dl = dlog != nullptr ? dlog + grow * 27 : nullptr;
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const float h = acc[k] + sBo[k], lv = acc[9 + k] + sBo[9 + k], lo = acc[18 + k] + sBo[18 + k];
        const float sg = 1.f / (1.f + expf(-h)), v = 1.f / (1.f + expf(-lv)), o = 0.5f * tanhf(lo);
        out[k] = thres >= 0.f ? (sg > thres ? 1.f : 0.f) : h;
        out[9 + k] = v;
        out[18 + k] = o;
        if (yt != nullptr) {
          const float yh = yt[k], yv = yt[9 + k], yo = yt[18 + k];
          const float w = (yh == 1.f) ? 1.f : penalty;
          a_bce = fmaf(fmaxf(h, 0.f) - h * yh + log1pf(expf(-fabsf(h))), w, a_bce);
          const float dv = v - yv, dof = o - yo;
          a_mv = fmaf(dv * dv, w, a_mv);
          a_mo = fmaf(dof * dof, w, a_mo);
          a_ok += ((sg > 0.5f ? 1.f : 0.f) == yh) ? 1.f : 0.f;
          dl[k] = gscale * w * (sg - yh);
          dl[9 + k] = gscale * 2.f * w * dv * v * (1.f - v);
          dl[18 + k] = gscale * 2.f * w * dof * (0.5f - 2.f * o * o);
        }
      }
    }
  }
  if (partials != nullptr) {
    __syncthreads();
    const float q0 = e256_warp_sum(a_bce), q1 = e256_warp_sum(a_mv), q2 = e256_warp_sum(a_mo), q3 = e256_warp_sum(a_ok);
    if (lane == 0) { sRed[warp] = q0; sRed[16 + warp] = q1; sRed[32 + warp] = q2; sRed[48 + warp] = q3; }
    __syncthreads();
    if (tid < 4) {
      float s = 0.f;
      for (int i = 0; i < 16; ++i) s += sRed[tid * 16 + i];
      partials[(int64_t)blockIdx.x * 4 + tid] = s;
    }
  }
}

int64_t edge256_loss_partials() { return 4 * 160; }

int edge256_tail_fwd(const uint8_t *x_img, const float *gamma, const float *beta, const float *Wout, const float *bout, float *hvo, float *mean,
                     float *rstd, int64_t M, int n_tiles, float thres, const float *y, float penalty, float *dlog, float *partials,
                     float *metrics6, cudaStream_t st) {
  const size_t smem = (size_t)TailSmem::total * sizeof(float);
  GT_CUDA(cudaFuncSetAttribute(tail256_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = e256_grid(n_tiles);
  { LaunchScope _ls(KC_TC_HEAD, st);
    tail256_fwd_kernel<<<grid, E256_THREADS, smem, st>>>(x_img, gamma, beta, Wout, bout, hvo, mean, rstd, M, n_tiles, thres, y, penalty,
                                                        y ? 1.f / (float)M : 0.f, y ? dlog : nullptr, y ? partials : nullptr); }
  GT_CUDA(cudaGetLastError());
  if (y != nullptr) return loss_finalize(partials, grid, M, metrics6, st);
  return 0;
}

// ---- tail backward: rows kernel -------------------------------------------------------------------------------------
// shared: WT [256][28] | gamma [256] | sStat [128][4][2]
__global__ void __launch_bounds__(E256_THREADS, 1) tail256_bwd_rows_kernel(const float *__restrict__ d_in, const float *__restrict__ hvo,
                                                                           const uint8_t *__restrict__ x_img, const float *__restrict__ mean,
                                                                           const float *__restrict__ rstd, const float *__restrict__ gamma,
                                                                           const float *__restrict__ Wout, float *__restrict__ dx_tiled,
                                                                           uint8_t *__restrict__ dl_img, int64_t M, int n_tiles) {
  extern __shared__ __align__(16) float esm[];
  float *sWT = esm, *sG = sWT + 256 * 28, *sStat = sG + 256;
  const int tid = threadIdx.x, r = tid & 127, part = tid >> 7;
  for (int i = tid; i < 256 * 28; i += E256_THREADS) { const int c = i / 28, j = i % 28; sWT[i] = j < 27 ? Wout[j * 256 + c] : 0.f; }
  if (tid < 256) sG[tid] = gamma[tid];
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    __syncthreads();
    const int64_t grow = (int64_t)tile * 128 + r;
    const bool valid = grow < M;
    float dl[32];                       // 27 channels, zero padded to the 32 columns of the dl' image
#pragma unroll
    for (int j = 0; j < 32; ++j) dl[j] = 0.f;
    float mu = 0.f, rs = 0.f;
    if (valid) {
      mu = __ldg(mean + grow); rs = __ldg(rstd + grow);
#pragma unroll
      for (int j = 0; j < 27; ++j) dl[j] = __ldg(d_in + grow * 27 + j);
      if (hvo != nullptr) {
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          const float v = __ldg(hvo + grow * 27 + 9 + k), o = __ldg(hvo + grow * 27 + 18 + k);
          dl[9 + k] *= v * (1.f - v);
          dl[18 + k] *= 0.5f - 2.f * o * o;
        }
      }
    }
    if (part == 0) {
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8)
        *reinterpret_cast<uint4 *>(dl_img + (size_t)tile * 8192 + kmajor_off(r, c8 * 8, 128)) =
            make_uint4(pack_bf16(dl[c8 * 8] * rs, dl[c8 * 8 + 1] * rs), pack_bf16(dl[c8 * 8 + 2] * rs, dl[c8 * 8 + 3] * rs),
                       pack_bf16(dl[c8 * 8 + 4] * rs, dl[c8 * 8 + 5] * rs), pack_bf16(dl[c8 * 8 + 6] * rs, dl[c8 * 8 + 7] * rs));
    }
    const uint8_t *xt = x_img + (size_t)tile * T256_TILE_IMG;
    float gd[64];                       // dz * gamma of this thread's 64 columns
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c8 = 0; c8 < 8; ++c8) {
      float f[8];
      e256_unpack8(*reinterpret_cast<const uint4 *>(xt + kmajor_off(r, part * 64 + c8 * 8, 128)), f);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int c = part * 64 + c8 * 8 + q;
        float d0 = 0.f, d1 = 0.f;
#pragma unroll
        for (int j = 0; j < 28; j += 4) {
          const float4 w = *reinterpret_cast<const float4 *>(sWT + c * 28 + j);
          d0 = fmaf(dl[j], w.x, d0); d1 = fmaf(dl[j + 1], w.y, d1); d0 = fmaf(dl[j + 2], w.z, d0); d1 = fmaf(dl[j + 3], w.w, d1);
        }
        const float g = (d0 + d1) * sG[c];
        gd[c8 * 8 + q] = g;
        s1 += g;
        s2 = fmaf(g, (f[q] - mu) * rs, s2);
      }
    }
    sStat[(r * 4 + part) * 2] = s1; sStat[(r * 4 + part) * 2 + 1] = s2;
    __syncthreads();
    {
      const float4 a = *reinterpret_cast<const float4 *>(sStat + r * 8), b4 = *reinterpret_cast<const float4 *>(sStat + r * 8 + 4);
      s1 = ((a.x + a.z) + (b4.x + b4.z)) * (1.f / 256);
      s2 = ((a.y + a.w) + (b4.y + b4.w)) * (1.f / 256);
    }
    float *dxt = dx_tiled + (size_t)tile * T256_TILE_F32;
#pragma unroll
    for (int c8 = 0; c8 < 8; ++c8) {
      float f[8], o[8];
      e256_unpack8(*reinterpret_cast<const uint4 *>(xt + kmajor_off(r, part * 64 + c8 * 8, 128)), f);
#pragma unroll
      for (int q = 0; q < 8; ++q) o[q] = valid ? (gd[c8 * 8 + q] - s1 - (f[q] - mu) * rs * s2) * rs : 0.f;
      const int c = part * 64 + c8 * 8;
      *reinterpret_cast<float4 *>(dxt + ((size_t)(c >> 2) * 128 + r) * 4) = make_float4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<float4 *>(dxt + ((size_t)((c >> 2) + 1) * 128 + r) * 4) = make_float4(o[4], o[5], o[6], o[7]);
    }
  }
}

// column sums of the [M, 27] arrays: sums[j] = sum_r dl[r][j] ; sums[32 + j] = sum_r dl[r][j] rstd[r] mean[r]   (lane = channel)
__global__ void __launch_bounds__(256) tail256_colsums_kernel(const float *__restrict__ d_in, const float *__restrict__ hvo,
                                                              const float *__restrict__ mean, const float *__restrict__ rstd, int64_t M,
                                                              float *sums) {
  __shared__ float sacc[64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < 64) sacc[threadIdx.x] = 0.f;
  __syncthreads();
  const int role = lane / 9;
  float a0 = 0.f, a1 = 0.f;
  if (lane < 27)
    for (int64_t row = (int64_t)blockIdx.x * 8 + warp; row < M; row += (int64_t)gridDim.x * 8) {
      float d = __ldg(d_in + row * 27 + lane);
      if (hvo != nullptr && role > 0) {
        const float a = __ldg(hvo + row * 27 + lane);
        d *= role == 1 ? a * (1.f - a) : (0.5f - 2.f * a * a);
      }
      a0 += d;
      a1 = fmaf(d, __ldg(rstd + row) * __ldg(mean + row), a1);
    }
  if (lane < 27) { atomicAdd(&sacc[lane], a0); atomicAdd(&sacc[32 + lane], a1); }
  __syncthreads();
  if (threadIdx.x < 64 && (threadIdx.x & 31) < 27) atomicAdd(sums + threadIdx.x, sacc[threadIdx.x]);
}

// T [256][32] (T[c][j] = sum_r x[r][c] dl'[r][j]), sums [64] -> parameter gradients of the final LayerNorm and the head
__global__ void __launch_bounds__(256) tail256_finalize_kernel(const float *__restrict__ Tm, const float *__restrict__ sums,
                                                               const float *__restrict__ gamma, const float *__restrict__ beta,
                                                               const float *__restrict__ Wout, float *gW, float *gb, float *gg, float *gbe) {
  const int c = threadIdx.x;
  const float g = gamma[c], b = beta[c];
  float dg = 0.f, dbt = 0.f;
  for (int j = 0; j < 27; ++j) {
    const float t = Tm[c * 32 + j] - sums[32 + j], sb = sums[j], w = Wout[j * 256 + c];
    gW[j * 256 + c] += g * t + b * sb;
    dg = fmaf(w, t, dg);
    dbt = fmaf(w, sb, dbt);
  }
  gg[c] += dg;
  gbe[c] += dbt;
  if (c < 27) gb[c] += sums[c];
}

// Tin [256][32] = g^T [src | 1] -> dW_in [256][E], db_in [256]
__global__ void __launch_bounds__(256) stem256_finalize_kernel(const float *__restrict__ Tin, int E, float *gW, float *gb) {
  const int c = threadIdx.x;
  for (int k = 0; k < E; ++k) gW[c * E + k] += Tin[c * 32 + k];
  gb[c] += Tin[c * 32 + E];
}

int edge256_tail_bwd(const float *d_in, const float *hvo, const uint8_t *x_img, const float *mean, const float *rstd, const float *gamma,
                     const float *beta, const float *Wout, float *dx_tiled, uint8_t *dl_img, float *scratch /* [256*32 + 64] */,
                     void *job_buf, float *gW, float *gb, float *gg, float *gbe, int64_t M, int n_tiles, cudaStream_t st) {
  const size_t smem = (size_t)(256 * 28 + 256 + 128 * 8) * sizeof(float);
  GT_CUDA(cudaFuncSetAttribute(tail256_bwd_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  GT_CUDA(cudaMemsetAsync(scratch, 0, (size_t)(256 * 32 + 64) * sizeof(float), st));
  { LaunchScope _ls(KC_TC_HEAD, st);
    tail256_bwd_rows_kernel<<<e256_grid(n_tiles), E256_THREADS, smem, st>>>(d_in, hvo, x_img, mean, rstd, gamma, Wout, dx_tiled, dl_img, M, n_tiles); }
  GT_CUDA(cudaGetLastError());
  { LaunchScope _ls(KC_TC_HEAD, st);
    const int64_t blocks = (M + 8 * 64 - 1) / (8 * 64);
    tail256_colsums_kernel<<<(unsigned)(blocks < 1 ? 1 : (blocks > 592 ? 592 : blocks)), 256, 0, st>>>(d_in, hvo, mean, rstd, M, scratch + 256 * 32); }
  GT_CUDA(cudaGetLastError());
  GT_TRY(t256_wgrad_pair(x_img, dl_img, 32, scratch, n_tiles, job_buf, st));
  { LaunchScope _ls(KC_TC_HEAD, st);
    tail256_finalize_kernel<<<1, 256, 0, st>>>(scratch, scratch + 256 * 32, gamma, beta, Wout, gW, gb, gg, gbe); }
  GT_CUDA(cudaGetLastError());
  return 0;
}

int edge256_stem_bwd(const float *dx_tiled, const float *src, int E, const float *W, const float *b, uint8_t *g_img, uint8_t *src_img,
                     float *scratch /* [256*32] */, void *job_buf, float *gW, float *gb, int64_t M, int n_tiles, const Drop &drop, int64_t row0,
                     cudaStream_t st) {
  GT_CUDA(cudaMemsetAsync(scratch, 0, (size_t)(256 * 32) * sizeof(float), st));
  GT_TRY(edge256_stem_bwd_rows(dx_tiled, src, E, W, b, g_img, src_img, M, n_tiles, drop, row0, st));
  GT_TRY(t256_wgrad_pair(g_img, src_img, 32, scratch, n_tiles, job_buf, st));
  { LaunchScope _ls(KC_TC_INPUT, st);
    stem256_finalize_kernel<<<1, 256, 0, st>>>(scratch, E, gW, gb); }
  GT_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace gt
