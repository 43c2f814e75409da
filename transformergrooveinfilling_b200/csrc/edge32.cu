// edge32.cu — the two ends of the d_model = 32 fused path, each as ONE HBM-bound kernel per direction:
//
//   stem  forward : x0 = dropout(relu(src W_in^T + b) + pe)                    BGT/models/io_layers.py:17-22, utils.py:49-50
//   stem  backward: g = dx0 * dropmask * (r > 0) (r recomputed) ; dW_in += g^T src ; db_in += colsum(g)
//   tail  forward : z = LayerNorm(x_L) ; logits = z W_out^T + b ; h | sigmoid(v) | 0.5 tanh(o)      encoder.py:8-16, io_layers.py:36-48
//                   (+ fused calculate_loss, BGT/models/train.py:9-40: loss partial sums and dL/dlogits, in the train step)
//   tail  backward: dW_out += dlogits^T z ; db_out ; dz = dlogits W_out ; LayerNorm backward -> dx_L, dgamma, dbeta
//
// They replace 17 launches of the generic fp32 kernels (5 GEMMs with K or N of 16..32, 2 colsums, LayerNorm forward/backward,
// 4 element-wise passes, the loss partial pass) whose intermediates (r0, z, dlogits, dz, g0: [tokens x 32] fp32 each) made
// them 15 % of a C2 training step.  Algorithmic bytes per token (fp32): stem fwd 4 E + 128, stem bwd 4 E + 128, tail fwd
// 128 + 108 (+ 108 y + 108 dlogits with the loss), tail bwd 108 + 128 + 128.
//
// Mapping (all four): a warp owns 4 consecutive tokens per pass; lane = feature column c (d_model = 32) for the row-wise
// parts and lane = output channel j (27 of 32 lanes) for the head parts; the token vector crosses between the two roles
// through 512 B of shared memory per warp.  Every global access of a warp is one contiguous 108 / 128 byte row.
#include "common.cuh"

namespace gt {

constexpr int EG_WARPS = 8, EG_TOK = 4, EG_MAX_BLOCKS = 148 * 8;

__device__ __forceinline__ float eg_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

static int eg_blocks(int64_t M) {
  int64_t b = (M + EG_WARPS * EG_TOK - 1) / (EG_WARPS * EG_TOK);
  return (int)(b < 1 ? 1 : (b > EG_MAX_BLOCKS ? EG_MAX_BLOCKS : b));
}
int64_t edge32_loss_partials(int64_t) { return 4 * (int64_t)EG_MAX_BLOCKS; }

// ---- stem -----------------------------------------------------------------------------------------------------------
template <int E>
struct StemSmem {
  static constexpr int EP = (E + 3) & ~3;       // padded source row (float4 reads)
};

template <int E, bool BWD>
__global__ void __launch_bounds__(EG_WARPS * 32) stem32_kernel(const float *__restrict__ src, const float *__restrict__ W,
                                                               const float *__restrict__ b, const float *__restrict__ pe,
                                                               float *__restrict__ x0, const float *__restrict__ dx0,
                                                               float *gW, float *gb, int64_t M, Drop drop, int64_t e0) {
  constexpr int EP = StemSmem<E>::EP;
  griddep_wait();                                    // launched with launch_pdl (common.cuh): nothing is read before the predecessor is done
  griddep_launch();
  drop_resolve(drop);                                // graph replay: key from the device step counter
  __shared__ __align__(16) float sx[EG_WARPS][EG_TOK][EP];
  __shared__ float sacc[BWD ? 32 * E + 32 : 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float w[EP];
#pragma unroll
  for (int k = 0; k < EP; ++k) w[k] = k < E ? W[lane * E + k] : 0.f;
  const float bias = b[lane];
  for (int i = lane; i < EG_TOK * EP; i += 32) (&sx[warp][0][0])[i] = 0.f;
  float acc[BWD ? EP : 1], accb = 0.f;
  if (BWD) {
#pragma unroll
    for (int k = 0; k < EP; ++k) acc[k] = 0.f;
    for (int i = threadIdx.x; i < 32 * E + 32; i += blockDim.x) sacc[i] = 0.f;
    __syncthreads();
  }
  __syncwarp();
  const int64_t n_groups = (M + EG_TOK - 1) / EG_TOK;
  const int64_t gstride = (int64_t)gridDim.x * EG_WARPS;
  // software pipeline: the NEXT group's source rows (EG_TOK * E contiguous floats: NS per lane) and gradient rows are requested
  // before the current group is processed, so their latency overlaps a group of arithmetic instead of opening every group
  constexpr int NS = (EG_TOK * E + 31) / 32;
  float ns[NS], ndv[BWD ? EG_TOK : 1];
  auto fetch = [&](int64_t gn) {
    const int64_t t0 = gn * EG_TOK;
    const int nt = gn < n_groups ? (int)min((int64_t)EG_TOK, M - t0) : 0;
#pragma unroll
    for (int u = 0; u < NS; ++u) ns[u] = (lane + 32 * u) < nt * E ? __ldg(src + t0 * E + lane + 32 * u) : 0.f;
    if (BWD) {
#pragma unroll
      for (int q = 0; q < EG_TOK; ++q) ndv[q] = q < nt ? __ldg(dx0 + (t0 + q) * 32 + lane) : 0.f;
    }
  };
  fetch((int64_t)blockIdx.x * EG_WARPS + warp);
  for (int64_t g = (int64_t)blockIdx.x * EG_WARPS + warp; g < n_groups; g += gstride) {
    const int64_t tok0 = g * EG_TOK;
    const int ntok = (int)min((int64_t)EG_TOK, M - tok0);
    // the group's source rows are contiguous in memory: coalesced loads, scattered into padded rows
#pragma unroll
    for (int u = 0; u < NS; ++u) {
      const int i = lane + 32 * u;
      if (i < EG_TOK * E) sx[warp][i / E][i % E] = ns[u];
    }
    float dv[EG_TOK];
    if (BWD) {
#pragma unroll
      for (int q = 0; q < EG_TOK; ++q) dv[q] = ndv[q];
    }
    fetch(g + gstride);
    __syncwarp();
#pragma unroll
    for (int q = 0; q < EG_TOK; ++q) {
      if (q < ntok) {
        const int64_t tok = tok0 + q;
        float r = bias;
#pragma unroll
        for (int k = 0; k < EP; k += 4) {
          const float4 xv = *reinterpret_cast<const float4 *>(&sx[warp][q][k]);
          r = fmaf(xv.x, w[k], r); r = fmaf(xv.y, w[k + 1], r); r = fmaf(xv.z, w[k + 2], r); r = fmaf(xv.w, w[k + 3], r);
        }
        const bool keep = drop.thr == 0 || drop_keep(drop.key, drop.thr, (uint64_t)(e0 + tok * 32 + lane));
        if (!BWD) {
          const float v = fmaxf(r, 0.f) + __ldg(pe + (tok & 31) * 32 + lane);
          x0[tok * 32 + lane] = keep ? v * drop.scale : 0.f;
        } else {
          const float gq = (keep && r > 0.f) ? dv[q] * drop.scale : 0.f;
          accb += gq;
#pragma unroll
          for (int k = 0; k < EP; k += 4) {
            const float4 xv = *reinterpret_cast<const float4 *>(&sx[warp][q][k]);
            acc[k] = fmaf(gq, xv.x, acc[k]); acc[k + 1] = fmaf(gq, xv.y, acc[k + 1]);
            acc[k + 2] = fmaf(gq, xv.z, acc[k + 2]); acc[k + 3] = fmaf(gq, xv.w, acc[k + 3]);
          }
        }
      }
    }
    __syncwarp();
  }
  if (BWD) {
#pragma unroll
    for (int k = 0; k < E; ++k) atomicAdd(&sacc[lane * E + k], acc[k]);
    atomicAdd(&sacc[32 * E + lane], accb);
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * E; i += blockDim.x) atomicAdd(gW + i, sacc[i]);
    if (threadIdx.x < 32) atomicAdd(gb + threadIdx.x, sacc[32 * E + threadIdx.x]);
  }
}

int edge32_stem_fwd(const float *src, int E, const float *W, const float *b, const float *pe, float *x0, int64_t M, const Drop &drop,
                    int64_t row0, cudaStream_t st) {
  if (M == 0) return 0;
  GT_CHECK(E == 16 || E == 27, "edge32: embedding_size_src must be 16 or 27");
  LaunchScope _ls(KC_TC_INPUT, st);
  if (E == 16) GT_CUDA(launch_pdl(stem32_kernel<16, false>, dim3(eg_blocks(M)), dim3(EG_WARPS * 32), 0, st, src, W, b, pe, x0, (const float *)nullptr, (float *)nullptr, (float *)nullptr, M, drop, row0 * 32));
  else GT_CUDA(launch_pdl(stem32_kernel<27, false>, dim3(eg_blocks(M)), dim3(EG_WARPS * 32), 0, st, src, W, b, pe, x0, (const float *)nullptr, (float *)nullptr, (float *)nullptr, M, drop, row0 * 32));
  GT_CUDA(cudaGetLastError());
  return 0;
}
int edge32_stem_bwd(const float *dx0, const float *src, int E, const float *W, const float *b, float *gW, float *gb, int64_t M,
                    const Drop &drop, int64_t row0, cudaStream_t st) {
  if (M == 0) return 0;
  GT_CHECK(E == 16 || E == 27, "edge32: embedding_size_src must be 16 or 27");
  LaunchScope _ls(KC_TC_INPUT, st);
  if (E == 16) GT_CUDA(launch_pdl(stem32_kernel<16, true>, dim3(eg_blocks(M)), dim3(EG_WARPS * 32), 0, st, src, W, b, (const float *)nullptr, (float *)nullptr, dx0, gW, gb, M, drop, row0 * 32));
  else GT_CUDA(launch_pdl(stem32_kernel<27, true>, dim3(eg_blocks(M)), dim3(EG_WARPS * 32), 0, st, src, W, b, (const float *)nullptr, (float *)nullptr, dx0, gW, gb, M, drop, row0 * 32));
  GT_CUDA(cudaGetLastError());
  return 0;
}

// ---- tail forward (+ optional fused loss) ---------------------------------------------------------------------------
__global__ void __launch_bounds__(EG_WARPS * 32) tail32_fwd_kernel(const float *__restrict__ x, const float *__restrict__ gamma,
                                                                   const float *__restrict__ beta, const float *__restrict__ Wout,
                                                                   const float *__restrict__ bout, float *__restrict__ hvo,
                                                                   float *mean, float *rstd, int64_t M, float thres,
                                                                   const float *__restrict__ y, float penalty, float gscale,
                                                                   float *__restrict__ dlog, float *partials) {
  __shared__ __align__(16) float sz[EG_WARPS][EG_TOK][32];
  __shared__ float red[4][EG_WARPS];
  griddep_wait();
  griddep_launch();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int role = lane / 9;                       // 0 hits, 1 velocities, 2 offsets, 3 idle lanes (27..31)
  float wr[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) wr[c] = lane < 27 ? Wout[lane * 32 + c] : 0.f;
  const float bj = lane < 27 ? bout[lane] : 0.f, gm = gamma[lane], be = beta[lane];
  float a_loss = 0.f, a_ok = 0.f;                  // role 0: bce / correct hits ; role 1: velocity mse ; role 2: offset mse
  const int64_t n_groups = (M + EG_TOK - 1) / EG_TOK;
  const int64_t gstride = (int64_t)gridDim.x * EG_WARPS;
  // software pipeline (as in tail32_bwd_kernel): the rows of the warp's NEXT group are requested before the current group is
  // processed — at ~16 resident warps per SM (32 weight registers per lane) the load latency was exposed once per group
  float nx[EG_TOK], nyj[EG_TOK], nyh[EG_TOK];
  auto fetch = [&](int64_t gn) {
    const int64_t t0 = gn * EG_TOK;
#pragma unroll
    for (int q = 0; q < EG_TOK; ++q) {
      const bool ok = gn < n_groups && t0 + q < M;
      nx[q] = ok ? __ldg(x + (t0 + q) * 32 + lane) : 0.f;
      nyj[q] = 0.f; nyh[q] = 0.f;
      if (y != nullptr && ok && lane < 27) {
        nyj[q] = __ldg(y + (t0 + q) * 27 + lane);
        nyh[q] = __ldg(y + (t0 + q) * 27 + (lane - 9 * role));
      }
    }
  };
  fetch((int64_t)blockIdx.x * EG_WARPS + warp);
  for (int64_t g = (int64_t)blockIdx.x * EG_WARPS + warp; g < n_groups; g += gstride) {
    const int64_t tok0 = g * EG_TOK;
    const int ntok = (int)min((int64_t)EG_TOK, M - tok0);
    float xv[EG_TOK], yj[EG_TOK], yh[EG_TOK];
#pragma unroll
    for (int q = 0; q < EG_TOK; ++q) { xv[q] = nx[q]; yj[q] = nyj[q]; yh[q] = nyh[q]; }
    fetch(g + gstride);
#pragma unroll
    for (int q = 0; q < EG_TOK; ++q) {
      const float mu = eg_warp_sum(xv[q]) * (1.f / 32);
      const float t = xv[q] - mu;
      const float rs = rsqrtf(eg_warp_sum(t * t) * (1.f / 32) + LN_EPS);
      sz[warp][q][lane] = t * rs * gm + be;
      if (lane == 0 && mean != nullptr && q < ntok) { mean[tok0 + q] = mu; rstd[tok0 + q] = rs; }
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < EG_TOK; ++q) {
      float a0 = bj, a1 = 0.f;
#pragma unroll
      for (int c = 0; c < 32; c += 8) {
        const float4 z0 = *reinterpret_cast<const float4 *>(&sz[warp][q][c]), z1 = *reinterpret_cast<const float4 *>(&sz[warp][q][c + 4]);
        a0 = fmaf(z0.x, wr[c], a0); a0 = fmaf(z0.y, wr[c + 1], a0); a0 = fmaf(z0.z, wr[c + 2], a0); a0 = fmaf(z0.w, wr[c + 3], a0);
        a1 = fmaf(z1.x, wr[c + 4], a1); a1 = fmaf(z1.y, wr[c + 5], a1); a1 = fmaf(z1.z, wr[c + 6], a1); a1 = fmaf(z1.w, wr[c + 7], a1);
      }
      const float lg = a0 + a1;
      if (q < ntok && lane < 27) {
        const int64_t o = (tok0 + q) * 27 + lane;
        // one sigmoid serves the three roles (no divergent expf / tanhf / log1pf paths): v = sigmoid(x), o = 0.5 tanh(x) =
        // sigmoid(2 x) - 0.5, and the hit BCE softplus(h) - h y = h (1 - y) + log(1 + exp(-h)) reuses exp(-h)
        const float e = __expf(role == 2 ? -2.f * lg : -lg);
        const float sg = __fdividef(1.f, 1.f + e);
        float out;
        if (role == 0) out = thres >= 0.f ? (sg > thres ? 1.f : 0.f) : lg;
        else if (role == 1) out = sg;
        else out = sg - 0.5f;
        hvo[o] = out;
        if (y != nullptr) {
          const float w = (yh[q] == 1.f) ? 1.f : penalty;
          float dl;
          if (role == 0) {
            a_loss = fmaf(fmaf(lg, 1.f - yh[q], __logf(1.f + e)), w, a_loss);
            a_ok += ((sg > 0.5f ? 1.f : 0.f) == yh[q]) ? 1.f : 0.f;
            dl = gscale * w * (sg - yh[q]);
          } else {
            const float df = out - yj[q];
            a_loss = fmaf(df * df, w, a_loss);
            dl = gscale * 2.f * w * df * (role == 1 ? out * (1.f - out) : (0.5f - 2.f * out * out));
          }
          dlog[o] = dl;
        }
      }
    }
    __syncwarp();
  }
  if (partials != nullptr) {
    const float s0 = eg_warp_sum(role == 0 ? a_loss : 0.f), s1 = eg_warp_sum(role == 1 ? a_loss : 0.f);
    const float s2 = eg_warp_sum(role == 2 ? a_loss : 0.f), s3 = eg_warp_sum(a_ok);
    if (lane == 0) { red[0][warp] = s0; red[1][warp] = s1; red[2][warp] = s2; red[3][warp] = s3; }
    __syncthreads();
    if (threadIdx.x < 4) {
      float s = 0.f;
      for (int i = 0; i < EG_WARPS; ++i) s += red[threadIdx.x][i];
      partials[(int64_t)blockIdx.x * 4 + threadIdx.x] = s;
    }
  }
}

int loss_finalize(const float *partials, int64_t blocks, int64_t M, float *metrics6, cudaStream_t st);   // kernels_simt.cu

int edge32_tail_fwd(const float *x, const float *gamma, const float *beta, const float *Wout, const float *bout, float *hvo,
                    float *mean, float *rstd, int64_t M, float thres, cudaStream_t st) {
  if (M == 0) return 0;
  { LaunchScope _ls(KC_TC_HEAD, st);
    GT_CUDA(launch_pdl(tail32_fwd_kernel, dim3(eg_blocks(M)), dim3(EG_WARPS * 32), 0, st, x, gamma, beta, Wout, bout, hvo, mean, rstd, M, thres,
                       (const float *)nullptr, 0.f, 0.f, (float *)nullptr, (float *)nullptr)); }
  GT_CUDA(cudaGetLastError());
  return 0;
}
// forward tail + calculate_loss: hvo, the six metrics and dL/dlogits (scaled by 1 / M) in one pass
int edge32_tail_fwd_loss(const float *x, const float *gamma, const float *beta, const float *Wout, const float *bout, float *hvo,
                         float *mean, float *rstd, int64_t M, const float *y, float penalty, float *dlog, float *partials,
                         float *metrics6, cudaStream_t st) {
  GT_CHECK(M > 0, "empty batch");
  const int blocks = eg_blocks(M);
  { LaunchScope _ls(KC_TC_HEAD, st);
    GT_CUDA(launch_pdl(tail32_fwd_kernel, dim3(blocks), dim3(EG_WARPS * 32), 0, st, x, gamma, beta, Wout, bout, hvo, mean, rstd, M, -1.f, y,
                       penalty, 1.f / (float)M, dlog, partials)); }
  GT_CUDA(cudaGetLastError());
  return loss_finalize(partials, blocks, M, metrics6, st);
}

// ---- tail backward --------------------------------------------------------------------------------------------------
// d_in: dL/dlogits when hvo == nullptr, else dL/d(h, v, o) (the activation derivative is applied here from hvo)
__global__ void __launch_bounds__(EG_WARPS * 32, 2) tail32_bwd_kernel(const float *__restrict__ d_in, const float *__restrict__ hvo,
                                                                   const float *__restrict__ x, const float *__restrict__ mean,
                                                                   const float *__restrict__ rstd, const float *__restrict__ gamma,
                                                                   const float *__restrict__ beta, const float *__restrict__ Wout,
                                                                   float *__restrict__ dx, float *gW, float *gb, float *gg, float *gbe,
                                                                   int64_t M) {
  __shared__ __align__(16) float sz[EG_WARPS][EG_TOK][32], sd[EG_WARPS][EG_TOK][32];
  __shared__ float sacc[27 * 32 + 32 + 64];
  griddep_wait();
  griddep_launch();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int role = lane / 9;
  // lane c keeps column c of W_out [27][32] (+ a zero row) in registers: read from shared memory it was 28 of the 45 shared-memory
  // wavefronts per token, and the kernel ran at the shared-memory pipe's rate (2.6 x its HBM time)
  float wc[28];
#pragma unroll
  for (int j = 0; j < 28; ++j) wc[j] = j < 27 ? __ldg(Wout + j * 32 + lane) : 0.f;
  const float gm = gamma[lane], be = beta[lane];
  float accw[32], accb = 0.f, adg = 0.f, adb = 0.f;
#pragma unroll
  for (int c = 0; c < 32; ++c) accw[c] = 0.f;
  for (int i = threadIdx.x; i < 27 * 32 + 96; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int64_t n_groups = (M + EG_TOK - 1) / EG_TOK;
  const int64_t gstride = (int64_t)gridDim.x * EG_WARPS;
  // software pipeline: the rows of the warp's NEXT group are loaded before the current group is processed, so the global
  // load latency (this kernel runs at 16 warps per SM) overlaps a whole group of arithmetic
  // (row statistics: lane q holds those of token q — one load per lane instead of EG_TOK broadcast loads, handed out by shuffles)
  float nx[EG_TOK], nd[EG_TOK], nmu = 0.f, nrs = 0.f;
  auto fetch = [&](int64_t g) {
    const int64_t tok0 = g * EG_TOK;
    const bool okl = g < n_groups && lane < EG_TOK && tok0 + lane < M;
    nmu = okl ? __ldg(mean + tok0 + lane) : 0.f;
    nrs = okl ? __ldg(rstd + tok0 + lane) : 0.f;
#pragma unroll
    for (int q = 0; q < EG_TOK; ++q) {
      const bool ok = g < n_groups && tok0 + q < M;
      nx[q] = ok ? __ldg(x + (tok0 + q) * 32 + lane) : 0.f;
      float d = 0.f;
      if (ok && lane < 27) {
        d = __ldg(d_in + (tok0 + q) * 27 + lane);
        if (hvo != nullptr && role > 0) {
          const float a = __ldg(hvo + (tok0 + q) * 27 + lane);
          d *= role == 1 ? a * (1.f - a) : (0.5f - 2.f * a * a);
        }
      }
      nd[q] = d;
    }
  };
  int64_t g = (int64_t)blockIdx.x * EG_WARPS + warp;
  fetch(g);
  for (; g < n_groups; g += gstride) {
    const int64_t tok0 = g * EG_TOK;
    const int ntok = (int)min((int64_t)EG_TOK, M - tok0);
    float xh[EG_TOK], rs[EG_TOK], dq[EG_TOK];
#pragma unroll
    for (int q = 0; q < EG_TOK; ++q) {
      rs[q] = __shfl_sync(0xffffffffu, nrs, q);
      xh[q] = (nx[q] - __shfl_sync(0xffffffffu, nmu, q)) * rs[q];
      dq[q] = nd[q];
      sz[warp][q][lane] = q < ntok ? xh[q] * gm + be : 0.f;
      sd[warp][q][lane] = dq[q];
      accb += dq[q];
    }
    fetch(g + gstride);
    __syncwarp();
#pragma unroll
    for (int q = 0; q < EG_TOK; ++q) {
      // lane j: dW_out[j][:] += dlogit_j z ;  lane c: dz_c = sum_j dlogit_j W_out[j][c]
      float dz0 = 0.f, dz1 = 0.f;
#pragma unroll
      for (int c = 0; c < 32; c += 4) {
        const float4 zz = *reinterpret_cast<const float4 *>(&sz[warp][q][c]);
        accw[c] = fmaf(dq[q], zz.x, accw[c]); accw[c + 1] = fmaf(dq[q], zz.y, accw[c + 1]);
        accw[c + 2] = fmaf(dq[q], zz.z, accw[c + 2]); accw[c + 3] = fmaf(dq[q], zz.w, accw[c + 3]);
      }
#pragma unroll
      for (int j = 0; j < 28; j += 4) {
        const float4 dd = *reinterpret_cast<const float4 *>(&sd[warp][q][j]);
        dz0 = fmaf(dd.x, wc[j], dz0); dz1 = fmaf(dd.y, wc[j + 1], dz1);
        dz0 = fmaf(dd.z, wc[j + 2], dz0); dz1 = fmaf(dd.w, wc[j + 3], dz1);
      }
      const float dz = dz0 + dz1, gd = dz * gm;
      const float s1 = eg_warp_sum(gd) * (1.f / 32), s2 = eg_warp_sum(gd * xh[q]) * (1.f / 32);
      if (q < ntok) dx[(tok0 + q) * 32 + lane] = (gd - s1 - xh[q] * s2) * rs[q];
      adg = fmaf(dz, xh[q], adg);
      adb += dz;
    }
    __syncwarp();
  }
  if (lane < 27) {
#pragma unroll
    for (int c = 0; c < 32; ++c) atomicAdd(&sacc[lane * 32 + c], accw[c]);
    atomicAdd(&sacc[27 * 32 + lane], accb);
  }
  atomicAdd(&sacc[27 * 32 + 32 + lane], adg);
  atomicAdd(&sacc[27 * 32 + 64 + lane], adb);
  __syncthreads();
  for (int i = threadIdx.x; i < 27 * 32; i += blockDim.x) atomicAdd(gW + i, sacc[i]);
  if (threadIdx.x < 27) atomicAdd(gb + threadIdx.x, sacc[27 * 32 + threadIdx.x]);
  if (threadIdx.x < 32) {
    atomicAdd(gg + threadIdx.x, sacc[27 * 32 + 32 + threadIdx.x]);
    atomicAdd(gbe + threadIdx.x, sacc[27 * 32 + 64 + threadIdx.x]);
  }
}

int edge32_tail_bwd(const float *d_in, const float *hvo, const float *x, const float *mean, const float *rstd, const float *gamma,
                    const float *beta, const float *Wout, float *dx, float *gW, float *gb, float *gg, float *gbe, int64_t M,
                    cudaStream_t st) {
  if (M == 0) return 0;
  { LaunchScope _ls(KC_TC_HEAD, st);
    GT_CUDA(launch_pdl(tail32_bwd_kernel, dim3(eg_blocks(M)), dim3(EG_WARPS * 32), 0, st, d_in, hvo, x, mean, rstd, gamma, beta, Wout, dx, gW, gb,
                       gg, gbe, M)); }
  GT_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace gt
