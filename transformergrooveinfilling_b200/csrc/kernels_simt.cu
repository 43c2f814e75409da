// kernels_simt.cu — fp32 CUDA-core kernels of the groove hot path (precision mode GT_PREC_FP32).
//
// These are the exact-arithmetic path (per-step loss within 1e-4 of the reference) and the generic
// building blocks for every shape the reference's sweeps can produce (any d_model / nhead /
// dim_feedforward).  The bf16 tensor-core path (tc_*.cu) replaces the hot stages for the shapes
// named in BASELINE.json.
#include "common.cuh"

namespace gt {

// =============================================================================================
// GEMM: C[m,n] = epi( sum_k A(m,k) B(n,k) ), generic strides, optional split-K with atomics
// =============================================================================================
constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;

struct GemmArgs {
  const float *A, *B;
  float *C;
  int64_t sam, sak, sbn, sbk, ldc, M, N, K, kchunk;
  GemmEpi e;
};

__global__ void __launch_bounds__(256) gemm_f32_kernel(GemmArgs g) {
  drop_resolve(g.e.drop);                            // graph replay: key from the device step counter
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int64_t n0 = (int64_t)blockIdx.y * BN;
  const int64_t k_begin = (int64_t)blockIdx.z * g.kchunk;
  const int64_t k_end = min(g.K, k_begin + g.kchunk);
  const bool a_kcontig = (g.sak == 1), b_kcontig = (g.sbk == 1);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int64_t k0 = k_begin; k0 < k_end; k0 += BK) {
#pragma unroll
    for (int i = 0; i < (BM * BK) / 256; ++i) {
      int e = tid + i * 256;
      int mm, kk;
      if (a_kcontig) { mm = e / BK; kk = e % BK; } else { mm = e % BM; kk = e / BM; }
      int64_t m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < g.M && k < k_end) ? g.A[m * g.sam + k * g.sak] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < (BN * BK) / 256; ++i) {
      int e = tid + i * 256;
      int nn, kk;
      if (b_kcontig) { nn = e / BK; kk = e % BK; } else { nn = e % BN; kk = e / BN; }
      int64_t n = n0 + nn, k = k0 + kk;
      Bs[kk][nn] = (n < g.N && k < k_end) ? g.B[n * g.sbn + k * g.sbk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a4 = *reinterpret_cast<const float4 *>(&As[kk][ty * TM]);
      float4 b4 = *reinterpret_cast<const float4 *>(&Bs[kk][tx * TN]);
      float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  const GemmEpi &e = g.e;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int64_t m = m0 + ty * TM + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int64_t n = n0 + tx * TN + j;
      if (n >= g.N) continue;
      float v = acc[i][j];
      if (e.bias && blockIdx.z == 0) v += e.bias[n];
      if (e.relu) v = fmaxf(v, 0.f);
      if (e.pe) v += e.pe[(m % T) * g.N + n];
      if (e.drop.thr) v = drop_keep(e.drop.key, e.drop.thr, (uint64_t)((e.drop_row0 + m) * g.N + n)) ? v * e.drop.scale : 0.f;
      if (e.mask_pos) v = (e.mask_pos[m * e.ld_mask + n] > 0.f) ? v * e.mask_scale : 0.f;
      if (e.residual) v += e.residual[m * e.ld_res + n];
      float *c = g.C + m * g.ldc + n;
      if (e.atomic) atomicAdd(c, v);
      else if (e.accumulate) *c += v;
      else *c = v;
    }
  }
}

int gemm_f32(const float *A, int64_t sam, int64_t sak, const float *B, int64_t sbn, int64_t sbk, float *C,
             int64_t ldc, int64_t M, int64_t N, int64_t K, const GemmEpi &epi, int64_t split_k_chunk,
             cudaStream_t st) {
  if (M == 0 || N == 0) return 0;
  GemmArgs g;
  g.A = A; g.B = B; g.C = C; g.sam = sam; g.sak = sak; g.sbn = sbn; g.sbk = sbk; g.ldc = ldc;
  g.M = M; g.N = N; g.K = K; g.e = epi;
  int64_t splits = 1;
  if (split_k_chunk > 0 && K > split_k_chunk) {
    split_k_chunk = (split_k_chunk + BK - 1) / BK * BK;
    splits = (K + split_k_chunk - 1) / split_k_chunk;
    while (splits > 32768) { split_k_chunk *= 2; splits = (K + split_k_chunk - 1) / split_k_chunk; }
    g.kchunk = split_k_chunk;
    GT_CHECK(epi.atomic, "split-K GEMM needs an atomic epilogue");
  } else {
    g.kchunk = (K + BK - 1) / BK * BK;
    if (g.kchunk == 0) g.kchunk = BK;
  }
  dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((N + BN - 1) / BN), (unsigned)splits);
  GT_CHECK(grid.y <= 65535, "GEMM N too large");
  { LaunchScope _ls(KC_GEMM_F32, st);
  gemm_f32_kernel<<<grid, 256, 0, st>>>(g); }
  GT_CUDA(cudaGetLastError());
  return 0;
}

// out[n] += sum_m X[m*ld+n]
__global__ void colsum_kernel(const float *__restrict__ X, int64_t ld, int64_t M, int N, float *out, int rows_per_block) {
  int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  int64_t r1 = min(M, r0 + rows_per_block);
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    float s = 0.f;
    for (int64_t r = r0; r < r1; ++r) s += X[r * ld + n];
    atomicAdd(out + n, s);
  }
}
int colsum_f32(const float *X, int64_t ld, int64_t M, int N, float *out, cudaStream_t st) {
  if (M == 0) return 0;
  int rows = 256;
  int64_t blocks = (M + rows - 1) / rows;
  int threads = N >= 256 ? 256 : ((N + 31) / 32 * 32);
  { LaunchScope _ls(KC_ELEMWISE, st);
  colsum_kernel<<<(unsigned)blocks, threads, 0, st>>>(X, ld, M, N, out, rows); }
  GT_CUDA(cudaGetLastError());
  return 0;
}

// =============================================================================================
// Attention: one warp per (sequence, head); lane = query row (forward) / key row (dK,dV)
// =============================================================================================
__global__ void attention_fwd_kernel(AttnArgs a, int warps) {
  drop_resolve(a.drop);
  extern __shared__ float sm[];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int64_t pair = (int64_t)blockIdx.x * warps + warp;
  if (pair >= a.n_seq * a.H) return;
  const int64_t seq = pair / a.H;
  const int head = (int)(pair % a.H);
  const int dh = a.dh, ls = dh + 1;
  float *Qs = sm + (size_t)warp * 3 * T * ls, *Ks = Qs + T * ls, *Vs = Ks + T * ls;
  for (int e = lane; e < T * dh; e += 32) {
    int r = e / dh, c = e % dh;
    int64_t row = seq * T + r;
    Qs[r * ls + c] = a.q[row * a.ldq + head * dh + c];
    Ks[r * ls + c] = a.k[row * a.ldk + head * dh + c];
    Vs[r * ls + c] = a.v[row * a.ldv + head * dh + c];
  }
  __syncwarp();
  float s[T];
#pragma unroll
  for (int j = 0; j < T; ++j) s[j] = 0.f;
  for (int c = 0; c < dh; ++c) {
    float qc = Qs[lane * ls + c];
#pragma unroll
    for (int j = 0; j < T; ++j) s[j] = fmaf(qc, Ks[j * ls + c], s[j]);
  }
  const float scale = rsqrtf((float)dh);
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < T; ++j) {
    s[j] *= scale;
    if (a.causal && j > lane) s[j] = -INFINITY;
    mx = fmaxf(mx, s[j]);
  }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < T; ++j) { s[j] = expf(s[j] - mx); sum += s[j]; }
  const float inv = 1.f / sum;
  const uint64_t base = (uint64_t)((((a.seq0 + seq) * a.H + head) * T + lane) * T);
#pragma unroll
  for (int j = 0; j < T; ++j) {
    float p = s[j] * inv;
    if (a.drop.thr) p = drop_keep(a.drop.key, a.drop.thr, base + key_perm(j)) ? p * a.drop.scale : 0.f;
    s[j] = p;
  }
  __syncwarp();
  for (int c = 0; c < dh; ++c) {
    float o = 0.f;
#pragma unroll
    for (int j = 0; j < T; ++j) o = fmaf(s[j], Vs[j * ls + c], o);
    Qs[lane * ls + c] = o;     // own row only: no cross-lane hazard
  }
  __syncwarp();
  for (int e = lane; e < T * dh; e += 32) {
    int r = e / dh, c = e % dh;
    a.o[(seq * T + r) * a.ldo + head * dh + c] = Qs[r * ls + c];
  }
}

// Backward.  The head's feature columns are processed in chunks of ATTN_BWD_CH (shared memory per warp is bounded by the
// chunk, not by head_dim: d_model = 512 with one head — inside the reference's sweep ranges, configs/*_sweep.yaml — needs
// 271 KB unchunked).  Pass 1 accumulates the scores and dP over the chunks in column order (the same order as the unchunked
// loop, so results are bit-identical for head dims that fit one chunk); pass 2 re-stages each chunk of q | k | dO and emits
// that chunk's columns of dq | dk | dv.
constexpr int ATTN_BWD_CH = 128;
__global__ void attention_bwd_kernel(AttnArgs a, int warps) {
  drop_resolve(a.drop);
  extern __shared__ float sm[];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int64_t pair = (int64_t)blockIdx.x * warps + warp;
  if (pair >= a.n_seq * a.H) return;
  const int64_t seq = pair / a.H;
  const int head = (int)(pair % a.H);
  const int dh = a.dh, ch = dh < ATTN_BWD_CH ? dh : ATTN_BWD_CH, ls = ch + 1;
  const size_t per_warp = (size_t)4 * T * ls + 2 * T * (T + 1);
  float *Qs = sm + warp * per_warp, *Ks = Qs + T * ls, *Vs = Ks + T * ls, *Gs = Vs + T * ls;
  float *dSs = Gs + T * ls, *Ps = dSs + T * (T + 1);
  auto stage = [&](int c0, int cw, bool with_v) {
    __syncwarp();
    for (int e = lane; e < T * cw; e += 32) {
      int r = e / cw, c = e % cw;
      int64_t row = seq * T + r;
      Qs[r * ls + c] = a.q[row * a.ldq + head * dh + c0 + c];
      Ks[r * ls + c] = a.k[row * a.ldk + head * dh + c0 + c];
      if (with_v) Vs[r * ls + c] = a.v[row * a.ldv + head * dh + c0 + c];
      Gs[r * ls + c] = a.d_o[row * a.ld_do + head * dh + c0 + c];
    }
    __syncwarp();
  };
  float s[T], dp[T];
#pragma unroll
  for (int j = 0; j < T; ++j) { s[j] = 0.f; dp[j] = 0.f; }
  for (int c0 = 0; c0 < dh; c0 += ch) {
    const int cw = dh - c0 < ch ? dh - c0 : ch;
    stage(c0, cw, true);
    for (int c = 0; c < cw; ++c) {
      float qc = Qs[lane * ls + c], gc = Gs[lane * ls + c];
#pragma unroll
      for (int j = 0; j < T; ++j) {
        s[j] = fmaf(qc, Ks[j * ls + c], s[j]);
        dp[j] = fmaf(gc, Vs[j * ls + c], dp[j]);
      }
    }
  }
  const float scale = rsqrtf((float)dh);
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < T; ++j) {
    s[j] *= scale;
    if (a.causal && j > lane) s[j] = -INFINITY;
    mx = fmaxf(mx, s[j]);
  }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < T; ++j) { s[j] = expf(s[j] - mx); sum += s[j]; }
  const float inv = 1.f / sum;
  const uint64_t base = (uint64_t)((((a.seq0 + seq) * a.H + head) * T + lane) * T);
  float delta = 0.f;
#pragma unroll
  for (int j = 0; j < T; ++j) {
    float p = s[j] * inv;
    float keep = 1.f;
    if (a.drop.thr) keep = drop_keep(a.drop.key, a.drop.thr, base + key_perm(j)) ? a.drop.scale : 0.f;
    Ps[lane * (T + 1) + j] = p * keep;   // dropped probabilities (feed dV)
    dp[j] *= keep;                       // gradient w.r.t. the un-dropped probability
    delta = fmaf(dp[j], p, delta);
    s[j] = p;
  }
#pragma unroll
  for (int j = 0; j < T; ++j) {
    float ds = s[j] * (dp[j] - delta) * scale;
    dSs[lane * (T + 1) + j] = ds;
    dp[j] = ds;
  }
  for (int c0 = 0; c0 < dh; c0 += ch) {
    const int cw = dh - c0 < ch ? dh - c0 : ch;
    if (dh > ch) stage(c0, cw, false);     // single chunk: q | k | dO are still staged
    // dQ[i,c] = sum_j dS[i,j] K[j,c]
    for (int c = 0; c < cw; ++c) {
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < T; ++j) acc = fmaf(dp[j], Ks[j * ls + c], acc);
      a.dq[(seq * T + lane) * a.ld_dq + head * dh + c0 + c] = acc;
    }
    __syncwarp();
    // lane = key row j: dK[j,c] = sum_i dS[i,j] Q[i,c] ; dV[j,c] = sum_i Pd[i,j] dO[i,c]
    for (int c = 0; c < cw; ++c) {
      float dk = 0.f, dv = 0.f;
#pragma unroll
      for (int i = 0; i < T; ++i) {
        dk = fmaf(dSs[i * (T + 1) + lane], Qs[i * ls + c], dk);
        dv = fmaf(Ps[i * (T + 1) + lane], Gs[i * ls + c], dv);
      }
      a.dk[(seq * T + lane) * a.ld_dk + head * dh + c0 + c] = dk;
      a.dv[(seq * T + lane) * a.ld_dv + head * dh + c0 + c] = dv;
    }
  }
}

static int attn_launch(const AttnArgs &a, bool bwd, cudaStream_t st) {
  if (a.n_seq == 0) return 0;
  const int ls = a.dh + 1, ls_b = (a.dh < ATTN_BWD_CH ? a.dh : ATTN_BWD_CH) + 1;
  size_t per_warp = bwd ? ((size_t)4 * T * ls_b + 2 * T * (T + 1)) * sizeof(float) : (size_t)3 * T * ls * sizeof(float);
  int warps = 8;
  while (warps > 1 && per_warp * warps > 96 * 1024) warps >>= 1;
  size_t smem = per_warp * warps;
  GT_CHECK(smem <= 200 * 1024, "head dim too large for the SIMT attention kernel");
  int64_t pairs = a.n_seq * a.H;
  int64_t blocks = (pairs + warps - 1) / warps;
  if (bwd) {
    GT_CUDA(cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    { LaunchScope _ls(KC_ATTN_BWD, st);
    attention_bwd_kernel<<<(unsigned)blocks, warps * 32, smem, st>>>(a, warps); }
  } else {
    GT_CUDA(cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    { LaunchScope _ls(KC_ATTN_FWD, st);
    attention_fwd_kernel<<<(unsigned)blocks, warps * 32, smem, st>>>(a, warps); }
  }
  GT_CUDA(cudaGetLastError());
  return 0;
}
int attention_fwd(const AttnArgs &a, cudaStream_t st) { return attn_launch(a, false, st); }
int attention_bwd(const AttnArgs &a, cudaStream_t st) { return attn_launch(a, true, st); }

// =============================================================================================
// LayerNorm (+ residual + dropout) forward / backward — one warp per token row
// =============================================================================================
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void ln_fwd_kernel(const float *__restrict__ a, const float *__restrict__ res, const float *__restrict__ gamma,
                              const float *__restrict__ beta, float *u_out, float *y, float *mean, float *rstd,
                              int64_t M, int d, Drop drop, int64_t row0) {
  drop_resolve(drop);
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x / 32) + warp;
  if (row >= M) return;
  constexpr int MAXV = 16;             // d <= 512
  float u[MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int c = lane + i * 32;
    float v = 0.f;
    if (c < d) {
      v = a[row * d + c];
      if (drop.thr) v = drop_keep(drop.key, drop.thr, (uint64_t)((row0 + row) * d + c)) ? v * drop.scale : 0.f;
      if (res) v += res[row * d + c];
      s += v;
    }
    u[i] = v;
  }
  const float mu = warp_sum(s) / d;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int c = lane + i * 32;
    if (c < d) { float t = u[i] - mu; q = fmaf(t, t, q); }
  }
  const float rs = rsqrtf(warp_sum(q) / d + LN_EPS);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int c = lane + i * 32;
    if (c < d) {
      if (u_out) u_out[row * d + c] = u[i];
      y[row * d + c] = (u[i] - mu) * rs * gamma[c] + beta[c];
    }
  }
  if (lane == 0 && mean) { mean[row] = mu; rstd[row] = rs; }
}

// ---- d_model a multiple of 128: float4 per lane (NV vectors of 4 consecutive columns), one dropout hash per vector ----------
__device__ __forceinline__ float4 ln_drop4(const Drop &drop, float4 v, uint64_t e) {   // e: element index of v.x (multiple of 4)
  const uint64_t w = e >> 2;
  uint32_t lo, hi;
  hash_quad((uint32_t)w ^ ((uint32_t)(w >> 32) * 0x85EBCA6Bu), drop.key, lo, hi);
  v.x = ((lo & 0xFFFFu) >= drop.thr) ? v.x * drop.scale : 0.f;
  v.y = ((lo >> 16) >= drop.thr) ? v.y * drop.scale : 0.f;
  v.z = ((hi & 0xFFFFu) >= drop.thr) ? v.z * drop.scale : 0.f;
  v.w = ((hi >> 16) >= drop.thr) ? v.w * drop.scale : 0.f;
  return v;
}

template <int NV>
__global__ void __launch_bounds__(256) ln_fwd_vec4_kernel(const float *__restrict__ a, const float *__restrict__ res,
                                                          const float *__restrict__ gamma, const float *__restrict__ beta, float *u_out,
                                                          float *y, float *mean, float *rstd, int64_t M, Drop drop, int64_t row0) {
  drop_resolve(drop);
  constexpr int d = NV * 128;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x / 32) + warp;
  if (row >= M) return;
  float4 u[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (lane + 32 * i) * 4;
    float4 v = __ldg(reinterpret_cast<const float4 *>(a + row * d + c));
    if (drop.thr) v = ln_drop4(drop, v, (uint64_t)((row0 + row) * d + c));
    if (res) {
      const float4 r = __ldg(reinterpret_cast<const float4 *>(res + row * d + c));
      v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    u[i] = v;
    s += (v.x + v.y) + (v.z + v.w);
  }
  const float mu = warp_sum(s) / d;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float t0 = u[i].x - mu, t1 = u[i].y - mu, t2 = u[i].z - mu, t3 = u[i].w - mu;
    q = fmaf(t0, t0, q); q = fmaf(t1, t1, q); q = fmaf(t2, t2, q); q = fmaf(t3, t3, q);
  }
  const float rs = rsqrtf(warp_sum(q) / d + LN_EPS);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (lane + 32 * i) * 4;
    if (u_out) *reinterpret_cast<float4 *>(u_out + row * d + c) = u[i];
    const float4 g = __ldg(reinterpret_cast<const float4 *>(gamma + c)), b = __ldg(reinterpret_cast<const float4 *>(beta + c));
    *reinterpret_cast<float4 *>(y + row * d + c) =
        make_float4((u[i].x - mu) * rs * g.x + b.x, (u[i].y - mu) * rs * g.y + b.y, (u[i].z - mu) * rs * g.z + b.z, (u[i].w - mu) * rs * g.w + b.w);
  }
  if (lane == 0 && mean) { mean[row] = mu; rstd[row] = rs; }
}

template <int NV>
__global__ void __launch_bounds__(256) ln_bwd_vec4_kernel(const float *__restrict__ dy, const float *__restrict__ u, const float *__restrict__ mean,
                                                          const float *__restrict__ rstd, const float *__restrict__ gamma, float *du, float *da,
                                                          float *dgamma, float *dbeta, int64_t M, Drop drop, int64_t row0, int rows_per_warp) {
  drop_resolve(drop);
  constexpr int d = NV * 128;
  __shared__ float sm[2 * d];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, wpb = blockDim.x / 32;
  for (int i = threadIdx.x; i < 2 * d; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  float4 dg[NV], db[NV], gm[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    dg[i] = make_float4(0, 0, 0, 0); db[i] = dg[i];
    gm[i] = __ldg(reinterpret_cast<const float4 *>(gamma + (lane + 32 * i) * 4));
  }
  const int64_t r_begin = ((int64_t)blockIdx.x * wpb + warp) * rows_per_warp;
  for (int64_t row = r_begin; row < min(M, r_begin + rows_per_warp); ++row) {
    const float mu = __ldg(mean + row), rs = __ldg(rstd + row);
    float4 xh[NV], g[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (lane + 32 * i) * 4;
      const float4 dv = __ldg(reinterpret_cast<const float4 *>(dy + row * d + c)), uv = __ldg(reinterpret_cast<const float4 *>(u + row * d + c));
      xh[i] = make_float4((uv.x - mu) * rs, (uv.y - mu) * rs, (uv.z - mu) * rs, (uv.w - mu) * rs);
      g[i] = make_float4(dv.x * gm[i].x, dv.y * gm[i].y, dv.z * gm[i].z, dv.w * gm[i].w);
      s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
      s2 = fmaf(g[i].x, xh[i].x, s2); s2 = fmaf(g[i].y, xh[i].y, s2); s2 = fmaf(g[i].z, xh[i].z, s2); s2 = fmaf(g[i].w, xh[i].w, s2);
      dg[i].x = fmaf(dv.x, xh[i].x, dg[i].x); dg[i].y = fmaf(dv.y, xh[i].y, dg[i].y); dg[i].z = fmaf(dv.z, xh[i].z, dg[i].z); dg[i].w = fmaf(dv.w, xh[i].w, dg[i].w);
      db[i].x += dv.x; db[i].y += dv.y; db[i].z += dv.z; db[i].w += dv.w;
    }
    s1 = warp_sum(s1) / d;
    s2 = warp_sum(s2) / d;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (lane + 32 * i) * 4;
      float4 v = make_float4((g[i].x - s1 - xh[i].x * s2) * rs, (g[i].y - s1 - xh[i].y * s2) * rs, (g[i].z - s1 - xh[i].z * s2) * rs,
                             (g[i].w - s1 - xh[i].w * s2) * rs);
      *reinterpret_cast<float4 *>(du + row * d + c) = v;
      if (da) {
        if (drop.thr) v = ln_drop4(drop, v, (uint64_t)((row0 + row) * d + c));
        *reinterpret_cast<float4 *>(da + row * d + c) = v;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = (lane + 32 * i) * 4;
    atomicAdd(&sm[c], dg[i].x); atomicAdd(&sm[c + 1], dg[i].y); atomicAdd(&sm[c + 2], dg[i].z); atomicAdd(&sm[c + 3], dg[i].w);
    atomicAdd(&sm[d + c], db[i].x); atomicAdd(&sm[d + c + 1], db[i].y); atomicAdd(&sm[d + c + 2], db[i].z); atomicAdd(&sm[d + c + 3], db[i].w);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    atomicAdd(dgamma + c, sm[c]);
    atomicAdd(dbeta + c, sm[d + c]);
  }
}

static bool ln_vec4_ok(const void *a, const void *b, const void *c, const void *e, const void *f, int d) {
  auto al = [](const void *p) { return p == nullptr || ((uintptr_t)p & 15) == 0; };
  return (d == 128 || d == 256 || d == 512) && al(a) && al(b) && al(c) && al(e) && al(f);
}

int ln_fwd(const float *a, const float *res, const float *gamma, const float *beta, float *u, float *y, float *mean,
           float *rstd, int64_t M, int d, const Drop &drop, int64_t row0, cudaStream_t st) {
  if (M == 0) return 0;
  GT_CHECK(d <= 512, "d_model > 512 not supported");
  if (ln_vec4_ok(a, res, gamma, u, y, d) && ((uintptr_t)beta & 15) == 0) {
    const int wpb = 8;
    const unsigned grid = (unsigned)((M + wpb - 1) / wpb);
    { LaunchScope _ls(KC_LN, st);
      if (d == 128) ln_fwd_vec4_kernel<1><<<grid, wpb * 32, 0, st>>>(a, res, gamma, beta, u, y, mean, rstd, M, drop, row0);
      else if (d == 256) ln_fwd_vec4_kernel<2><<<grid, wpb * 32, 0, st>>>(a, res, gamma, beta, u, y, mean, rstd, M, drop, row0);
      else ln_fwd_vec4_kernel<4><<<grid, wpb * 32, 0, st>>>(a, res, gamma, beta, u, y, mean, rstd, M, drop, row0); }
    GT_CUDA(cudaGetLastError());
    return 0;
  }
  const int wpb = 8;
  { LaunchScope _ls(KC_LN, st);
  ln_fwd_kernel<<<(unsigned)((M + wpb - 1) / wpb), wpb * 32, 0, st>>>(a, res, gamma, beta, u, y, mean, rstd, M, d, drop, row0); }
  GT_CUDA(cudaGetLastError());
  return 0;
}

// block = 8 warps x ROWS_PER_WARP rows; per-block dgamma/dbeta partials reduced in smem, then atomics
__global__ void ln_bwd_kernel(const float *__restrict__ dy, const float *__restrict__ u, const float *__restrict__ mean,
                              const float *__restrict__ rstd, const float *__restrict__ gamma, float *du, float *da,
                              float *dgamma, float *dbeta, int64_t M, int d, Drop drop, int64_t row0, int rows_per_warp) {
  drop_resolve(drop);
  extern __shared__ float sm[];          // [2][d] block partials
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, wpb = blockDim.x / 32;
  for (int i = threadIdx.x; i < 2 * d; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  constexpr int MAXV = 16;
  float dg[MAXV], db[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) { dg[i] = 0.f; db[i] = 0.f; }
  const int64_t r_begin = ((int64_t)blockIdx.x * wpb + warp) * rows_per_warp;
  for (int64_t row = r_begin; row < min(M, r_begin + rows_per_warp); ++row) {
    const float mu = mean[row], rs = rstd[row];
    float xh[MAXV], g[MAXV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      int c = lane + i * 32;
      xh[i] = 0.f; g[i] = 0.f;
      if (c < d) {
        float dyv = dy[row * d + c];
        xh[i] = (u[row * d + c] - mu) * rs;
        g[i] = dyv * gamma[c];
        s1 += g[i];
        s2 = fmaf(g[i], xh[i], s2);
        dg[i] = fmaf(dyv, xh[i], dg[i]);
        db[i] += dyv;
      }
    }
    s1 = warp_sum(s1) / d;
    s2 = warp_sum(s2) / d;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      int c = lane + i * 32;
      if (c < d) {
        float v = (g[i] - s1 - xh[i] * s2) * rs;
        du[row * d + c] = v;
        if (da) {
          if (drop.thr) v = drop_keep(drop.key, drop.thr, (uint64_t)((row0 + row) * d + c)) ? v * drop.scale : 0.f;
          da[row * d + c] = v;
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int c = lane + i * 32;
    if (c < d) { atomicAdd(&sm[c], dg[i]); atomicAdd(&sm[d + c], db[i]); }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    atomicAdd(dgamma + c, sm[c]);
    atomicAdd(dbeta + c, sm[d + c]);
  }
}

int ln_bwd(const float *dy, const float *u, const float *mean, const float *rstd, const float *gamma, float *du,
           float *da, float *dgamma, float *dbeta, int64_t M, int d, const Drop &drop, int64_t row0, cudaStream_t st) {
  if (M == 0) return 0;
  GT_CHECK(d <= 512, "d_model > 512 not supported");
  const int wpb = 8, rpw = 16;
  int64_t blocks = (M + wpb * rpw - 1) / (wpb * rpw);
  if (ln_vec4_ok(dy, u, gamma, du, da, d)) {
    { LaunchScope _ls(KC_LN, st);
      if (d == 128) ln_bwd_vec4_kernel<1><<<(unsigned)blocks, wpb * 32, 0, st>>>(dy, u, mean, rstd, gamma, du, da, dgamma, dbeta, M, drop, row0, rpw);
      else if (d == 256) ln_bwd_vec4_kernel<2><<<(unsigned)blocks, wpb * 32, 0, st>>>(dy, u, mean, rstd, gamma, du, da, dgamma, dbeta, M, drop, row0, rpw);
      else ln_bwd_vec4_kernel<4><<<(unsigned)blocks, wpb * 32, 0, st>>>(dy, u, mean, rstd, gamma, du, da, dgamma, dbeta, M, drop, row0, rpw); }
    GT_CUDA(cudaGetLastError());
    return 0;
  }
  { LaunchScope _ls(KC_LN, st);
  ln_bwd_kernel<<<(unsigned)blocks, wpb * 32, 2 * d * sizeof(float), st>>>(dy, u, mean, rstd, gamma, du, da, dgamma, dbeta,
                                                                           M, d, drop, row0, rpw); }
  GT_CUDA(cudaGetLastError());
  return 0;
}

// =============================================================================================
// small element-wise kernels
// =============================================================================================
__global__ void pe_dropout_fwd_kernel(const float *__restrict__ r, const float *__restrict__ pe, float *x0, int64_t n,
                                      int d, Drop drop, int64_t e0) {
  drop_resolve(drop);
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = r[i] + pe[i % ((int64_t)T * d)];
  if (drop.thr) v = drop_keep(drop.key, drop.thr, (uint64_t)(e0 + i)) ? v * drop.scale : 0.f;
  x0[i] = v;
}
int pe_dropout_fwd(const float *r, const float *pe, float *x0, int64_t M, int d, const Drop &drop, int64_t row0, cudaStream_t st) {
  int64_t n = M * d;
  if (n == 0) return 0;
  { LaunchScope _ls(KC_ELEMWISE, st);
  pe_dropout_fwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(r, pe, x0, n, d, drop, row0 * d); }
  GT_CUDA(cudaGetLastError());
  return 0;
}
__global__ void pe_dropout_bwd_kernel(const float *__restrict__ dx0, const float *__restrict__ r, float *g, int64_t n,
                                      Drop drop, int64_t e0) {
  drop_resolve(drop);
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = dx0[i];
  if (drop.thr) v = drop_keep(drop.key, drop.thr, (uint64_t)(e0 + i)) ? v * drop.scale : 0.f;
  g[i] = r[i] > 0.f ? v : 0.f;
}
int pe_dropout_bwd(const float *dx0, const float *r, float *g, int64_t M, int d, const Drop &drop, int64_t row0, cudaStream_t st) {
  int64_t n = M * d;
  if (n == 0) return 0;
  { LaunchScope _ls(KC_ELEMWISE, st);
  pe_dropout_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dx0, r, g, n, drop, row0 * d); }
  GT_CUDA(cudaGetLastError());
  return 0;
}

__global__ void head_activation_kernel(float *hvo, int64_t n, int e_tgt, float thres) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int v9 = e_tgt / 3, c = (int)(i % e_tgt);
  float x = hvo[i];
  if (c < v9) { if (thres >= 0.f) hvo[i] = (1.f / (1.f + expf(-x)) > thres) ? 1.f : 0.f; }
  else if (c < 2 * v9) hvo[i] = 1.f / (1.f + expf(-x));
  else hvo[i] = 0.5f * tanhf(x);
}
int head_activation(float *hvo, int64_t M, int e_tgt, float thres, cudaStream_t st) {
  int64_t n = M * e_tgt;
  if (n == 0) return 0;
  { LaunchScope _ls(KC_ELEMWISE, st);
  head_activation_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(hvo, n, e_tgt, thres); }
  GT_CUDA(cudaGetLastError());
  return 0;
}
__global__ void head_activation_bwd_kernel(const float *__restrict__ d_hvo, const float *__restrict__ hvo, float *dl,
                                           int64_t n, int e_tgt) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int v9 = e_tgt / 3, c = (int)(i % e_tgt);
  float g = d_hvo[i], a = hvo[i];
  if (c < v9) dl[i] = g;
  else if (c < 2 * v9) dl[i] = g * a * (1.f - a);
  else dl[i] = g * (0.5f - 2.f * a * a);       // d/dx 0.5 tanh x = 0.5 (1 - tanh^2) = 0.5 - 2 o^2
}
int head_activation_bwd(const float *d_hvo, const float *hvo, float *dlogits, int64_t M, int e_tgt, cudaStream_t st) {
  int64_t n = M * e_tgt;
  if (n == 0) return 0;
  { LaunchScope _ls(KC_ELEMWISE, st);
  head_activation_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_hvo, hvo, dlogits, n, e_tgt); }
  GT_CUDA(cudaGetLastError());
  return 0;
}

// ---- loss: BGT/models/train.py:9-40 ----------------------------------------------------------
constexpr int LOSS_ROWS_PER_BLOCK = 256;   // token rows per block (one thread per row)
int64_t loss_scratch_floats(int64_t n_seq) {
  int64_t blocks = (n_seq * T + LOSS_ROWS_PER_BLOCK - 1) / LOSS_ROWS_PER_BLOCK;
  return 4 * (blocks > 0 ? blocks : 1);
}
// VT = 9: the reference's nine drum voices, fully unrolled; VT = 0: any voice count V (embedding_size_tgt = 3 V — the reference's
// calculate_loss splits y into thirds, BGT/models/train.py:12-13)
template <int VT>
__global__ void loss_partial_kernel(const float *__restrict__ hvo, const float *__restrict__ y, int64_t M, int Vrt, float penalty,
                                    float *d_hvo, float gscale, float *partials) {
  __shared__ float red[4][LOSS_ROWS_PER_BLOCK / 32];
  const int V = VT > 0 ? VT : Vrt;
  const int64_t row = (int64_t)blockIdx.x * LOSS_ROWS_PER_BLOCK + threadIdx.x;
  float bce = 0.f, mv = 0.f, mo = 0.f, ok = 0.f;
  if (row < M) {
    const float *p = hvo + row * 3 * V, *t = y + row * 3 * V;
    float *g = d_hvo ? d_hvo + row * 3 * V : nullptr;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      float h = p[k], v = p[V + k], o = p[2 * V + k];
      float yh = t[k], yv = t[V + k], yo = t[2 * V + k];
      float w = (yh == 1.f) ? 1.f : penalty;
      float sp = fmaxf(h, 0.f) - h * yh + log1pf(expf(-fabsf(h)));
      bce = fmaf(sp, w, bce);
      float dv = v - yv, dof = o - yo;
      mv = fmaf(dv * dv, w, mv);
      mo = fmaf(dof * dof, w, mo);
      float sg = 1.f / (1.f + expf(-h));
      float hit = sg > 0.5f ? 1.f : 0.f;
      ok += (hit == yh) ? 1.f : 0.f;
      if (g) {
        g[k] = gscale * w * (sg - yh);
        g[V + k] = gscale * 2.f * w * dv;
        g[2 * V + k] = gscale * 2.f * w * dof;
      }
    }
  }
  float vals[4] = {bce, mv, mo, ok};
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float s = warp_sum(vals[q]);
    if (lane == 0) red[q][warp] = s;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    float s = 0.f;
    for (int w = 0; w < LOSS_ROWS_PER_BLOCK / 32; ++w) s += red[threadIdx.x][w];
    partials[(int64_t)blockIdx.x * 4 + threadIdx.x] = s;
  }
}
__global__ void loss_final_kernel(const float *__restrict__ partials, int64_t blocks, int64_t M, int V, float *metrics6) {
  __shared__ double red[4][256];
  griddep_wait();                                      // launch_pdl (common.cuh)
  griddep_launch();
  double acc[4] = {0, 0, 0, 0};
  for (int64_t b = threadIdx.x; b < blocks; b += blockDim.x)
    for (int q = 0; q < 4; ++q) acc[q] += (double)partials[b * 4 + q];
  for (int q = 0; q < 4; ++q) red[q][threadIdx.x] = acc[q];
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s)
      for (int q = 0; q < 4; ++q) red[q][threadIdx.x] += red[q][threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double bce = red[0][0] / (double)M, mv = red[1][0] / (double)M, mo = red[2][0] / (double)M;
    metrics6[0] = (float)(bce + mv + mo);
    metrics6[1] = (float)(red[3][0] / ((double)M * (double)V));
    metrics6[2] = (float)exp(bce);
    metrics6[3] = (float)bce;
    metrics6[4] = (float)mv;
    metrics6[5] = (float)mo;
  }
}
int loss_finalize(const float *partials, int64_t blocks, int64_t M, float *metrics6, cudaStream_t st) {
  { LaunchScope _ls(KC_LOSS, st);
  GT_CUDA(launch_pdl(loss_final_kernel, dim3(1), dim3(256), 0, st, partials, blocks, M, 9, metrics6)); }
  GT_CUDA(cudaGetLastError());
  return 0;
}
int loss_fwd_bwd(const float *hvo, const float *y, int64_t n_seq, float penalty, float *metrics6, float *d_hvo,
                 float grad_scale, float *partials, cudaStream_t st, int n_voices) {
  GT_NVTX("groove.loss");
  int64_t M = n_seq * T;
  GT_CHECK(M > 0, "empty batch");
  GT_CHECK(n_voices >= 1, "loss: embedding_size_tgt must be a positive multiple of 3");
  int64_t blocks = (M + LOSS_ROWS_PER_BLOCK - 1) / LOSS_ROWS_PER_BLOCK;
  { LaunchScope _ls(KC_LOSS, st);
  if (n_voices == 9) loss_partial_kernel<9><<<(unsigned)blocks, LOSS_ROWS_PER_BLOCK, 0, st>>>(hvo, y, M, 9, penalty, d_hvo, grad_scale / (float)M, partials);
  else loss_partial_kernel<0><<<(unsigned)blocks, LOSS_ROWS_PER_BLOCK, 0, st>>>(hvo, y, M, n_voices, penalty, d_hvo, grad_scale / (float)M, partials); }
  GT_CUDA(cudaGetLastError());
  { LaunchScope _ls(KC_LOSS, st);
  GT_CUDA(launch_pdl(loss_final_kernel, dim3(1), dim3(256), 0, st, partials, blocks, M, n_voices, metrics6)); }
  GT_CUDA(cudaGetLastError());
  return 0;
}

// ---- evaluator metrics: GrooveEvaluator/GrooveEvaluator/evaluator.py:189-251 ----------------------
// get_hits_accuracies / get_velocity_errors / get_micro_timing_errors over prediction and ground-truth arrays
// [n_seq, 32, 3 V]: per voice i the mean over examples of (pred == gt).sum(steps) / 32 and of ((gt - pred)^2).mean(steps);
// "Overall" = the same over the flattened [steps x voices] axis.  Every example has the same 32 steps, so each of these is a
// plain mean over (example, step) [and voice]; the kernel produces the 3 V channel sums, HBM-bound at 2 x 32 x 3 V x 4 bytes
// per sequence.  Thread t of a block owns channel t % 3V of row t / 3V: a block reads 9 consecutive rows (243 consecutive
// floats at V = 9) per pass; partial sums per block, fixed-order double accumulation in the final kernel (deterministic).
constexpr int EVAL_BLOCK = 256, EVAL_MAX_BLOCKS = 148 * 8;
static int64_t eval_blocks(int64_t M, int C3) {
  const int rows_per_pass = EVAL_BLOCK / C3;
  int64_t b = (M + (int64_t)rows_per_pass * 8 - 1) / ((int64_t)rows_per_pass * 8);
  return b < 1 ? 1 : (b > EVAL_MAX_BLOCKS ? EVAL_MAX_BLOCKS : b);
}
int64_t eval_scratch_floats(int64_t n_seq, int n_voices) { return eval_blocks(n_seq * T, 3 * n_voices) * 3 * n_voices; }
__global__ void eval_partial_kernel(const float *__restrict__ pred, const float *__restrict__ gt, int64_t M, int V, float *partials) {
  __shared__ float red[EVAL_BLOCK];
  const int C3 = 3 * V, rows_per_pass = EVAL_BLOCK / C3, t = threadIdx.x;
  const int c = t % C3, rl = t / C3;
  float acc = 0.f;
  if (rl < rows_per_pass) {
    for (int64_t r = (int64_t)blockIdx.x * rows_per_pass + rl; r < M; r += (int64_t)gridDim.x * rows_per_pass) {
      const float p = __ldg(pred + r * C3 + c), g = __ldg(gt + r * C3 + c);
      acc += c < V ? (p == g ? 1.f : 0.f) : (g - p) * (g - p);
    }
  }
  red[t] = acc;
  __syncthreads();
  if (t < C3) {
    float s = 0.f;
    for (int j = 0; j < rows_per_pass; ++j) s += red[t + j * C3];
    partials[(int64_t)blockIdx.x * C3 + t] = s;
  }
}
__global__ void eval_final_kernel(const float *__restrict__ partials, int64_t blocks, int64_t M, int V, float *out) {
  __shared__ double ch[96];
  const int C3 = 3 * V, t = threadIdx.x;
  if (t < C3) {
    double s = 0.0;
    for (int64_t b = 0; b < blocks; ++b) s += (double)partials[b * C3 + t];
    ch[t] = s / (double)M;
  }
  __syncthreads();
  if (t < 3) {                                       // out: [hits | velocity | micro-timing] x (V per-voice values, then Overall)
    double o = 0.0;
    for (int i = 0; i < V; ++i) { out[t * (V + 1) + i] = (float)ch[t * V + i]; o += ch[t * V + i]; }
    out[t * (V + 1) + V] = (float)(o / V);
  }
}
int eval_metrics(const float *pred, const float *gt, int64_t n_seq, int n_voices, float *out, float *partials, cudaStream_t st) {
  const int64_t M = n_seq * T;
  GT_CHECK(M > 0, "eval_metrics: empty batch");
  GT_CHECK(n_voices >= 1 && n_voices <= 32, "eval_metrics: n_voices must be in [1, 32]");
  const int64_t blocks = eval_blocks(M, 3 * n_voices);
  { LaunchScope _ls(KC_LOSS, st);
  eval_partial_kernel<<<(unsigned)blocks, EVAL_BLOCK, 0, st>>>(pred, gt, M, n_voices, partials); }
  GT_CUDA(cudaGetLastError());
  { LaunchScope _ls(KC_LOSS, st);
  eval_final_kernel<<<1, 96, 0, st>>>(partials, blocks, M, n_voices, out); }
  GT_CUDA(cudaGetLastError());
  return 0;
}

__global__ void shift_right_kernel(const float *__restrict__ y, float *out, int64_t n, int e) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t row = i / e;
  out[i] = (row % T == 0) ? 0.f : y[i - e];
}
// out[i, :] = data[perm[start + i], :]   (rows of `row` floats, row % 4 == 0)
__global__ void gather_rows_kernel(const float *__restrict__ data, const int64_t *__restrict__ perm, int64_t start, float *out,
                                   int64_t n_rows, int row4, const unsigned long long *start_ptr) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * row4) return;
  if (start_ptr != nullptr) start += (int64_t)*start_ptr;          // graph replay: the row offset lives in device memory
  const int64_t r = i / row4;
  const int c = (int)(i % row4);
  reinterpret_cast<float4 *>(out)[i] = __ldg(reinterpret_cast<const float4 *>(data) + perm[start + r] * row4 + c);
}
int gather_rows(const float *data, const int64_t *perm, int64_t start, float *out, int64_t n_rows, int64_t row_floats, cudaStream_t st,
                const unsigned long long *start_ptr) {
  if (n_rows == 0) return 0;
  GT_CHECK(row_floats % 4 == 0, "gather_rows: row length must be a multiple of 4 floats");
  const int64_t n = n_rows * (row_floats / 4);
  { LaunchScope _ls(KC_ELEMWISE, st);
    gather_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(data, perm, start, out, n_rows, (int)(row_floats / 4), start_ptr); }
  GT_CUDA(cudaGetLastError());
  return 0;
}

int shift_right(const float *y, float *out, int64_t n_seq, int e, cudaStream_t st) {
  int64_t n = n_seq * T * e;
  if (n == 0) return 0;
  { LaunchScope _ls(KC_ELEMWISE, st);
  shift_right_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(y, out, n, e); }
  GT_CUDA(cudaGetLastError());
  return 0;
}

// out[n, i, :] = act(hvo[n, i, :]) with thresholded hits; tgt[n, i+1, :] = same (if i+1 < 32)
__global__ void predict_feedback_kernel(const float *__restrict__ hvo, float *tgt, float *out, int64_t n_seq, int e, int step_i,
                                        float thres) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_seq * e) return;
  int64_t s = i / e;
  int c = (int)(i % e);
  float v = hvo[(s * T + step_i) * e + c];
  if (c < e / 3) v = (1.f / (1.f + expf(-v)) > thres) ? 1.f : 0.f;
  out[(s * T + step_i) * e + c] = v;
  if (step_i + 1 < T) tgt[(s * T + step_i + 1) * e + c] = v;
}
int predict_feedback(const float *hvo, float *tgt, float *out, int64_t n_seq, int e, int step_i, float thres, cudaStream_t st) {
  int64_t n = n_seq * e;
  if (n == 0) return 0;
  { LaunchScope _ls(KC_ELEMWISE, st);
  predict_feedback_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(hvo, tgt, out, n_seq, e, step_i, thres); }
  GT_CUDA(cudaGetLastError());
  return 0;
}

// ---- KV-cached autoregressive decode (gt_predict, encoder-decoder) -----------------------------------------------
// One warp per (sequence, head); lane = key position.  q: one row per sequence; K/V: `nkeys` cached rows per sequence.
__global__ void attn_decode_kernel(const float *__restrict__ q, int64_t ldq, const float *__restrict__ k, const float *__restrict__ v,
                                   int64_t ld_key, int64_t ld_seq, int nkeys, float *__restrict__ o, int64_t ldo, int64_t n_seq,
                                   int H, int dh) {
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int64_t pair = (int64_t)blockIdx.x * (blockDim.x / 32) + warp;
  if (pair >= n_seq * H) return;
  const int64_t seq = pair / H;
  const int head = (int)(pair % H);
  const float *qr = q + seq * ldq + head * dh;
  const float *kr = k + seq * ld_seq + (int64_t)lane * ld_key + head * dh;
  const float *vr = v + seq * ld_seq + (int64_t)lane * ld_key + head * dh;
  const bool live = lane < nkeys;
  float sc = 0.f;
  if (live)
    for (int c = 0; c < dh; ++c) sc = fmaf(qr[c], kr[c], sc);
  sc = live ? sc * rsqrtf((float)dh) : -INFINITY;
  float mx = sc;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  float p = live ? expf(sc - mx) : 0.f;
  float sum = p;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
  p /= sum;
  for (int c = 0; c < dh; ++c) {
    float t = live ? p * vr[c] : 0.f;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
    if (lane == 0) o[seq * ldo + head * dh + c] = t;
  }
}
int attention_decode(const float *q, int64_t ldq, const float *k, const float *v, int64_t ld_key, int64_t ld_seq, int nkeys,
                     float *o, int64_t ldo, int64_t n_seq, int H, int dh, cudaStream_t st) {
  GT_CHECK(nkeys >= 1 && nkeys <= T, "attention_decode: nkeys out of range");
  const int64_t pairs = n_seq * H;
  { LaunchScope _ls(KC_ATTN_FWD, st);
  attn_decode_kernel<<<(unsigned)((pairs + 7) / 8), 256, 0, st>>>(q, ldq, k, v, ld_key, ld_seq, nkeys, o, ldo, n_seq, H, dh); }
  GT_CUDA(cudaGetLastError());
  return 0;
}

// x[n, d] += pe[pos, :]   (InputLayer: relu(linear) + positional row of decode position `pos`)
__global__ void add_pe_row_kernel(float *x, const float *__restrict__ pe_row, int64_t n, int d) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] += pe_row[i % d];
}
int add_pe_row(float *x, const float *pe_row, int64_t n_rows, int d, cudaStream_t st) {
  const int64_t n = n_rows * d;
  { LaunchScope _ls(KC_ELEMWISE, st);
  add_pe_row_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, pe_row, n, d); }
  GT_CUDA(cudaGetLastError());
  return 0;
}

// decode step i: hvo_step[n, e] (hits = raw logits, v/o activated) -> out[n, i, :] with thresholded hits; the same row is
// the decoder's input token of step i+1 (BGT/models/transformer.py:66-72)
__global__ void decode_feedback_kernel(const float *__restrict__ hvo_step, float *tok, float *out, int64_t n_seq, int e, int step_i,
                                       float thres) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_seq * e) return;
  int64_t s = i / e;
  int c = (int)(i % e);
  float v = hvo_step[i];
  if (c < e / 3) v = (1.f / (1.f + expf(-v)) > thres) ? 1.f : 0.f;
  out[(s * T + step_i) * e + c] = v;
  tok[i] = v;
}
int decode_feedback(const float *hvo_step, float *tok, float *out, int64_t n_seq, int e, int step_i, float thres, cudaStream_t st) {
  const int64_t n = n_seq * e;
  { LaunchScope _ls(KC_ELEMWISE, st);
  decode_feedback_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(hvo_step, tok, out, n_seq, e, step_i, thres); }
  GT_CUDA(cudaGetLastError());
  return 0;
}

// ---- optimizers over the flat vectors ---------------------------------------------------------
__global__ void sgd_kernel(float *p, const float *__restrict__ g, int64_t n, float lr, float gs) {
  griddep_wait();                                      // launch_pdl (common.cuh); no early launch_dependents: the layer kernels read
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // parameters before their own wait
  if (i < n) p[i] = p[i] - lr * (g[i] * gs);
}
int sgd_step(float *p, const float *g, int64_t n, float lr, float gs, cudaStream_t st) {
  GT_NVTX("groove.optimizer");
  if (n == 0) return 0;
  { LaunchScope _ls(KC_OPT, st);
  GT_CUDA(launch_pdl(sgd_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, p, g, n, lr, gs)); }
  GT_CUDA(cudaGetLastError());
  return 0;
}
__global__ void adam_kernel(float *p, const float *__restrict__ g, float *m, float *v, int64_t n, float lr, float b1, float b2,
                            float eps, float bc1, float bc2_sqrt, float gs, const unsigned long long *t_ptr) {
  griddep_wait();                                      // launch_pdl (common.cuh); no early launch_dependents (see sgd_kernel)
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (t_ptr != nullptr) {                              // graph replay: bias corrections from the device-resident step counter
    const float t = (float)(*t_ptr + 1ull);
    bc1 = 1.f - powf(b1, t);
    bc2_sqrt = sqrtf(1.f - powf(b2, t));
  }
  // torch/optim/adam.py _single_tensor_adam: exp_avg.lerp_(grad, 1-b1); exp_avg_sq = b2*v + (1-b2) g^2;
  // denom = sqrt(v)/sqrt(bias_correction2) + eps ; p -= (lr / bias_correction1) * m / denom
  float gi = g[i] * gs;
  float mi = m[i] + (gi - m[i]) * (1.f - b1);
  float vi = v[i] * b2 + (1.f - b2) * gi * gi;
  m[i] = mi; v[i] = vi;
  float denom = sqrtf(vi) / bc2_sqrt + eps;
  p[i] = p[i] - (lr / bc1) * (mi / denom);
}
int adam_step(float *p, const float *g, float *m, float *v, int64_t n, float lr, float b1, float b2, float eps, int64_t step,
              float gs, cudaStream_t st, const unsigned long long *t_ptr) {
  GT_NVTX("groove.optimizer");
  if (n == 0) return 0;
  GT_CHECK(step >= 1 || t_ptr != nullptr, "Adam step is 1-based");
  if (step < 1) step = 1;
  double bc1 = 1.0 - pow((double)b1, (double)step), bc2 = 1.0 - pow((double)b2, (double)step);
  { LaunchScope _ls(KC_OPT, st);
  GT_CUDA(launch_pdl(adam_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, p, g, m, v, n, lr, b1, b2, eps, (float)bc1, (float)sqrt(bc2), gs, t_ptr)); }
  GT_CUDA(cudaGetLastError());
  return 0;
}

// end of one replayed step: the step's six metrics go to slot counters[3] % ring_slots of the ring (if there is one), then
// counters[0] (dropout step) += 1, [1] (Adam steps taken) += 1, [2] (row offset into the permutation) += batch, [3] (slot) += 1
__global__ void counter_advance_kernel(unsigned long long *c, unsigned long long batch, const float *metrics6, float *ring,
                                       unsigned long long ring_slots) {
  const unsigned long long slot = c[3];
  if (ring != nullptr && threadIdx.x < 6) ring[(slot % ring_slots) * 6 + threadIdx.x] = metrics6[threadIdx.x];
  __syncthreads();
  if (threadIdx.x == 0) { c[0] += 1ull; c[1] += 1ull; c[2] += batch; c[3] = slot + 1ull; }
}
// forces the (lazily loaded) optimizer kernels into the context: a first launch must not happen inside a stream capture
int optimizer_kernels_warm() {
  cudaFuncAttributes fa;
  GT_CUDA(cudaFuncGetAttributes(&fa, sgd_kernel));
  GT_CUDA(cudaFuncGetAttributes(&fa, adam_kernel));
  GT_CUDA(cudaFuncGetAttributes(&fa, counter_advance_kernel));
  GT_CUDA(cudaFuncGetAttributes(&fa, gather_rows_kernel));
  return 0;
}
int counter_advance(unsigned long long *counters, int64_t batch, const float *metrics6, float *ring, int64_t ring_slots, cudaStream_t st) {
  { LaunchScope _ls(KC_OPT, st);
    counter_advance_kernel<<<1, 32, 0, st>>>(counters, (unsigned long long)batch, metrics6, ring, (unsigned long long)(ring_slots > 0 ? ring_slots : 1)); }
  GT_CUDA(cudaGetLastError());
  return 0;
}

__global__ void debug_mask_kernel(uint32_t key, uint32_t thr, int64_t idx0, int64_t n, uint8_t *keep) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keep[i] = (thr == 0 || drop_keep(key, thr, (uint64_t)(idx0 + i))) ? 1 : 0;
}
int debug_dropout_mask(uint32_t key, uint32_t thr, int64_t idx0, int64_t n, uint8_t *keep, cudaStream_t st) {
  if (n == 0) return 0;
  { LaunchScope _ls(KC_ELEMWISE, st);
  debug_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(key, thr, idx0, n, keep); }
  GT_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace gt
