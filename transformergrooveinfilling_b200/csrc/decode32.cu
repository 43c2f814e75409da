// decode32.cu — one decoder layer of the KV-cached autoregressive predict (runner.cu:predict_decode) as ONE kernel for
// d_model = 32: BGT/models/transformer.py:48-83 decodes a groove one step at a time, so each pass handles ONE token per
// sequence and the per-op formulation was 13 launches of a few microseconds of work per layer and step (2500 launches per
// predict() of a 6-layer decoder).  Here a warp owns a sequence (lane = feature column, d_model = 32 = warp width):
//
//   q | k | v = y Wqkv^T + b        (k | v appended to the layer's self-attention cache at position i)
//   causal self-attention over cache rows 0..i  (online softmax; a head's dh lanes reduce their partial dot products by shuffles)
//   x1 = LN1(y + ctx Wo^T + bo) ;  q = x1 Wq^T + bq ; attention over the 32 cached cross keys / values of the encoder memory
//   x2 = LN2(x1 + ctx Wo^T + bo) ; out = LN3(x2 + relu(x2 W1^T + b1) W2^T + b2)
//
// The layer's fp32 weights are staged once per CTA, transposed ([in][out], odd row stride) so that lane = output column reads
// are conflict-free; the token vector is broadcast from 128 B of per-warp shared memory.  Eval mode only (no dropout).
#include "common.cuh"

namespace gt {

constexpr int DC_WARPS = 16;

struct Dec32Args {
  const float *y_in;
  float *y_out;
  float *kv_self;          // [n, T, 64]: k | v per cached position
  const float *kv_cross;   // [n, T, 64]
  const float *wqkv, *bqkv, *wo, *bo, *wcq, *bcq, *wco, *bco, *w1, *b1, *w2, *b2, *g1, *be1, *g2, *be2, *g3, *be3;
  int64_t n;
  int step, H, F;
  int do_ffn;              // 0: stop after the cross-attention block (y_out = x2); the FFN block then runs on the tensor cores
};

__device__ __forceinline__ float dc_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float dc_layernorm(float u, float g, float b) {
  const float mu = dc_warp_sum(u) * (1.f / 32);
  const float t = u - mu;
  const float rs = rsqrtf(dc_warp_sum(t * t) * (1.f / 32) + LN_EPS);
  return t * rs * g + b;
}
// out[lane] = bias + sum_c Wt[c * ld + lane] * xs[c]   (xs: this warp's 32-float broadcast buffer)
__device__ __forceinline__ float dc_matvec32(const float *Wt, int ld, const float *xs, float bias, int lane) {
  float a0 = bias, a1 = 0.f;
#pragma unroll
  for (int c = 0; c < 32; c += 4) {
    const float4 xv = *reinterpret_cast<const float4 *>(xs + c);
    a0 = fmaf(Wt[(c + 0) * ld + lane], xv.x, a0); a1 = fmaf(Wt[(c + 1) * ld + lane], xv.y, a1);
    a0 = fmaf(Wt[(c + 2) * ld + lane], xv.z, a0); a1 = fmaf(Wt[(c + 3) * ld + lane], xv.w, a1);
  }
  return a0 + a1;
}
// attention of this warp's token over nk cached keys: lane = feature column; q already scaled by log2(e) / sqrt(dh)
__device__ __forceinline__ float dc_attend(const float *kv, int nk, float q, int dh, int lane) {
  float m = -1e30f, l = 0.f, acc = 0.f;
  for (int t0 = 0; t0 < nk; t0 += 8) {            // 16 independent loads in flight per chunk of 8 keys (the rows come from L2 / HBM)
    float kt[8], vt[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const bool ok = t0 + u < nk;
      kt[u] = ok ? kv[(t0 + u) * 64 + lane] : 0.f;
      vt[u] = ok ? kv[(t0 + u) * 64 + 32 + lane] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      float s = q * kt[u];
      for (int o = 1; o < dh; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (t0 + u < nk) {
        const float mn = fmaxf(m, s);
        const float corr = ex2_ftz(m - mn), p = ex2_ftz(s - mn);
        l = fmaf(l, corr, p);
        acc = fmaf(acc, corr, p * vt[u]);
        m = mn;
      }
    }
  }
  return acc / l;
}

__global__ void __launch_bounds__(DC_WARPS * 32, 1) dec32_layer_step_kernel(const Dec32Args a) {
  extern __shared__ __align__(16) float dsm[];
  const int F = a.F, ldq = 97, ld32 = 33, ldf = F + 1;
  float *sWqkv = dsm;                       // [32][97]
  float *sWo = sWqkv + 32 * ldq;            // [32][33]
  float *sWcq = sWo + 32 * ld32, *sWco = sWcq + 32 * ld32;
  float *sW1 = sWco + 32 * ld32;            // [32][F + 1]
  float *sW2 = sW1 + 32 * ldf;              // [F][33]
  float *sB = sW2 + F * ld32;               // bqkv 96 | bo 32 | bcq 32 | bco 32 | b2 32 | g1 be1 g2 be2 g3 be3 (6 x 32) | b1 F
  float *sX = sB + 416 + F;                 // per warp: xs [32] | hs [F]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 96 * 32; i += DC_WARPS * 32) sWqkv[(i & 31) * ldq + (i >> 5)] = a.wqkv[i];
  for (int i = tid; i < 32 * 32; i += DC_WARPS * 32) {
    sWo[(i & 31) * ld32 + (i >> 5)] = a.wo[i];
    sWcq[(i & 31) * ld32 + (i >> 5)] = a.wcq[i];
    sWco[(i & 31) * ld32 + (i >> 5)] = a.wco[i];
  }
  if (a.do_ffn) {
    for (int i = tid; i < F * 32; i += DC_WARPS * 32) sW1[(i & 31) * ldf + (i >> 5)] = a.w1[i];        // W1 [F][32] -> [c][j]
    for (int i = tid; i < 32 * F; i += DC_WARPS * 32) sW2[(i % F) * ld32 + i / F] = a.w2[i];            // W2 [32][F] -> [j][c]
  }
  for (int i = tid; i < 96; i += DC_WARPS * 32) sB[i] = a.bqkv[i];
  if (tid < 32) {
    sB[96 + tid] = a.bo[tid]; sB[128 + tid] = a.bcq[tid]; sB[160 + tid] = a.bco[tid]; sB[192 + tid] = a.b2[tid];
    sB[224 + tid] = a.g1[tid]; sB[256 + tid] = a.be1[tid]; sB[288 + tid] = a.g2[tid]; sB[320 + tid] = a.be2[tid];
    sB[352 + tid] = a.g3[tid]; sB[384 + tid] = a.be3[tid];
  }
  for (int i = tid; i < F; i += DC_WARPS * 32) sB[416 + i] = a.b1[i];
  __syncthreads();
  float *xs = sX + warp * (32 + F), *hs = xs + 32;
  const int dh = 32 / a.H;
  const float qscale = rsqrtf((float)dh) * 1.4426950408889634f;
  for (int64_t s = (int64_t)blockIdx.x * DC_WARPS + warp; s < a.n; s += (int64_t)gridDim.x * DC_WARPS) {
    const float y = a.y_in[s * 32 + lane];
    // ---- causal self-attention ----
    xs[lane] = y;
    __syncwarp();
    const float q = dc_matvec32(sWqkv, ldq, xs, sB[lane], lane) * qscale;
    const float k = dc_matvec32(sWqkv + 32, ldq, xs, sB[32 + lane], lane);
    const float v = dc_matvec32(sWqkv + 64, ldq, xs, sB[64 + lane], lane);
    float *kvs = a.kv_self + s * (T * 64);
    kvs[a.step * 64 + lane] = k;
    kvs[a.step * 64 + 32 + lane] = v;
    __syncwarp();
    float ctx = dc_attend(kvs, a.step + 1, q, dh, lane);
    xs[lane] = ctx;
    __syncwarp();
    const float x1 = dc_layernorm(y + dc_matvec32(sWo, ld32, xs, sB[96 + lane], lane), sB[224 + lane], sB[256 + lane]);
    __syncwarp();
    // ---- cross-attention over the encoder memory ----
    xs[lane] = x1;
    __syncwarp();
    const float qc = dc_matvec32(sWcq, ld32, xs, sB[128 + lane], lane) * qscale;
    ctx = dc_attend(a.kv_cross + s * (T * 64), T, qc, dh, lane);
    __syncwarp();
    xs[lane] = ctx;
    __syncwarp();
    const float x2 = dc_layernorm(x1 + dc_matvec32(sWco, ld32, xs, sB[160 + lane], lane), sB[288 + lane], sB[320 + lane]);
    __syncwarp();
    if (!a.do_ffn) {                                  // bf16 mode: the FFN block of these tokens runs in tc_layer_fwd (TC_MODE_FFN)
      a.y_out[s * 32 + lane] = x2;
      __syncwarp();
      continue;
    }
    // ---- feed-forward ----
    xs[lane] = x2;
    __syncwarp();
    for (int j0 = 0; j0 < F; j0 += 32) hs[j0 + lane] = fmaxf(dc_matvec32(sW1 + j0, ldf, xs, sB[416 + j0 + lane], lane), 0.f);
    __syncwarp();
    float o0 = sB[192 + lane], o1 = 0.f;
#pragma unroll 4
    for (int j = 0; j < F; j += 4) {
      const float4 hv = *reinterpret_cast<const float4 *>(hs + j);
      o0 = fmaf(sW2[(j + 0) * ld32 + lane], hv.x, o0); o1 = fmaf(sW2[(j + 1) * ld32 + lane], hv.y, o1);
      o0 = fmaf(sW2[(j + 2) * ld32 + lane], hv.z, o0); o1 = fmaf(sW2[(j + 3) * ld32 + lane], hv.w, o1);
    }
    a.y_out[s * 32 + lane] = dc_layernorm(x2 + o0 + o1, sB[352 + lane], sB[384 + lane]);
    __syncwarp();
  }
}

bool dec32_supported(const gt_config &c) {
  return c.d_model == 32 && c.n_dec > 0 && c.dim_ff % 32 == 0 && c.dim_ff >= 32 && c.dim_ff <= 512 && 32 % c.nhead == 0;
}

int dec32_layer_step(const gt_config &c, const LayerP &p, const float *P, const float *y_in, float *y_out, float *kv_self,
                     const float *kv_cross, int64_t n, int step, bool do_ffn, cudaStream_t st) {
  GT_CHECK(dec32_supported(c), "dec32_layer_step: configuration not supported");
  Dec32Args a;
  a.y_in = y_in; a.y_out = y_out; a.kv_self = kv_self; a.kv_cross = kv_cross;
  a.wqkv = P + p.sa.w_in; a.bqkv = P + p.sa.b_in; a.wo = P + p.sa.w_out; a.bo = P + p.sa.b_out;
  a.wcq = P + p.ca.w_in; a.bcq = P + p.ca.b_in; a.wco = P + p.ca.w_out; a.bco = P + p.ca.b_out;
  a.w1 = P + p.w1; a.b1 = P + p.b1; a.w2 = P + p.w2; a.b2 = P + p.b2;
  a.g1 = P + p.g1; a.be1 = P + p.be1; a.g2 = P + p.g2; a.be2 = P + p.be2; a.g3 = P + p.g3; a.be3 = P + p.be3;
  a.n = n; a.step = step; a.H = c.nhead; a.F = c.dim_ff; a.do_ffn = do_ffn ? 1 : 0;
  const int F = c.dim_ff;
  const size_t smem = (size_t)(32 * 97 + 3 * 32 * 33 + 32 * (F + 1) + F * 33 + 416 + F + DC_WARPS * (32 + F)) * sizeof(float);
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  int64_t grid = (n + DC_WARPS - 1) / DC_WARPS;
  if (grid > sms) grid = sms;
  GT_CUDA(cudaFuncSetAttribute(dec32_layer_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  { LaunchScope _ls(KC_ATTN_FWD, st);
    dec32_layer_step_kernel<<<(unsigned)grid, DC_WARPS * 32, smem, st>>>(a); }
  GT_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace gt
