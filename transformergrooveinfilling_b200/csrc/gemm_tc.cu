// gemm_tc.cu — generic mixed-precision GEMM on tcgen05 for every shape the fused layer kernels do not cover
// (d_model other than 32 / 256, head_dim 64 / 128 at d_model = 256, and the whole encoder-decoder model).
//
//   C[m,n] = epi( sum_k A(m,k) * B(n,k) )      A(m,k) = A[m*sam + k*sak],  B(n,k) = B[n*sbn + k*sbk]
//
// Same contract as gemm_f32 (kernels_simt.cu): fp32 operands and fp32 results in HBM, the same fused epilogue
// (bias / ReLU / positional encoding / dropout / ReLU-mask / residual / accumulate / split-K atomics).  The
// operands are rounded to bf16 on their way into shared memory and contracted by UMMA with an fp32 accumulator
// in TMEM, i.e. exactly the arithmetic of the fused bf16 layer kernels.
//
// One CTA = one 128 x BN tile of C and one K range; two CTAs per SM overlap one CTA's epilogue with the other's main loop.  Three
// ways of getting the operands into the canonical no-swizzle core-matrix images UMMA reads (launch<BN>() picks; GT_GEMM_PRE):
//   * both operands pre-imaged (default whenever the problem has >= 32 CTAs and the scratch fits): gemm_tc_aimg_kernel /
//     gemm_tc_bimg_kernel convert A and B once per GEMM, at full occupancy and HBM speed, into contiguous 64-deep tile images;
//     the main kernel's K loop is then two bulk-TMA copies and four UMMAs per block, issued by one thread, two stages deep;
//   * weight operand pre-imaged only (scratch too small for A): B by bulk TMA, A staged by the CTA's threads as below;
//   * in-kernel staging (small problems, stream capture): per 64-deep K block all 256 threads load fp32 from global (32-byte
//     chunks, 8 rows x 4 chunks per warp so that the 16-byte st.shared of the packed chunk is conflict-free), convert and write
//     the image; one thread issues the 4 UMMAs of the block and commits them to the stage's mbarrier.
// An operand whose contraction index is NOT the contiguous one in memory (the data-gradient and weight-gradient forms) is
// imaged as its transpose and consumed as an MN-major operand — no transposition in registers or shared memory.  The epilogue
// drains TMEM per warp (lane = row), applies bias / ReLU / dropout, turns the chunk through shared memory and writes float4
// row segments with the mask / residual / accumulate / atomic options applied per quad.
#include "common.cuh"
#include "umma.cuh"
#include <algorithm>

namespace gt {
using namespace umma;

namespace {

constexpr int GBM = 128, GBK = 64, GSTG = 2, GTHREADS = 256;

struct GemmTcArgs {
  const float *A, *B;
  float *C;
  int64_t sam, sak, sbn, sbk, ldc, M, N, K, kchunk;
  int n_tiles;
  int a_vec, b_vec;                 // 16-byte loads are legal on A / B
  const uint8_t *b_img;             // pre-built bf16 images of B, one per (n-tile, K block) in that order (gemm_tc_bimg_kernel), or null
  const uint8_t *a_img;             // same for A, one per (m-tile, K block) (gemm_tc_aimg_kernel); only together with b_img
  int nkb_img;                      // K blocks of the whole contraction (per n-tile / m-tile in the images)
  int split;                        // 1: fp32 results on the tensor cores — every image block is a bf16 [x0 | x1 | x2] triple (see split_pair)
  GemmEpi e;
};

// image rows r0 + 8-row blocks, CH = 1 << LCH chunks of 8 columns; item i of the tile:
__device__ __forceinline__ void item_rc(int i, int lch, int &r, int &c) {
  const int b = i >> 3;
  c = b & ((1 << lch) - 1);
  r = (i & 7) | ((b >> lch) << 3);
}

// fp32 tile -> registers.  Image row i = src row (row0 + i), image column j = src column (col0 + j); columns are the
// contiguous index in memory, `ld` floats between rows.  Out-of-range elements read as zero.
template <int ITEMS>
__device__ __forceinline__ void tile_load(float4 (&v)[ITEMS][2], const float *__restrict__ src, int64_t ld, int rows, int lch,
                                          int64_t row0, int64_t rows_end, int64_t col0, int64_t cols_end, bool vec, int tid) {
#pragma unroll
  for (int it = 0; it < ITEMS; ++it) {
    const int i = tid + it * GTHREADS;
    v[it][0] = make_float4(0.f, 0.f, 0.f, 0.f);
    v[it][1] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i >= (rows << lch)) continue;
    int r, c;
    item_rc(i, lch, r, c);
    const int64_t gr = row0 + r, gc = col0 + (int64_t)c * 8;
    if (gr >= rows_end || gc >= cols_end) continue;
    const float *p = src + gr * ld + gc;
    if (vec && gc + 8 <= cols_end) {
      v[it][0] = __ldg(reinterpret_cast<const float4 *>(p));
      v[it][1] = __ldg(reinterpret_cast<const float4 *>(p) + 1);
    } else {
      float t[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) t[j] = (gc + j < cols_end) ? __ldg(p + j) : 0.f;
      v[it][0] = make_float4(t[0], t[1], t[2], t[3]);
      v[it][1] = make_float4(t[4], t[5], t[6], t[7]);
    }
  }
}

template <int ITEMS>
__device__ __forceinline__ void tile_store(const float4 (&v)[ITEMS][2], uint8_t *img, int rows, int lch, int tid) {
#pragma unroll
  for (int it = 0; it < ITEMS; ++it) {
    const int i = tid + it * GTHREADS;
    if (i >= (rows << lch)) continue;
    int r, c;
    item_rc(i, lch, r, c);
    *reinterpret_cast<uint4 *>(img + kmajor_off(r, c * 8, rows)) =
        make_uint4(pack_bf16(v[it][0].x, v[it][0].y), pack_bf16(v[it][0].z, v[it][0].w), pack_bf16(v[it][1].x, v[it][1].y),
                   pack_bf16(v[it][1].z, v[it][1].w));
  }
}

// Split form (GT_PREC_FP32_TC): x = x0 + x1 + x2 exactly, x0 = bf16(x), x1 = bf16(x - x0), x2 = bf16(x - x0 - x1) (8 + 8 + 8
// significand bits; the subtractions are exact in fp32).  The main kernel contracts the six products of weight >= 2^-18 —
// x0.y0, x0.y1, x1.y0, x1.y1, x0.y2, x2.y0 — into the same fp32 accumulator: what is dropped is <= 2^-26 of a product, below the
// accumulator's own rounding, i.e. fp32 results (the 1e-4 parity mode) at a sixth of the bf16 tensor rate instead of the FFMA rate.
// (Two-term splits — three products, 2^-16 — were measured first: 1e-5 gradient errors, not enough for 20-step trajectories.)
__device__ __forceinline__ void split_pair(float x, float y, uint32_t &p0, uint32_t &p1, uint32_t &p2) {
  p0 = pack_bf16(x, y);
  x -= __uint_as_float(p0 << 16); y -= __uint_as_float(p0 & 0xFFFF0000u);
  p1 = pack_bf16(x, y);
  x -= __uint_as_float(p1 << 16); y -= __uint_as_float(p1 & 0xFFFF0000u);
  p2 = pack_bf16(x, y);
}
template <int ITEMS>
__device__ __forceinline__ void tile_store_split(const float4 (&v)[ITEMS][2], uint8_t *img, uint32_t img_bytes, int rows, int lch, int tid) {
#pragma unroll
  for (int it = 0; it < ITEMS; ++it) {
    const int i = tid + it * GTHREADS;
    if (i >= (rows << lch)) continue;
    int r, c;
    item_rc(i, lch, r, c);
    uint4 h, m, l;
    split_pair(v[it][0].x, v[it][0].y, h.x, m.x, l.x);
    split_pair(v[it][0].z, v[it][0].w, h.y, m.y, l.y);
    split_pair(v[it][1].x, v[it][1].y, h.z, m.z, l.z);
    split_pair(v[it][1].z, v[it][1].w, h.w, m.w, l.w);
    uint8_t *dst = img + kmajor_off(r, c * 8, rows);
    *reinterpret_cast<uint4 *>(dst) = h;
    *reinterpret_cast<uint4 *>(dst + img_bytes) = m;
    *reinterpret_cast<uint4 *>(dst + 2u * img_bytes) = l;
  }
}

__device__ __forceinline__ int ilog2(int x) { return 31 - __clz(x); }

// Fast path of tile_load / tile_store for a tile that lies wholly inside its operand, with 16-byte loads legal and every K block
// full.  Item it of thread tid sits at image row r0 + it * (256 >> lch), chunk c = (tid >> 3) & (CH - 1) (32 is a multiple of
// CH <= 32, so the chunk does not depend on it): the global pointer and the image offset advance by constants per item, and the
// pointer by a constant per K block — no per-item index arithmetic or bounds checks in the main loop.
struct FastOp {
  const float *p;          // this thread's first item of the current K block
  int64_t istep;           // floats between consecutive items
  int64_t kadv;            // floats between consecutive K blocks
  uint32_t soff, sstep;    // image byte offset of the first item, bytes between items
};
__device__ __forceinline__ FastOp fast_op(const float *src, int64_t ld, int rows, int lch, int64_t row0, int64_t col0, bool k_is_row, int tid) {
  FastOp f;
  const int c = (tid >> 3) & ((1 << lch) - 1);
  const int r0 = (tid & 7) | (((tid >> 3) >> lch) << 3);
  const int rstep = GTHREADS >> lch;
  f.p = src + (row0 + r0) * ld + col0 + (int64_t)c * 8;
  f.istep = (int64_t)rstep * ld;
  f.kadv = k_is_row ? (int64_t)GBK * ld : (int64_t)GBK;
  f.soff = kmajor_off(r0, c * 8, rows);
  f.sstep = (uint32_t)(rstep >> 3) * 128u;
  return f;
}
template <int ITEMS>
__device__ __forceinline__ void fast_load(float4 (&v)[ITEMS][2], const FastOp &f) {
#pragma unroll
  for (int it = 0; it < ITEMS; ++it) {
    const float4 *q = reinterpret_cast<const float4 *>(f.p + it * f.istep);
    v[it][0] = __ldg(q);
    v[it][1] = __ldg(q + 1);
  }
}
template <int ITEMS>
__device__ __forceinline__ void fast_store(const float4 (&v)[ITEMS][2], uint8_t *img, const FastOp &f) {
#pragma unroll
  for (int it = 0; it < ITEMS; ++it)
    *reinterpret_cast<uint4 *>(img + f.soff + it * f.sstep) =
        make_uint4(pack_bf16(v[it][0].x, v[it][0].y), pack_bf16(v[it][0].z, v[it][0].w), pack_bf16(v[it][1].x, v[it][1].y),
                   pack_bf16(v[it][1].z, v[it][1].w));
}

// B as the weight operand of a forward / data-gradient GEMM is tiny and shared by every one of the thousands of M tiles, yet it
// was two thirds of each CTA's staging work (256 of the 384 rows converted per K block).  This pre-pass converts it ONCE per
// GEMM into the exact shared-memory images (zero-padded at the edges), one contiguous BN x 64 block per (n-tile, K block); the
// main kernel then fetches a block with a single bulk-TMA copy — no registers, no conversion, no bounds logic for B.
template <int BN>
__global__ void __launch_bounds__(GTHREADS) gemm_tc_bimg_kernel(const GemmTcArgs g, uint8_t *img) {
  constexpr int BI = (BN * GBK / 8 + GTHREADS - 1) / GTHREADS;
  constexpr uint32_t B_BYTES = BN * GBK * 2;
  const int tid = threadIdx.x;
  const int64_t n0 = (int64_t)blockIdx.x * BN, k0 = (int64_t)blockIdx.y * GBK;
  const bool b_mn = g.sbk != 1;
  const int b_rows = b_mn ? GBK : BN, b_lch = b_mn ? ilog2(BN / 8) : 3;
  const int64_t b_ld = b_mn ? g.sbk : g.sbn;
  float4 rb[BI][2];
  if (b_mn) tile_load<BI>(rb, g.B, b_ld, b_rows, b_lch, k0, g.K, n0, g.N, g.b_vec != 0, tid);
  else tile_load<BI>(rb, g.B, b_ld, b_rows, b_lch, n0, g.N, k0, g.K, g.b_vec != 0, tid);
  if (g.split) tile_store_split<BI>(rb, img + ((size_t)blockIdx.x * gridDim.y + blockIdx.y) * 3u * B_BYTES, B_BYTES, b_rows, b_lch, tid);
  else tile_store<BI>(rb, img + ((size_t)blockIdx.x * gridDim.y + blockIdx.y) * B_BYTES, b_rows, b_lch, tid);
}

// The same for A (activations): one streaming pass fp32 -> bf16 images, after which the main kernel's K loop is two bulk-TMA
// copies and four UMMAs per block issued by one thread — the conversion runs at full occupancy here instead of latency-bound
// inside a 2-CTA-per-SM GEMM, and an A block shared by several n-tiles is converted once.
__global__ void __launch_bounds__(GTHREADS) gemm_tc_aimg_kernel(const GemmTcArgs g, uint8_t *img) {
  constexpr int AI = GBM * GBK / 8 / GTHREADS;
  constexpr uint32_t A_BYTES = GBM * GBK * 2;
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * GBM, k0 = (int64_t)blockIdx.y * GBK;
  const bool a_mn = g.sak != 1;
  const int a_rows = a_mn ? GBK : GBM, a_lch = a_mn ? 4 : 3;
  const int64_t a_ld = a_mn ? g.sak : g.sam;
  float4 ra[AI][2];
  if (a_mn) tile_load<AI>(ra, g.A, a_ld, a_rows, a_lch, k0, g.K, m0, g.M, g.a_vec != 0, tid);
  else tile_load<AI>(ra, g.A, a_ld, a_rows, a_lch, m0, g.M, k0, g.K, g.a_vec != 0, tid);
  if (g.split) tile_store_split<AI>(ra, img + ((size_t)blockIdx.x * gridDim.y + blockIdx.y) * 3u * A_BYTES, A_BYTES, a_rows, a_lch, tid);
  else tile_store<AI>(ra, img + ((size_t)blockIdx.x * gridDim.y + blockIdx.y) * A_BYTES, a_rows, a_lch, tid);
}

template <int BN>
__global__ void __launch_bounds__(GTHREADS, 2) gemm_tc_kernel(const GemmTcArgs g) {
  constexpr uint32_t A_BYTES = GBM * GBK * 2, B_BYTES = BN * GBK * 2, STG = A_BYTES + B_BYTES;
  constexpr int AI = GBM * GBK / 8 / GTHREADS, BI = (BN * GBK / 8 + GTHREADS - 1) / GTHREADS;
  // split mode (BN <= 128) keeps a second accumulator for the five correction products: columns [BN, 2 BN)
  constexpr uint32_t TCOLS = BN <= 128 ? (2 * BN < 32 ? 32 : 2 * BN) : BN;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_free[GSTG], bar_b[GSTG], bar_done;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t m0 = (int64_t)(blockIdx.x / g.n_tiles) * GBM;
  const int64_t n0 = (int64_t)(blockIdx.x % g.n_tiles) * BN;
  const int64_t k_begin = (int64_t)blockIdx.y * g.kchunk;
  const int64_t k_end = min(g.K, k_begin + g.kchunk);
  const bool a_mn = g.sak != 1, b_mn = g.sbk != 1;
  const int nrem = (int)min((int64_t)BN, g.N - n0);
  const int nn = (nrem + 15) & ~15;                       // UMMA N of this tile

  if (warp == 0) tmem_alloc(&tmem_slot, TCOLS);
  if (tid == 0) {
    for (int i = 0; i < GSTG; ++i) { mbar_init(&bar_free[i], 1); mbar_init(&bar_b[i], 1); }
    mbar_init(&bar_done, 1);
    fence_mbar_init();
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = make_idesc_bf16(GBM, nn, a_mn ? 1 : 0, b_mn ? 1 : 0);
  // image geometry: K-major source -> image [rows = M/N extent] x [64 k columns]; MN-major source -> image [64 k rows] x [extent]
  const int a_rows = a_mn ? GBK : GBM, a_lch = a_mn ? 4 : 3;
  const int b_rows = b_mn ? GBK : BN, b_lch = b_mn ? ilog2(BN / 8) : 3;
  const int64_t a_ld = a_mn ? g.sak : g.sam, b_ld = b_mn ? g.sbk : g.sbn;
  const uint32_t a_lbo = a_mn ? 128u : (GBM / 8) * 128u, a_sbo = a_mn ? (GBK / 8) * 128u : 128u;
  const uint32_t b_lbo = b_mn ? 128u : (BN / 8) * 128u, b_sbo = b_mn ? (GBK / 8) * 128u : 128u;
  const uint32_t a_kstep = a_mn ? 256u : 2u * (GBM / 8) * 128u, b_kstep = b_mn ? 256u : 2u * (BN / 8) * 128u;   // bytes per k16

  const int nkb = (int)((k_end - k_begin + GBK - 1) / GBK);
  if (g.a_img != nullptr) {
    // both operands pre-imaged: thread 0 alone runs the K loop (copies of block kb + 2 are issued as soon as the MMAs of block
    // kb have retired), the other warps go straight to the epilogue's wait; its own warp parks at the __syncwarp instead of
    // spinning on the barrier next to it
    if (warp == 0) {
    if (lane == 0) {
      const int64_t kb0 = k_begin / GBK;
      // split mode: block kb of an operand is a triple of images [x0 | x1 | x2]; a stage holds both triples (1 CTA per SM, BN <= 128)
      // and is contracted in six passes, smallest products first
      const int im = g.split ? 3 : 1, npass = g.split ? 6 : 1;
      const uint32_t stg = (uint32_t)im * STG;
      const uint8_t *a_src = g.a_img + ((size_t)(blockIdx.x / g.n_tiles) * g.nkb_img + kb0) * im * A_BYTES;
      const uint8_t *b_src = g.b_img + ((size_t)(blockIdx.x % g.n_tiles) * g.nkb_img + kb0) * im * B_BYTES;
      auto issue = [&](int kb) {
        const int s = kb % GSTG;
        uint8_t *sA = smem + (uint32_t)s * stg;
        mbar_expect_tx(&bar_b[s], stg);
        for (int i = 0; i < im; ++i) {
          tma_load_1d(sA + (uint32_t)i * A_BYTES, a_src + ((size_t)kb * im + i) * A_BYTES, A_BYTES, &bar_b[s]);
          tma_load_1d(sA + (uint32_t)im * A_BYTES + (uint32_t)i * B_BYTES, b_src + ((size_t)kb * im + i) * B_BYTES, B_BYTES, &bar_b[s]);
        }
      };
      for (int kb = 0; kb < nkb && kb < GSTG; ++kb) issue(kb);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % GSTG;
        mbar_wait(&bar_b[s], (uint32_t)(kb / GSTG) & 1u);
        fence_after_sync();
        const int64_t krem = k_end - (k_begin + (int64_t)kb * GBK);
        const int nk16 = krem >= GBK ? GBK / 16 : (int)((krem + 15) / 16);
        const uint32_t aA = smem_u32(smem + (uint32_t)s * stg), aB = aA + (uint32_t)im * A_BYTES;
        for (int p = 0; p < npass; ++p) {
          // (A image, B image) of pass p: (2,0) (0,2) (1,1) (1,0) (0,1) (0,0); plain mode: (0,0)
          const uint32_t ia = g.split ? ((0x001102u >> (4 * p)) & 3u) : 0u, ib = g.split ? ((0x010120u >> (4 * p)) & 3u) : 0u;
          // The tensor core TRUNCATES when it adds into the fp32 accumulator: every UMMA costs up to one ulp of the accumulator,
          // always towards zero (measured: 96 accumulations of a K = 256 contraction left 1e-5).  The x0.y0 products therefore have
          // the main accumulator to themselves and the five corrections (2^-9 of it and less) add up in a second one, whose
          // truncation is invisible; the epilogue adds the two in fp32.
          const bool corr = g.split && p < 5;
          for (int k = 0; k < nk16; ++k)
            mma_bf16_ss(tmem + (corr ? (uint32_t)BN : 0u), make_desc(aA + ia * A_BYTES + (uint32_t)k * a_kstep, a_lbo, a_sbo),
                        make_desc(aB + ib * B_BYTES + (uint32_t)k * b_kstep, b_lbo, b_sbo), idesc,
                        (corr ? (kb | p | k) : (kb | k)) > 0 ? 1u : 0u);
        }
        mma_commit(&bar_free[s]);
        if (kb + 1 == nkb) mma_commit(&bar_done);
        if (kb + GSTG < nkb) {
          mbar_wait(&bar_free[s], (uint32_t)(kb / GSTG) & 1u);
          issue(kb + GSTG);
        }
      }
    }
    __syncwarp();
    }
  } else {
  // interior tiles (the common case: every dimension of the reference's models is a multiple of the tile) take the fast path
  const bool k_full = ((k_end - k_begin) % GBK) == 0;
  const bool a_fast = g.a_vec != 0 && k_full && m0 + GBM <= g.M;
  const bool b_fast = g.b_vec != 0 && k_full && nrem == BN && (BN * GBK / 8) % GTHREADS == 0;
  const bool b_pre = g.b_img != nullptr;                  // B blocks come ready-made by bulk TMA (never with split-K: k_begin = 0)
  const uint8_t *b_src = b_pre ? g.b_img + (size_t)(blockIdx.x % g.n_tiles) * g.nkb_img * B_BYTES : nullptr;
  FastOp fa = a_mn ? fast_op(g.A, a_ld, a_rows, a_lch, k_begin, m0, true, tid) : fast_op(g.A, a_ld, a_rows, a_lch, m0, k_begin, false, tid);
  FastOp fb = b_mn ? fast_op(g.B, b_ld, b_rows, b_lch, k_begin, n0, true, tid) : fast_op(g.B, b_ld, b_rows, b_lch, n0, k_begin, false, tid);
  for (int kb = 0; kb < nkb; ++kb) {
    const int s = kb % GSTG;
    const int64_t k0 = k_begin + (int64_t)kb * GBK;
    float4 ra[AI][2], rb[BI][2];
    if (b_pre && tid == 0) {                                // first thing in the block: the copy lands while A is fetched and converted
      if (kb >= GSTG) mbar_wait(&bar_free[s], (uint32_t)((kb / GSTG) - 1) & 1u);
      mbar_expect_tx(&bar_b[s], B_BYTES);
      tma_load_1d(smem + (uint32_t)s * STG + A_BYTES, b_src + (size_t)kb * B_BYTES, B_BYTES, &bar_b[s]);
    }
    if (a_fast) { fast_load<AI>(ra, fa); fa.p += fa.kadv; }
    else if (a_mn) tile_load<AI>(ra, g.A, a_ld, a_rows, a_lch, k0, k_end, m0, g.M, g.a_vec != 0, tid);
    else tile_load<AI>(ra, g.A, a_ld, a_rows, a_lch, m0, g.M, k0, k_end, g.a_vec != 0, tid);
    if (b_pre) {}
    else if (b_fast) { fast_load<BI>(rb, fb); fb.p += fb.kadv; }
    else if (b_mn) tile_load<BI>(rb, g.B, b_ld, b_rows, b_lch, k0, k_end, n0, g.N, g.b_vec != 0, tid);
    else tile_load<BI>(rb, g.B, b_ld, b_rows, b_lch, n0, g.N, k0, k_end, g.b_vec != 0, tid);
    if (kb >= GSTG) mbar_wait(&bar_free[s], (uint32_t)((kb / GSTG) - 1) & 1u);    // the MMAs that read this stage have retired
    uint8_t *sA = smem + (uint32_t)s * STG, *sB = sA + A_BYTES;
    if (a_fast) fast_store<AI>(ra, sA, fa); else tile_store<AI>(ra, sA, a_rows, a_lch, tid);
    if (b_pre) {}
    else if (b_fast) fast_store<BI>(rb, sB, fb);
    else tile_store<BI>(rb, sB, b_rows, b_lch, tid);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      fence_after_sync();
      if (b_pre) mbar_wait(&bar_b[s], (uint32_t)(kb / GSTG) & 1u);
      const int64_t krem = k_end - k0;
      const int nk16 = krem >= GBK ? GBK / 16 : (int)((krem + 15) / 16);
      const uint32_t aA = smem_u32(sA), aB = smem_u32(sB);
      for (int k = 0; k < nk16; ++k)
        mma_bf16_ss(tmem, make_desc(aA + (uint32_t)k * a_kstep, a_lbo, a_sbo), make_desc(aB + (uint32_t)k * b_kstep, b_lbo, b_sbo), idesc,
                    (kb | k) > 0 ? 1u : 0u);
      mma_commit(&bar_free[s]);
      if (kb + 1 == nkb) mma_commit(&bar_done);
    }
  }
  }
  mbar_wait(&bar_done, 0);
  fence_after_sync();

  // ---- epilogue ----
  // Warp w drains TMEM lanes 32 (w & 3) .. in 32-column chunks (w >> 2), (w >> 2) + 2, ...  Phase 1 (lane = row, the TMEM
  // mapping): bias, ReLU, dropout with one hash per quad.  The chunk then goes through a padded per-warp shared-memory
  // buffer (the operand stages are free: every MMA has retired) so that phase 2 runs with lane = column: every access to
  // C, the residual and the ReLU mask is a 128-byte row segment per warp instruction instead of 32 scattered 16-byte ones.
  const GemmEpi &e = g.e;
  Drop edrop = g.e.drop;
  drop_resolve(edrop);                                    // graph replay: key from the device step counter
  const int q = warp & 3;
  const int64_t mw0 = m0 + q * 32;                        // first row of this warp
  const bool first_split = blockIdx.y == 0;
  // vector form of phase 2 (every case but the positional-encoding epilogue of the input layer, given 16-byte alignment): lane
  // = (row i & 3, column quad): one 16-byte access per operand per lane, four whole 128-byte row segments per warp instruction
  auto al16 = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  const bool vec_epi = !e.pe && (g.N & 3) == 0 && (g.ldc & 3) == 0 && al16(g.C) && (!e.bias || al16(e.bias)) &&
                       (!e.residual || ((e.ld_res & 3) == 0 && al16(e.residual))) && (!e.mask_pos || ((e.ld_mask & 3) == 0 && al16(e.mask_pos)));
  constexpr int BS = 36;                                  // floats per buffer row: 16-byte aligned rows, conflict-free float4 access
  float *buf = reinterpret_cast<float *>(smem) + warp * (32 * BS);
  for (int cb = (warp >> 2) * 32; cb < nn; cb += 64) {
    float v[32];
    tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)cb, v);
    tmem_ld_wait();
    if (g.split) {                                        // + the correction accumulator
      float c2[32];
      tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(BN + cb), c2);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] += c2[j];
    }
    const int64_t nb = n0 + cb;
    {
      const int64_t m = mw0 + lane;
      if (e.bias && first_split) {
        if (vec_epi) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            if (nb + j < g.N) {
              const float4 b4 = __ldg(reinterpret_cast<const float4 *>(e.bias + nb + j));
              v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
            }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) if (nb + j < g.N) v[j] += __ldg(e.bias + nb + j);
        }
      }
      if (e.relu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
      }
      if (edrop.thr && !e.pe && m < g.M) {
#pragma unroll
        for (int j4 = 0; j4 < 32; j4 += 4) {
          const int64_t n = nb + j4;
          if (n >= g.N) break;
          const uint64_t idx = (uint64_t)((e.drop_row0 + m) * g.N + n);
          if (n + 4 <= g.N && (idx & 3) == 0) {             // the four elements are one quad of the site: one hash
            const uint64_t wq = idx >> 2;
            uint32_t lo, hi;
            hash_quad((uint32_t)wq ^ ((uint32_t)(wq >> 32) * 0x85EBCA6Bu), edrop.key, lo, hi);
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j4 + j] = quad_keep(lo, hi, j, edrop.thr) ? v[j4 + j] * edrop.scale : 0.f;
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (n + j < g.N) v[j4 + j] = drop_keep(edrop.key, edrop.thr, idx + j) ? v[j4 + j] * edrop.scale : 0.f;
          }
        }
      }
    }
    __syncwarp();                                           // the previous chunk's phase 2 has finished reading buf
#pragma unroll
    for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4 *>(&buf[lane * BS + j]) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    __syncwarp();
    const int rows = (int)min((int64_t)32, g.M - mw0);
    if (vec_epi) {
      const int cq = (lane & 7) * 4;
      const int64_t n = nb + cq;
      if (n < g.N) {                                        // N % 4 == 0: the quad is in range as a whole
        for (int i = lane >> 3; i < rows; i += 4) {
          const int64_t m = mw0 + i;
          float4 w = *reinterpret_cast<const float4 *>(&buf[i * BS + cq]);
          if (e.mask_pos) {
            const float4 k4 = __ldg(reinterpret_cast<const float4 *>(e.mask_pos + m * e.ld_mask + n));
            w.x = k4.x > 0.f ? w.x * e.mask_scale : 0.f; w.y = k4.y > 0.f ? w.y * e.mask_scale : 0.f;
            w.z = k4.z > 0.f ? w.z * e.mask_scale : 0.f; w.w = k4.w > 0.f ? w.w * e.mask_scale : 0.f;
          }
          if (e.residual) {
            const float4 r4 = __ldg(reinterpret_cast<const float4 *>(e.residual + m * e.ld_res + n));
            w.x += r4.x; w.y += r4.y; w.z += r4.z; w.w += r4.w;
          }
          float *c = g.C + m * g.ldc + n;
          if (e.atomic) { atomicAdd(c, w.x); atomicAdd(c + 1, w.y); atomicAdd(c + 2, w.z); atomicAdd(c + 3, w.w); }
          else if (e.accumulate) {
            const float4 o = *reinterpret_cast<const float4 *>(c);
            *reinterpret_cast<float4 *>(c) = make_float4(o.x + w.x, o.y + w.y, o.z + w.z, o.w + w.w);
          } else *reinterpret_cast<float4 *>(c) = w;
        }
      }
      continue;
    }
    const int64_t n = nb + lane;
    if (n < g.N) {
      for (int i = 0; i < rows; ++i) {
        const int64_t m = mw0 + i;
        float w = buf[i * BS + lane];
        if (e.pe) {
          w += __ldg(e.pe + (m % T) * g.N + n);
          if (edrop.thr) w = drop_keep(edrop.key, edrop.thr, (uint64_t)((e.drop_row0 + m) * g.N + n)) ? w * edrop.scale : 0.f;
        }
        if (e.mask_pos) w = __ldg(e.mask_pos + m * e.ld_mask + n) > 0.f ? w * e.mask_scale : 0.f;
        if (e.residual) w += __ldg(e.residual + m * e.ld_res + n);
        float *c = g.C + m * g.ldc + n;
        if (e.atomic) atomicAdd(c, w);
        else if (e.accumulate) *c += w;
        else *c = w;
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TCOLS);
}

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// scratch for the operand images: CALLER-OWNED (include/groove_b200.h: the library keeps no device memory).  The pass drivers
// carve it out of the workspace (runner.cu: Plan::gemm_img, sized by gemm_tc_scratch_bytes) and bind it to the calling thread
// for the duration of the pass; GEMMs of one pass run in stream order, so the next pre-pass cannot overwrite images a previous
// main kernel still reads, and two models never share a buffer (each has its own workspace).  Without a bound scratch — or when
// a problem's images do not fit — the GEMM stages its operands itself.
constexpr size_t IMG_MAX_BYTES = (size_t)8 << 30;
constexpr int GEMM_TC_NO_SPLIT = -77;      // launch(): a split-mode problem whose operands cannot both be pre-imaged (gemm_tc falls back to FFMA)
thread_local uint8_t *g_img_scratch = nullptr;
thread_local size_t g_img_scratch_bytes = 0;
uint8_t *img_scratch(size_t need) { return (g_img_scratch != nullptr && need <= g_img_scratch_bytes) ? g_img_scratch : nullptr; }

// 0: operands staged by the GEMM's own threads; 1: weight operand pre-imaged (GEMMs without split-K); 2 (default): both operands
// pre-imaged whenever the scratch fits, weight-only otherwise
int pre_mode() {
  static const int m = getenv("GT_GEMM_PRE") ? atoi(getenv("GT_GEMM_PRE")) : 2;
  return m;
}

template <int BN>
int launch(GemmTcArgs g, int64_t m_tiles, int64_t splits, cudaStream_t st) {
  constexpr size_t smem1 = (size_t)GSTG * (GBM * GBK * 2 + BN * GBK * 2);
  // split mode: a stage holds three images per operand (BN <= 128: 192 KB, one CTA per SM)
  const size_t smem = g.split ? 3 * smem1 : smem1;
  if (g.split && BN > 128) return GEMM_TC_NO_SPLIT;
  static bool attr_done = false;
  if (!attr_done) {
    GT_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(BN <= 128 ? 3 * smem1 : smem1)));
    attr_done = true;
  }
  g.b_img = nullptr; g.a_img = nullptr; g.nkb_img = 0;
  const int64_t nkb = (g.K + GBK - 1) / GBK;
  const size_t im = g.split ? 3 : 1;
  const size_t need_b = (size_t)g.n_tiles * nkb * BN * GBK * 2 * im, need_a = (size_t)m_tiles * nkb * GBM * GBK * 2 * im;
  const int mode = pre_mode();                                    // the scratch is a fixed workspace region: valid under stream capture too
  const bool big = m_tiles * g.n_tiles * splits >= 32;            // small problems are launch-latency bound: no extra kernels
  const bool all = mode >= 2 && (big || g.split) && nkb <= 65535 && need_a + need_b <= IMG_MAX_BYTES && need_a + need_b <= g_img_scratch_bytes;
  const bool only_b = !all && mode >= 1 && splits == 1 && m_tiles >= 16 && need_b <= ((size_t)4 << 20);
  if (g.split && !(all && img_scratch(need_a + need_b) != nullptr)) return GEMM_TC_NO_SPLIT;     // small problem / no scratch
  if (all || only_b) {
    uint8_t *img = img_scratch(all ? need_a + need_b : need_b);
    if (img != nullptr) {
      { LaunchScope _ls(KC_GEMM_TC, st);
        gemm_tc_bimg_kernel<BN><<<dim3((unsigned)g.n_tiles, (unsigned)nkb), GTHREADS, 0, st>>>(g, img); }
      GT_CUDA(cudaGetLastError());
      g.b_img = img; g.nkb_img = (int)nkb;
      if (all) {
        { LaunchScope _ls(KC_GEMM_TC, st);
          gemm_tc_aimg_kernel<<<dim3((unsigned)m_tiles, (unsigned)nkb), GTHREADS, 0, st>>>(g, img + need_b); }
        GT_CUDA(cudaGetLastError());
        g.a_img = img + need_b;
      }
    }
  }
  dim3 grid((unsigned)(m_tiles * g.n_tiles), (unsigned)splits);
  { LaunchScope _ls(KC_GEMM_TC, st);
    gemm_tc_kernel<BN><<<grid, GTHREADS, smem, st>>>(g); }
  GT_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

void gemm_tc_bind_scratch(void *p, int64_t bytes) {
  g_img_scratch = static_cast<uint8_t *>(p);
  g_img_scratch_bytes = p != nullptr && bytes > 0 ? (size_t)bytes : 0;
}

// upper bound of the operand images of any Linear forward / dgrad / wgrad GEMM of a model with these widths over `tokens` rows:
// forward / dgrad: A = [tokens x K], B = [N x K]; wgrad: A = [N_w x tokens], B = [K_w x tokens] (both <= the widest layer dims)
int64_t gemm_tc_scratch_bytes(int64_t tokens, int64_t d, int64_t F, int split) {
  auto pad = [](int64_t v, int64_t m) { return (v + m - 1) / m * m; };
  const int64_t wide = pad(std::max<int64_t>(3 * d, F), 256), narrow = pad(std::max<int64_t>(d, F), 256);
  const int64_t need = (split ? 6 : 2) * pad(tokens, 128) * (wide + narrow) + ((int64_t)1 << 20);
  return std::min<int64_t>(need, (int64_t)IMG_MAX_BYTES);
}

bool gemm_tc_supported(int64_t sam, int64_t sak, int64_t sbn, int64_t sbk, int64_t M, int64_t N, int64_t K) {
  // one of the two indices of each operand must be contiguous; tiny contractions / outputs stay on the SIMT kernel
  if (!((sak == 1) || (sam == 1)) || !((sbk == 1) || (sbn == 1))) return false;
  return M >= 32 && N >= 32 && K >= 32;
}

int gemm_tc(const float *A, int64_t sam, int64_t sak, const float *B, int64_t sbn, int64_t sbk, float *C, int64_t ldc, int64_t M,
            int64_t N, int64_t K, const GemmEpi &epi, int64_t split_k_chunk, cudaStream_t st, int split) {
  if (M == 0 || N == 0) return 0;
  const int64_t split_k_chunk_in = split_k_chunk;
  GT_CHECK(gemm_tc_supported(sam, sak, sbn, sbk, M, N, K), "gemm_tc: unsupported operand layout / shape");
  GemmTcArgs g;
  g.A = A; g.B = B; g.C = C; g.sam = sam; g.sak = sak; g.sbn = sbn; g.sbk = sbk; g.ldc = ldc;
  g.M = M; g.N = N; g.K = K; g.e = epi; g.split = split ? 1 : 0;
  // tile width: the narrowest of {32, 64, 128, 256} that covers N, 256-wide tiles beyond that (a 384 / 768-wide output uses 128 / 256)
  // (measured on C3, batch 8192: capping the tile at 128 / 64 columns costs 23 % / 69 % of the GEMM time — every extra n-tile
  //  re-stages and re-converts the 128-row A block, which is what bounds this kernel)
  int bn;
  if (N <= 32) bn = 32;
  else if (N <= 64) bn = 64;
  else if (N <= 128) bn = 128;
  else if (N <= 256) bn = 256;
  else bn = (N % 256 == 0 || N % 256 > 128) ? 256 : ((N % 128 == 0 || N % 128 > 64) ? 128 : 256);
  if (split && bn > 128) bn = 128;                        // a split-mode stage (three images per operand) fits for BN <= 128
  g.n_tiles = (int)((N + bn - 1) / bn);
  const int64_t m_tiles = (M + GBM - 1) / GBM;
  int64_t splits = 1;
  if (split_k_chunk > 0 && K > split_k_chunk) {
    split_k_chunk = (split_k_chunk + GBK - 1) / GBK * GBK;
    splits = (K + split_k_chunk - 1) / split_k_chunk;
    // Weight gradients have a handful of output tiles and a very long contraction: the requested chunk only bounds the split from
    // above.  Shrink it so that tiles x splits fills whole waves of the 2 x 148 resident CTAs (C3 dW1: 4 tiles x 128 splits =
    // 1.73 waves of 2048-token chunks -> 4 x 147 = 1.99 waves of 1792-token chunks).
    static int slots = 0;
    if (slots == 0) {
      int dev = 0, sms = 148;
      if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      slots = 2 * sms;
    }
    const int64_t tiles = m_tiles * g.n_tiles;
    if (splits >= 8 && tiles < slots) {
      const int64_t waves = (tiles * splits + slots - 1) / slots;
      const int64_t want = waves * slots / tiles;
      if (want > splits) {
        const int64_t chunk = ((K + want - 1) / want + GBK - 1) / GBK * GBK;
        if (chunk >= 4 * GBK) { split_k_chunk = chunk; splits = (K + chunk - 1) / chunk; }
      }
    }
    while (splits > 65535) { split_k_chunk *= 2; splits = (K + split_k_chunk - 1) / split_k_chunk; }
    g.kchunk = split_k_chunk;
    GT_CHECK(epi.atomic, "split-K GEMM needs an atomic epilogue");
  } else {
    g.kchunk = (K + GBK - 1) / GBK * GBK;
  }
  const int64_t a_ld = sak != 1 ? sak : sam, b_ld = sbk != 1 ? sbk : sbn;
  g.a_vec = aligned16(A) && (a_ld % 4 == 0);
  g.b_vec = aligned16(B) && (b_ld % 4 == 0);
  // a K-major source starts its chunks at k0 (multiple of 64) and an MN-major one at m0 / n0 (multiples of 128 / BN): always 16-byte
  // aligned relative to the base; K-split offsets are multiples of 64 as well.
  GT_CHECK(m_tiles * ((N + 31) / 32) < (int64_t)1 << 31, "gemm_tc: too many tiles");
  int rc;
  switch (bn) {
    case 32: rc = launch<32>(g, m_tiles, splits, st); break;
    case 64: rc = launch<64>(g, m_tiles, splits, st); break;
    case 128: rc = launch<128>(g, m_tiles, splits, st); break;
    default: rc = launch<256>(g, m_tiles, splits, st); break;
  }
  // split mode promises fp32-class results: a problem that cannot be pre-imaged (a handful of tiles, no scratch bound) runs on the
  // exact FFMA kernel instead of the in-kernel bf16 staging
  if (rc == GEMM_TC_NO_SPLIT) return gemm_f32(A, sam, sak, B, sbn, sbk, C, ldc, M, N, K, epi, split_k_chunk_in, st);
  return rc;
}

}  // namespace gt
