// umma.cuh — thin inline-PTX layer over the Blackwell tensor-core programming model:
// tcgen05.mma (UMMA) with shared-memory descriptors, TMEM alloc / ld, mbarrier, bulk-TMA copies.
// sm_100a only.  Bit layouts follow the PTX ISA "tcgen05 matrix / instruction descriptor" tables.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace gt {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- canonical K-major, no-swizzle operand layout ---------------------------------------------
// An operand tile is R rows (M or N) x K columns of bf16.  It is stored as 8x8 "core matrices"
// (8 rows x 16 bytes, 128 contiguous bytes each).  Core matrix (rb, kb) = rows 8rb.., k 8kb.. lives
// at byte offset (kb * (R/8) + rb) * 128, so
//    stride between K-adjacent core matrices   ("leading byte offset")  = (R/8) * 128
//    stride between row-adjacent core matrices ("stride  byte offset")  = 128
// Thread-per-row producers write one 16-byte chunk per (row, kb): consecutive rows are consecutive
// 16-byte words -> conflict-free st.shared.v4.
__host__ __device__ __forceinline__ uint32_t kmajor_off(int r, int k, int R) {
  return (uint32_t)(((k >> 3) * (R >> 3) + (r >> 3)) * 128 + (r & 7) * 16 + (k & 7) * 2);
}

// 64-bit shared-memory matrix descriptor (SWIZZLE_NONE, version 1 = Blackwell)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// 32-bit instruction descriptor: D=f32, A=B=bf16, both K-major, dense
__host__ __device__ __forceinline__ uint32_t make_idesc_bf16(int M, int N, int a_mn_major = 0, int b_mn_major = 0) {
  uint32_t d = 0;
  d |= 1u << 4;                       // c_format = F32
  d |= 1u << 7;                       // a_format = BF16
  d |= 1u << 10;                      // b_format = BF16
  d |= (uint32_t)(a_mn_major & 1) << 15;
  d |= (uint32_t)(b_mn_major & 1) << 16;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread
__device__ __forceinline__ void mma_bf16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// make generic-proxy st.shared visible to the async proxy (UMMA / TMA reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM ------------------------------------------------------------------------------------
// one full warp; ncols power of two in [32,512]; the base address is written to *slot (smem)
__device__ __forceinline__ void tmem_alloc(uint32_t *slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// warp-wide: lane l receives 32 consecutive fp32 columns of TMEM lane (taddr.lane + l)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v) {
  uint32_t *r = reinterpret_cast<uint32_t *>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v) {
  uint32_t *r = reinterpret_cast<uint32_t *>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float *v) {
  uint32_t *r = reinterpret_cast<uint32_t *>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// registers -> TMEM: lane (32 w % 4 + l) of the warp's lane quadrant, 16 consecutive 32-bit columns (the mirror of tmem_ld16).
// A later tcgen05.mma that ACCUMULATES onto these columns sees them after tmem_st_wait + fence_before_sync + a thread sync.
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float *v) {
  const uint32_t *r = reinterpret_cast<const uint32_t *>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t *r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               :
               : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- mbarrier -----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a lost arrival becomes a trap (reported as a CUDA error) instead of a hung GPU
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
#pragma unroll 1
  for (uint32_t spin = 0; spin < (1u << 26); ++spin)
    if (mbar_try_wait(bar, parity)) return;
  __trap();
}

// warp-converged election of one issuing lane (the same lane every time for a full warp)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// named barrier over a subset of the CTA's warps (id 1..15; nthreads multiple of 32)
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// ---- bulk TMA (1-D): shared -> global ------------------------------------------------------------
__device__ __forceinline__ void tma_store_1d(void *gmem_dst, const void *smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- warp-level mma.sync (attention inside the fused d_model = 256 kernels) ---------------------------
__device__ __forceinline__ uint32_t lds32(const uint8_t *p) { return *reinterpret_cast<const uint32_t *>(p); }
// D(16x8, f32) += A(16x16 bf16, row) * B(16x8 bf16, col)
__device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// four transposed 8x8 b16 matrices; lane l supplies the address of row (l & 7) of matrix (l >> 3)
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void *smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(smem_row)));
}
// four 8x8 b16 matrices as they lie (row = 16 contiguous bytes); lane l supplies the address of row (l & 7) of matrix (l >> 3)
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void *smem_row) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(smem_row)));
}
// transpose one 8x8 b16 matrix held as a mma fragment (lane holds row lane/4, columns 2*(lane%4), +1)
__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t a) {
  uint32_t d;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
  return d;
}

// D(16x8, f32) += A(16x8 tf32, row) * B(8x8 tf32, col)
__device__ __forceinline__ void mma1688_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t f32_to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

// ---- column sums over the 32 lanes of a warp on the tensor cores (lane = token row, registers = feature columns) ----------
// The butterfly (31 shuffles + 62 selects + 31 adds per 32 values) was 15 % of the instructions of the d_model = 256 backward.
// Here the lane's 16 values go in as the B operand of four mma.m16n8k16: B[k][n] of lane (g = lane / 4, t = lane % 4) is
// "row-group g, slot (t, register, half)", and a 0/1 SELECTOR A operand routes slot (register r, half h) of group q's mma to
// output row 4 q + 2 r + h — so D1[column c][n] = sum of column c over the four lanes of row-group n.  One tf32 mma with an
// all-ones B then sums D1 over its 8 row-groups (the D1 accumulator fragment IS the A fragment layout of m16n8k8, up to the
// order of k, which a sum does not see).  Result: every lane (g, t) holds column g in s0 and column g + 8 in s1.
// Inputs are rounded to bf16 (like every other tensor-core column sum of this library), the partial sums to tf32.
struct ColsumSel {
  uint32_t l0, h0, l1, h1;
};
__device__ __forceinline__ ColsumSel colsum_sel(int lane) {
  const int g = lane >> 2, gi = g & 3;
  const uint32_t lo = gi == 0 ? 0x00003F80u : (gi == 1 ? 0x3F800000u : 0u);     // k < 8: even k (low half) / odd k (high half)
  const uint32_t hi = gi == 2 ? 0x00003F80u : (gi == 3 ? 0x3F800000u : 0u);     // k >= 8
  ColsumSel s;
  s.l0 = g < 4 ? lo : 0u; s.h0 = g < 4 ? hi : 0u; s.l1 = g < 4 ? 0u : lo; s.h1 = g < 4 ? 0u : hi;
  return s;
}
// pk[j] = packed bf16 pair (column 2 j, column 2 j + 1) of this lane's row, j = 0..7
__device__ __forceinline__ void warp_colsum16_packed(const uint32_t (&pk)[8], const ColsumSel &s, float &s0, float &s1) {
  float d[4] = {0.f, 0.f, 0.f, 0.f};
  mma16816(d, s.l0, 0u, s.h0, 0u, pk[0], pk[1]);      // columns 0..3   -> rows 0..3
  mma16816(d, s.l1, 0u, s.h1, 0u, pk[2], pk[3]);      // columns 4..7   -> rows 4..7
  mma16816(d, 0u, s.l0, 0u, s.h0, pk[4], pk[5]);      // columns 8..11  -> rows 8..11
  mma16816(d, 0u, s.l1, 0u, s.h1, pk[6], pk[7]);      // columns 12..15 -> rows 12..15
  float e[4] = {0.f, 0.f, 0.f, 0.f};
  mma1688_tf32(e, f32_to_tf32(d[0]), f32_to_tf32(d[2]), f32_to_tf32(d[1]), f32_to_tf32(d[3]), 0x3F800000u, 0x3F800000u);
  s0 = e[0]; s1 = e[2];
}

// 8 columns only (pk[j] = columns 2 j, 2 j + 1): every lane (g, t) receives the sum of column g
__device__ __forceinline__ float warp_colsum8_packed(const uint32_t (&pk)[4], const ColsumSel &s) {
  float d[4] = {0.f, 0.f, 0.f, 0.f};
  mma16816(d, s.l0, 0u, s.h0, 0u, pk[0], pk[1]);      // columns 0..3 -> rows 0..3
  mma16816(d, s.l1, 0u, s.h1, 0u, pk[2], pk[3]);      // columns 4..7 -> rows 4..7
  float e[4] = {0.f, 0.f, 0.f, 0.f};
  mma1688_tf32(e, f32_to_tf32(d[0]), 0u, f32_to_tf32(d[1]), 0u, 0x3F800000u, 0x3F800000u);
  return e[0];
}

// ---- bulk TMA (1-D): global -> shared, completion counted in bytes on an mbarrier -------------
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// the same copy with an L2 eviction-priority hint (createpolicy): the weight stage streams are re-read by every CTA for every tile
// (evict_last) while ~3 MB of activations per tile stream through L2 once (evict_first)
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void tma_load_1d_hint(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
               : "memory");
}

// asynchronous L2 prefetch of a contiguous global range (bytes: multiple of 16)
__device__ __forceinline__ void prefetch_l2_bulk(const void *gmem, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem), "r"(bytes) : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t *>(&v);
}

// two fp32 -> packed bf16 pair with ReLU fused into the conversion (low half = lo)
__device__ __forceinline__ uint32_t pack_bf16_relu(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
// Two dropout decisions at once.  `fields` holds two 14-bit fields in its 16-bit halves (common.cuh: hash_quad), `thr2` the
// 14-bit threshold in both halves: as fp16 bit patterns both are non-negative and finite, where fp16 order = integer order,
// so ONE packed half-precision compare yields 0xFFFF (keep) / 0x0000 (drop) per half — AND it onto a packed bf16 pair.
__device__ __forceinline__ uint32_t keep2(uint32_t fields, uint32_t thr2) {
  return __hge2_mask(*reinterpret_cast<const __half2 *>(&fields), *reinterpret_cast<const __half2 *>(&thr2));
}
// 0xFFFF per half where the packed bf16 (or fp16) value is non-zero (inputs are >= +0: ReLU outputs)
__device__ __forceinline__ uint32_t nonzero2(uint32_t v) {
  const uint32_t z = 0u;
  return __hne2_mask(*reinterpret_cast<const __half2 *>(&v), *reinterpret_cast<const __half2 *>(&z));
}
__device__ __forceinline__ float rcp_approx(float x) {      // 1 / x, one MUFU (x is a softmax denominator in [1, 32])
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

}  // namespace umma
}  // namespace gt
