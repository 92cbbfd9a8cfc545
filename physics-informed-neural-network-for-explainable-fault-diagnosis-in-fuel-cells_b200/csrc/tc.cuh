// tc.cuh -- tcgen05 / TMEM / mbarrier building blocks (inline PTX, sm_100a).
//
// The hidden-layer contractions of the DNN are [128 samples] x [64 in] x [64 out] GEMMs.
// They run on the 5th-gen tensor cores as `tcgen05.mma.kind::tf32` with fp32 accumulation in
// TMEM.  Plain TF32 (10-bit mantissa) would break the 1e-5 parity bar, so both operands are
// split x = hi + lo (hi = tf32-rounded x, lo = x - hi, exact) and three MMAs accumulate
// lo*hi + hi*lo + hi*hi into the same TMEM tile ("3xTF32"): the dropped lo*lo term is
// 2^-22 relative.  Operands live in shared memory in the canonical K-major, no-swizzle
// ("interleave") UMMA layout:
//
//     byte_offset(row, k) = (k/4) * LBO + (row/8) * SBO + (row%8) * 16 + (k%4) * 4
//
// with SBO = 128 B (8-row core matrices back to back) and LBO = rows * 16 B, so the thread
// that owns sample `row` writes its 16-byte chunks at  kc*LBO + row*16  -- consecutive
// lanes hit consecutive 16-byte slots: conflict-free 128-bit stores.
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"

namespace pinn {
namespace tc {

PINN_D uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// One lane of a converged warp (CUTLASS's elect_one_sync): keeps the surrounding code
// warp-uniform, so UMMA descriptors stay in uniform registers.
PINN_D bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, %1;\n\t"
      "@px mov.s32 %0, 1;\n\t}"
      : "+r"(pred)
      : "r"(0xFFFFFFFFu));
  return pred != 0;
}
// Warp index as a provably warp-uniform value (shuffle broadcast).
PINN_D int uniform_warp_idx() { return __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0); }

// ------------------------------------------------------------------------- mbarrier
PINN_D void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
PINN_D void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
PINN_D void mbar_wait(uint64_t* bar, uint32_t parity) {
  // try_wait with a suspend-time hint: the hardware parks the thread until the phase completes
  // (event-driven wake-up) or the hint expires, so waiting warps neither poll nor oversleep.
  const uint32_t addr = smem_u32(bar);
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity), "r"(20000u)
        : "memory");
  } while (!ok);
}
// non-blocking probe of a phase (the MMA warp polls two barriers)
PINN_D bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
PINN_D void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// TMA 1-D bulk copy global -> shared; completion is signalled on `bar` as transaction bytes.
PINN_D void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
PINN_D void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// TMA tiled copy global -> shared through a tensor map: box at (c0 = innermost coordinate, c1) of a 2-D tensor
PINN_D void tma_load_2d(void* dst_smem, const void* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_u32(dst_smem)), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}
// generic-proxy smem writes -> visible to the async proxy (UMMA operand reads)
PINN_D void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ----------------------------------------------------------------------------- TMEM
PINN_D void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
PINN_D void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// whole warp; ncols power of two >= 32; the TMEM base address lands in *dst (shared memory)
PINN_D void tmem_alloc(uint32_t* dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "r"(ncols)
               : "memory");
}
PINN_D void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
PINN_D void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// 32 lanes x 32 bit, 16 consecutive columns: thread i of the warp reads TMEM lane (lane_base + i)
PINN_D void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
PINN_D void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// 32 lanes x 32 bit, 16 consecutive columns: thread i of the warp writes TMEM lane (lane_base + i)
PINN_D void tmem_st16(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}
PINN_D void tmem_st8(uint32_t taddr, const float* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                 "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                 "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}
PINN_D void tmem_st4(uint32_t taddr, const float* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
               ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3]))
               : "memory");
}
PINN_D void tmem_st1(uint32_t taddr, float v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(__float_as_uint(v)) : "memory");
}
PINN_D void tmem_ld4(uint32_t taddr, float* v) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
PINN_D void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
PINN_D void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------ UMMA
// K-major, no swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1).
PINN_D uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;  // descriptor version (Blackwell)
  return d;         // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE / interleave)
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32; a_mn / b_mn
// select MN-major (bit 15 / 16) instead of K-major operands.
PINN_HD constexpr uint32_t make_idesc_tf32(int M, int N, bool a_mn = false, bool b_mn = false) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u) |
         (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
// The same plane  off(row, col) = (col/4)*LBO + row*16 + (col%4)*4  read as an MN-major operand
// (MN index = col, K index = row): canonical layout ((1,n),(8,k)):((X,SBO'),(1,LBO')) in 16-byte
// units with SBO' = LBO (distance between 4-column chunks) and LBO' = 128 B (8 rows).
PINN_D uint64_t make_desc_mn(uint32_t smem_addr, uint32_t plane_lbo_bytes) { return make_desc(smem_addr, 128, plane_lbo_bytes); }
// D[tmem] (+)= A[smem] * B[smem]^T, one K = 8 slab.  Issued by ONE thread.
PINN_D void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same with the A operand in TENSOR MEMORY (lane = row, one 32-bit column per k): the tensor core
// then fetches only B from shared memory.  With both operands in shared memory an M128 N64 K8 tf32
// MMA pulls 6 KB through the 128 B/clk shared-memory port -- 48 clk against a 32 clk math floor --
// and those reads compete with the epilogue's own LDS/STS (profiles/README.md, timeline_mc).
PINN_D void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on `bar` once every MMA issued so far by this thread has completed (implies
// tcgen05.fence::before_thread_sync).
PINN_D void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ------------------------------------------------------------------- operand staging
PINN_D float tf32_hi(float x) {
  uint32_t h;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
  return __uint_as_float(h);
}
// Write one 16-byte K-chunk (4 consecutive k) of row `row` into the hi and lo planes.
PINN_D void store_split4(float* hi_plane, float* lo_plane, uint32_t lbo_bytes, int row, int kc, float4 v) {
  float4 h = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
  float4 l = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
  const size_t off = (static_cast<size_t>(kc) * lbo_bytes + static_cast<size_t>(row) * 16) / sizeof(float);
  *reinterpret_cast<float4*>(hi_plane + off) = h;
  *reinterpret_cast<float4*>(lo_plane + off) = l;
}

// Activations are finite and far from overflow, so their tf32 rounding needs no NaN/Inf
// guard: add half an ulp (bit 12) to the magnitude and clear the low 13 bits -- two integer
// ops instead of the four cvt.rna.tf32 expands to; identical result (ties away from zero).
#ifdef PINN_TF32_TRUNC
PINN_D float tf32_hi_fast(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
#else
PINN_D float tf32_hi_fast(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }
#endif
// Eight consecutive columns c0..c0+7 (c0 % 8 == 0) of row `row` into the hi and lo planes.
PINN_D void store_split8_fast(float* hi_plane, float* lo_plane, uint32_t lbo_bytes, int row, int c0, const float (&v)[8]) {
  float h[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) h[q] = tf32_hi_fast(v[q]);
  const size_t off = (static_cast<size_t>(c0 >> 2) * lbo_bytes + static_cast<size_t>(row) * 16) / sizeof(float);
  const size_t off2 = off + lbo_bytes / sizeof(float);
  *reinterpret_cast<float4*>(hi_plane + off) = make_float4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<float4*>(lo_plane + off) = make_float4(v[0] - h[0], v[1] - h[1], v[2] - h[2], v[3] - h[3]);
  *reinterpret_cast<float4*>(hi_plane + off2) = make_float4(h[4], h[5], h[6], h[7]);
  *reinterpret_cast<float4*>(lo_plane + off2) = make_float4(v[4] - h[4], v[5] - h[5], v[6] - h[6], v[7] - h[7]);
}

// Issue the 3xTF32 product  D[M x N] = A[M x 64] * B[N x 64]^T  from one thread: 24 MMAs
// back to back.  Descriptors are built once by the caller; stepping one K = 8 slab adds
// 2*LBO (in 16-byte units) to the start-address field, so each MMA costs one add.
template <int K>
PINN_D void issue_3xtf32(uint32_t d_tmem, uint64_t a_hi0, uint64_t a_lo0, uint32_t lbo_a, uint64_t b_hi0, uint64_t b_lo0,
                         uint32_t lbo_b, uint32_t idesc) {
  const uint64_t a_step = (2u * lbo_a) >> 4, b_step = (2u * lbo_b) >> 4;
#if defined(PINN_ABL) && (PINN_ABL & 4)     // ablation build: one MMA instead of 24
  umma_tf32(d_tmem, a_hi0, b_hi0, idesc, 0u);
  return;
#endif
#pragma unroll
  for (int term = 0; term < 3; ++term) {      // small terms first: lo*hi, hi*lo, then hi*hi
    const uint64_t a = term == 0 ? a_lo0 : a_hi0;
    const uint64_t b = term == 1 ? b_lo0 : b_hi0;
#pragma unroll
    for (int k8 = 0; k8 < K / 8; ++k8)        // one MMA consumes K = 8 tf32 = two 16-byte chunks
      umma_tf32(d_tmem, a + k8 * a_step, b + k8 * b_step, idesc, (term | k8) != 0 ? 1u : 0u);
  }
}

// 3xTF32 product with A (hi and lo planes, K columns each) in tensor memory.
template <int K>
PINN_D void issue_3xtf32_ts(uint32_t d_tmem, uint32_t a_hi_t, uint32_t a_lo_t, uint64_t b_hi0, uint64_t b_lo0, uint32_t lbo_b,
                            uint32_t idesc) {
  const uint64_t b_step = (2u * lbo_b) >> 4;
#pragma unroll
  for (int term = 0; term < 3; ++term) {      // small terms first: lo*hi, hi*lo, then hi*hi
    const uint32_t a = term == 0 ? a_lo_t : a_hi_t;
    const uint64_t b = term == 1 ? b_lo0 : b_hi0;
#pragma unroll
    for (int k8 = 0; k8 < K / 8; ++k8)
      umma_tf32_ts(d_tmem, a + 8u * k8, b + k8 * b_step, idesc, (term | k8) != 0 ? 1u : 0u);
  }
}

// The same product issued slab by slab: the three terms of the two K = 8 slabs `s0`, `s1` (A columns 8 s .. 8 s + 7).
// `first`: this call starts the accumulation (its very first MMA overwrites D).
PINN_D void issue_3xtf32_ts_slabs(uint32_t d_tmem, uint32_t a_hi_t, uint32_t a_lo_t, uint64_t b_hi0, uint64_t b_lo0, uint32_t lbo_b,
                                  uint32_t idesc, int s0, int s1, bool first) {
  const uint64_t b_step = (2u * lbo_b) >> 4;
#pragma unroll
  for (int term = 0; term < 3; ++term) {      // small terms first: lo*hi, hi*lo, then hi*hi
    const uint32_t a = term == 0 ? a_lo_t : a_hi_t;
    const uint64_t b = term == 1 ? b_lo0 : b_hi0;
    umma_tf32_ts(d_tmem, a + 8u * static_cast<uint32_t>(s0), b + static_cast<uint64_t>(s0) * b_step, idesc, (first && term == 0) ? 0u : 1u);
    umma_tf32_ts(d_tmem, a + 8u * static_cast<uint32_t>(s1), b + static_cast<uint64_t>(s1) * b_step, idesc, 1u);
  }
}

// ------------------------------------------------------------------------------ fp16-pair split (kind::f16 products)
// a = a_h + a_l with a_h = fp16(a), a_l = fp16(a - a_h): 22 significant bits for O(1) values; the three products
// a_l*w_h + a_h*w_l + a_h*w_h run as kind::f16 MMAs (twice the tf32 rate, half the operand bytes), fp32 accumulation.
// two fp32 -> packed fp16 pair (hi) and the packed fp16 pair of the remainders (lo)
PINN_D void split_h2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const float2 d = __ffma2_rn(hf, make_float2(-1.0f, -1.0f), make_float2(a, b));      // a - a_h, exact; one packed instruction
  const __half2 l = __floats2half2_rn(d.x, d.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

PINN_HD constexpr uint32_t make_idesc_f16(int M, int N) {      // D = F32, A = B = F16, both K-major
  return (1u << 4) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
// D[tmem] (+)= A[tmem, packed fp16 pairs] * B[smem]^T, one K = 16 slab.  Issued by ONE thread.
PINN_D void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

}  // namespace tc
}  // namespace pinn
