// Blackwell (sm_100a) tensor-core plumbing used by the tcgen05 kernels: inline-PTX wrappers for
// tcgen05.mma (kind::tf32, operands in shared memory, accumulator in TMEM), TMEM allocation and
// loads, mbarriers, bulk async copies and proxy fences, plus the shared-memory operand layout and
// its matrix descriptors.
//
// Operand layout (no swizzle, "interleaved" canonical form of the UMMA descriptors): the unit is a
// CORE MATRIX of 8 rows x 16 bytes (4 fp32/tf32 values), 128 contiguous bytes, row r at byte 16*r.
// A row-major source X[rows][cols] (cols contiguous) is stored as core matrices
//     core(rg, cc) = X[8*rg .. 8*rg+7][4*cc .. 4*cc+3]
// and used as a K-major operand (rows = M or N index, cols = K index): SBO = stride between row
// groups, LBO = stride between 16-byte column chunks.
//
// MN-major operands (cols = M or N index, rows = K index; the dW = X^T.dZ GEMMs, whose reduction runs
// over batch rows): for 32-bit elements the tensor core only accepts the SWIZZLE_128B_BASE32B layout
// (every other layout type returned zeros on B200 -- measured with tools/tc_probe.py): a K-row holds
// 32 consecutive MN elements (128 bytes) whose four 32-byte chunks are XOR-ed with (row % 4); four rows
// make a 512-byte group; groups along K are SBO apart, 32-element blocks along MN are LBO apart.
//
// fp32 accuracy on TF32 tensor cores (3xTF32): x = hi + lo with hi = tf32(x) (round to nearest) and
// lo = tf32(x - hi); a.b ~= hi_a.hi_b + lo_a.hi_b + hi_a.lo_b accumulated in fp32 -- the dropped
// lo.lo term and the rounding of lo are ~2^-22 relative, inside the 1e-5 parity budget.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace mr {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- 3xTF32 split ------------------------------------------------------------------------------
__device__ __forceinline__ float tf32_rn(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
// Exact split used where it is computed once (weights): both halves rounded to TF32.
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = tf32_rn(x);
  lo = tf32_rn(x - hi);
}
// Cheap split for the per-element activation path (2 instructions instead of ~14: the producers were
// issue-bound on cvt.rna, which expands to a 6-instruction sequence): hi = x with the low 13 mantissa
// bits cleared (a valid TF32), lo = x - hi (exact in fp32, |lo| < 2^-10 |x|); the tensor core ignores the
// low 13 bits of lo, an error of at most 2^-20 |x|, the same order as the dropped lo.lo term.
__device__ __forceinline__ void split_tf32_fast(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
  lo = x - hi;
}
__device__ __forceinline__ void split_tf32x4(const float4& x, float4& hi, float4& lo) {
  split_tf32_fast(x.x, hi.x, lo.x);
  split_tf32_fast(x.y, hi.y, lo.y);
  split_tf32_fast(x.z, hi.z, lo.z);
  split_tf32_fast(x.w, hi.w, lo.w);
}

// ---- descriptors -------------------------------------------------------------------------------
// Shared-memory matrix descriptor, SWIZZLE_NONE, sm_100 version field = 1 (bits 46-47).
constexpr uint32_t kLayoutNone = 0;          // K-major operands
constexpr uint32_t kLayoutSw128Base32 = 1;   // MN-major 32-bit operands
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout_type = kLayoutNone) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | ((uint64_t)layout_type << 61);
}
// Instruction descriptor for kind::tf32, fp32 accumulate: c_format=F32 (bits 4-5 = 1), a/b format =
// TF32 (2) at bits 7-9 / 10-12, a_major bit 15, b_major bit 16 (0 = K-major, 1 = MN-major),
// N>>3 at bits 17-22, M>>4 at bits 24-28.
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] . B[smem]; issued by ONE thread on behalf of the CTA.
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// One lane of a converged warp, chosen by the hardware.  Code under `if (elect_one())` is compiled for the
// uniform datapath: descriptors live in uniform registers and consecutive tcgen05.mma are 1-2 instructions
// apart.  Under `if (lane == 0)` the compiler instead wraps EVERY mma in a R2UR + ELECT + BRA.U.ANY waterfall
// (~14 instructions), and the single issuing thread cannot keep the tensor pipe fed.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
// Arrive on an mbarrier when every tcgen05 op issued so far by this thread has completed.
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// Make generic-proxy shared-memory writes visible to the async proxy (tensor core / bulk copy).
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM --------------------------------------------------------------------------------------
// One full warp allocates `ncols` (power of two >= 32) columns; the base address lands in *slot.
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// Warp w reads TMEM lanes 32*(w%4)..+31: thread t gets 16 consecutive columns of lane 32*(w%4)+t.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Thread t of warp w writes 16 consecutive columns of TMEM lane 32*(w%4)+t (the A operand of a
// tcgen05.mma whose A comes from tensor memory: lane = row of A, column = K index).
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] . B[smem]: A read from tensor memory (no shared-memory operand fetch for A: the SS form of
// a 128 x 128 x 8 TF32 MMA sustains one per 105.6 cycles, this form one per 64 -- tools/tc_rate.py).
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ---- mbarrier ----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  // try_wait suspends the thread until the phase completes or the hint (in ns) runs out; without a hint the
  // suspension is short and 13 waiting warps re-poll through the LSU data pipe the tensor core's operand reads
  // share (ncu: ~12 % of that pipe's wavefronts were polls).
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "MR_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra MR_DONE_%=;\n\t"
      "bra MR_WAIT_%=;\n\t"
      "MR_DONE_%=:\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(0x989680u)
      : "memory");
}
// Warp-level wait: one lane polls the barrier, the rest of the warp parks on __syncwarp (32x fewer
// waiters on the barrier word).
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity) {
  if ((threadIdx.x & 31) == 0) mbar_wait(bar, parity);
  __syncwarp();
}
// 1-D bulk async copy global -> shared (bytes % 16 == 0, both 16-byte aligned); completes on `bar`.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Pull a line into L2 ahead of use (no register or shared-memory cost; the producers' register prefetch
// only covers one chunk, L2 prefetch covers the DRAM part of the latency several chunks ahead).
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ---- operand tile addressing ---------------------------------------------------------------------
// Byte offset of element (r, c) of a tile stored as core matrices [row group][column chunk]:
//   ncc = number of 16-byte column chunks per row (cols / 4).
__device__ __forceinline__ uint32_t core_off_rg_major(int r, int c, int ncc) {
  return (uint32_t)(((r >> 3) * ncc + (c >> 2)) * 128 + (r & 7) * 16 + (c & 3) * 4);
}

// MN-major (SWIZZLE_128B_BASE32B) tile of a row-major source [krows x cols]: byte offset of element
// (r = K index, c = MN index) with the tile stored as [c/32 block][r/4 group][4 rows][128 B];
// ngroups = krows / 4.  Descriptor: LBO = ngroups * 512, SBO = 512; one K=8 MMA step spans 1024 bytes.
__device__ __forceinline__ uint32_t mn_off(int r, int c, int ngroups) {
  const int p = r & 3, e = c & 31;
  return (uint32_t)(((c >> 5) * ngroups + (r >> 2)) * 512 + p * 128 + ((((e >> 3) ^ p) & 3) << 5) + (e & 7) * 4);
}

}  // namespace tc
}  // namespace mr
