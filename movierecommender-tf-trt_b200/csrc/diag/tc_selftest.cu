// Single-tile tcgen05 GEMM used to validate the operand layouts, matrix/instruction descriptors,
// TMEM allocation/loads and the mbarrier + proxy-fence protocol in isolation (tests/test_gpu_tc.py).
//   D[128 x N] = A . B^T   with A = [128 x K] (K-major) or given as [K x 128] (MN-major),
//                               B = [N x K]   (K-major) or given as [K x N]   (MN-major),
// fp32 in/out, computed as 3xTF32 (three_x = 1) or plain TF32 (three_x = 0).
#include "../launchers.h"
#include "../tc_common.cuh"

namespace mr {

__global__ void __launch_bounds__(128) tc_gemm_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                               float* __restrict__ D, int N, int K, int a_mn, int b_mn,
                                                               int three_x) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int a_elems = 128 * K, b_elems = N * K;
  uint8_t* a_hi = smem_raw;
  uint8_t* a_lo = a_hi + a_elems * 4;
  uint8_t* b_hi = a_lo + a_elems * 4;
  uint8_t* b_lo = b_hi + b_elems * 4;

  uint32_t ncols = 32;
  while (ncols < (uint32_t)N) ncols <<= 1;
  if (warp == 0) tc::tmem_alloc(&tmem_slot, ncols);
  if (tid == 0) {
    tc::mbar_init(&done_bar, 1);
    tc::mbar_init_fence();
  }

  // fill operands: source is row-major [R x C]; a_mn/b_mn only change what R and C mean
  {
    const int R = a_mn ? K : 128, Cc = a_mn ? 128 : K;
    for (int e = tid; e < R * Cc; e += blockDim.x) {
      const int r = e / Cc, c = e - r * Cc;
      float hi, lo;
      if (three_x == 2) tc::split_tf32_fast(A[e], hi, lo);  // the activation-path split
      else tc::split_tf32(A[e], hi, lo);
      const uint32_t off = a_mn ? tc::mn_off(r, c, K >> 2) : tc::core_off_rg_major(r, c, Cc >> 2);
      *reinterpret_cast<float*>(a_hi + off) = hi;
      *reinterpret_cast<float*>(a_lo + off) = lo;
    }
  }
  {
    const int R = b_mn ? K : N, Cc = b_mn ? N : K;
    for (int e = tid; e < R * Cc; e += blockDim.x) {
      const int r = e / Cc, c = e - r * Cc;
      float hi, lo;
      tc::split_tf32(B[e], hi, lo);
      const uint32_t off = b_mn ? tc::mn_off(r, c, K >> 2) : tc::core_off_rg_major(r, c, Cc >> 2);
      *reinterpret_cast<float*>(b_hi + off) = hi;
      *reinterpret_cast<float*>(b_lo + off) = lo;
    }
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = tmem_slot;

  if (tid == 0) {
    const uint32_t idesc = tc::idesc_tf32(128, N, a_mn, b_mn);
    const uint32_t a_lbo = a_mn ? (K >> 2) * 512 : 128, a_sbo = a_mn ? 512 : (K >> 2) * 128;
    const uint32_t b_lbo = b_mn ? (K >> 2) * 512 : 128, b_sbo = b_mn ? 512 : (K >> 2) * 128;
    const uint32_t a_step = a_mn ? 1024 : 2 * 128;  // per k-step of 8
    const uint32_t b_step = b_mn ? 1024 : 2 * 128;
    const uint32_t a_lt = a_mn ? tc::kLayoutSw128Base32 : tc::kLayoutNone;
    const uint32_t b_lt = b_mn ? tc::kLayoutSw128Base32 : tc::kLayoutNone;
    uint32_t acc = 0;
    for (int kk = 0; kk < K / 8; ++kk) {
      const uint64_t ah = tc::smem_desc(tc::smem_u32(a_hi) + kk * a_step, a_lbo, a_sbo, a_lt);
      const uint64_t al = tc::smem_desc(tc::smem_u32(a_lo) + kk * a_step, a_lbo, a_sbo, a_lt);
      const uint64_t bh = tc::smem_desc(tc::smem_u32(b_hi) + kk * b_step, b_lbo, b_sbo, b_lt);
      const uint64_t bl = tc::smem_desc(tc::smem_u32(b_lo) + kk * b_step, b_lbo, b_sbo, b_lt);
      tc::mma_tf32(tmem_base, ah, bh, idesc, acc);
      acc = 1;
      if (three_x) {
        tc::mma_tf32(tmem_base, al, bh, idesc, 1);
        tc::mma_tf32(tmem_base, ah, bl, idesc, 1);
      }
    }
    tc::mma_commit(&done_bar);
  }
  tc::mbar_wait(&done_bar, 0);
  tc::fence_after_sync();

  const int row = 32 * warp + lane;
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tc::tmem_ld16(tmem_base + ((uint32_t)(32 * warp) << 16) + c0, v);
#pragma unroll
    for (int i = 0; i < 16; ++i) D[(size_t)row * N + c0 + i] = v[i];
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, ncols);
}

// Descriptor explorer: A's shared-memory image is given verbatim (raw words), B is the 16x8 identity
// (K-major), one M=128,N=16,K=8 TF32 MMA runs with the caller's descriptor fields, so D[m][k] shows the
// word the hardware fetched for logical element (m, k).  Used only by tests to pin layout semantics.
__global__ void __launch_bounds__(128) tc_probe_kernel(const float* __restrict__ raw_a, int n_words, int start_off,
                                                       int lbo, int sbo, int a_mn, float* __restrict__ D) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* a_img = reinterpret_cast<float*>(smem_raw);
  uint8_t* b_img = smem_raw + (size_t)n_words * 4;
  if (warp == 0) tc::tmem_alloc(&tmem_slot, 32);
  if (tid == 0) {
    tc::mbar_init(&done_bar, 1);
    tc::mbar_init_fence();
  }
  for (int e = tid; e < n_words; e += blockDim.x) a_img[e] = raw_a[e];
  for (int e = tid; e < 16 * 8; e += blockDim.x) {
    const int n = e / 8, k = e - n * 8;
    *reinterpret_cast<float*>(b_img + tc::core_off_rg_major(n, k, 2)) = (n == k) ? 1.f : 0.f;
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = tc::idesc_tf32(128, 16, a_mn, 0);
    const uint64_t ad = tc::smem_desc(tc::smem_u32(a_img) + start_off, lbo & 0xFFFFF, sbo) | ((uint64_t)(lbo >> 24) << 61);  // lbo bits 24+ carry the layout type (probe only)
    const uint64_t bd = tc::smem_desc(tc::smem_u32(b_img), 128, 256);
    tc::mma_tf32(tmem_base, ad, bd, idesc, 0);
    tc::mma_commit(&done_bar);
  }
  tc::mbar_wait(&done_bar, 0);
  tc::fence_after_sync();
  float v[16];
  tc::tmem_ld16(tmem_base + ((uint32_t)(32 * warp) << 16), v);
  for (int i = 0; i < 16; ++i) D[(32 * warp + lane) * 16 + i] = v[i];
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, 32);
}

int launch_tc_probe(const float* raw_a, int n_words, int start_off, int lbo, int sbo, int a_mn, float* D,
                    cudaStream_t st) {
  const size_t smem = (size_t)n_words * 4 + 1024;
  if (n_words < 0 || smem > 200 * 1024 || (n_words & 255)) {
    set_error("tc probe: bad n_words=%d", n_words);
    return MR_ERR_INVALID;
  }
  MR_CUDA(cudaFuncSetAttribute(tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_probe_kernel<<<1, 128, smem, st>>>(raw_a, n_words, start_off, lbo, sbo, a_mn, D);
  MR_LAUNCH_CHECK("tc_probe_kernel");
  return MR_OK;
}

int launch_tc_selftest(const float* A, const float* B, float* D, int N, int K, int a_mn, int b_mn, int three_x,
                       cudaStream_t st) {
  const size_t smem = (size_t)(128 + N) * K * 4 * 2;
  if (N % 16 || N < 16 || N > 256 || K % 8 || K < 8 || smem > 200 * 1024 || (b_mn && N % 32)) {
    set_error("tc selftest: unsupported N=%d K=%d", N, K);
    return MR_ERR_INVALID;
  }
  MR_CUDA(cudaFuncSetAttribute(tc_gemm_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_gemm_selftest_kernel<<<1, 128, smem, st>>>(A, B, D, N, K, a_mn, b_mn, three_x);
  MR_LAUNCH_CHECK("tc_gemm_selftest_kernel");
  return MR_OK;
}

}  // namespace mr

// ---- tensor-pipe rate probe (tools/tc_rate.py; diagnostics only) --------------------------------------------
// One CTA per SM issues `iters` pipeline stages of 4 k-steps x 3 MMAs (the 3xTF32 pattern of the real kernels)
// on static operands rotating over `nbuf` shared-memory stage buffers, with no producer handshake, and
// reports clock64 cycles from the first issue to the completion of the last MMA.
//   flags bit 0: A operand from TMEM (tcgen05.mma [d], [a], b-desc) instead of shared memory
//         bit 1: MN-major operands (SWIZZLE_128B_BASE32B descriptors, 16-row K chunks as in tc_wgrad)
//   writers: extra warps that stream st.shared.v4 into a scratch area while the MMAs run (producer traffic)
namespace mr {

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

__global__ void __launch_bounds__(32 * 9, 1) tc_rate_kernel(int N, int iters, int nbuf, int flags, int writers,
                                                            int write_iters, long long* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t done_bar, side_bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t a_bytes = 128 * 32 * 4, b_bytes = (uint32_t)N * 32 * 4;
  const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
  const uint32_t scratch_off = (uint32_t)nbuf * stage_bytes;
  for (uint32_t e = tid; e < (scratch_off + 16384) / 16; e += blockDim.x)
    reinterpret_cast<float4*>(smem)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (warp == 0) tc::tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    tc::mbar_init(&done_bar, 1);
    tc::mbar_init(&side_bar, 1);
    tc::mbar_init_fence();
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = tmem_slot;
  const bool ts = flags & 1, mn = flags & 2;
  if (tid == 0) {
    const uint32_t idesc = tc::idesc_tf32(128, N, (mn && !ts) ? 1 : 0, mn ? 1 : 0);  // A in TMEM has no major-ness
    const uint32_t lbo = mn ? 8 * 512 : 128, sbo = mn ? 512 : 1024, lt = mn ? tc::kLayoutSw128Base32 : tc::kLayoutNone;
    const uint32_t kstep = mn ? 1024 : 256;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t sa = tc::smem_u32(smem) + (uint32_t)(it % nbuf) * stage_bytes;
      const uint32_t sb = sa + 2 * a_bytes;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint64_t bh = tc::smem_desc(sb + kk * kstep, lbo, sbo, lt);
        const uint64_t bl = tc::smem_desc(sb + b_bytes + kk * kstep, lbo, sbo, lt);
        if (ts) {
          // flags bit 3: A right behind the accumulator (column 128) instead of column 256; bit 4: A ring of 4 stages
          const uint32_t abase = tmem_base + ((flags & 8) ? 128 : 256) + ((flags & 16) ? (uint32_t)(it & 3) * 64 : 0);
          const uint32_t ah = abase + 8 * kk, al = abase + 32 + 8 * kk;
          mma_tf32_ts(tmem_base, ah, bh, idesc, 1);
          mma_tf32_ts(tmem_base, al, bh, idesc, 1);
          mma_tf32_ts(tmem_base, ah, bl, idesc, 1);
        } else {
          const uint64_t ah = tc::smem_desc(sa + kk * kstep, lbo, sbo, lt);
          const uint64_t al = tc::smem_desc(sa + a_bytes + kk * kstep, lbo, sbo, lt);
          tc::mma_tf32(tmem_base, ah, bh, idesc, 1);
          tc::mma_tf32(tmem_base, al, bh, idesc, 1);
          tc::mma_tf32(tmem_base, ah, bl, idesc, 1);
        }
      }
      if (flags & 4) tc::mma_commit(&side_bar);  // a commit per stage, as the pipelined kernels issue them
    }
    tc::mma_commit(&done_bar);
    tc::mbar_wait(&done_bar, 0);
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  } else if (warp >= 1 && warp <= writers) {
    float4* dst = reinterpret_cast<float4*>(smem + scratch_off) + (tid & 31) + 32 * ((warp - 1) & 7) * 4;
    const float4 v = make_float4(1.f, 2.f, 3.f, (float)tid);
    for (int i = 0; i < write_iters; ++i) {
#pragma unroll
      for (int q = 0; q < 4; ++q) asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(tc::smem_u32(dst + 32 * q)), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, 512);
}

int launch_tc_rate(int N, int iters, int nbuf, int flags, int writers, int write_iters, long long* out, int grid,
                   cudaStream_t st) {
  const size_t smem = (size_t)nbuf * (2 * 128 * 32 * 4 + 2 * (size_t)N * 32 * 4) + 16384 + 1024;
  if (N % 16 || N < 16 || N > 256 || nbuf < 1 || smem > 227 * 1024 || writers < 0 || writers > 8 || grid < 1) {
    set_error("tc rate: bad arguments");
    return MR_ERR_INVALID;
  }
  MR_CUDA(cudaFuncSetAttribute(tc_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_rate_kernel<<<grid, 32 * 9, smem, st>>>(N, iters, nbuf, flags, writers, write_iters, out);
  MR_LAUNCH_CHECK("tc_rate_kernel");
  return MR_OK;
}

}  // namespace mr
