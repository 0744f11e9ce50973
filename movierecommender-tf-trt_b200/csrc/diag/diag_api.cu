// Diagnostics (libmovierec_b200_diag.so): the single-tile GEMM self-test of the fused train kernel's operand form
// (tc_bf16x3.cuh), plus the C entry points of every diagnostic and the error plumbing the launchers expect.
#include <stdarg.h>
#include <stdio.h>

#include "../launchers.h"
#include "../tc_bf16x3.cuh"
#include "movierec_b200_diag.h"

namespace mr {

static thread_local char g_diag_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_diag_err, sizeof(g_diag_err), fmt, ap);
  va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error in %s: %s", what, cudaGetErrorString(e));
  return MR_ERR_CUDA;
}
void count_launch() {}
int sm_count() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    return kB200Sms;
  }
  return n;
}

// ---- single-tile GEMM on the same operand layout, descriptors and split (tests/test_gpu_tc.py) ----------------------
//   D[128 x N] = A . B^T,  A = [128 x K] (K-major) or given as [K x 128] (MN-major), B = [N x K] or given as [K x N].
namespace fz {
__global__ void __launch_bounds__(128) bf16x3_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                              float* __restrict__ D, int N, int K, int a_mn, int b_mn) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // operand image: row-major source [R x C] -> panels of 64 columns, each [R x 128 B] (R padded to 8)
  const int Ra = a_mn ? K : 128, Ca = a_mn ? 128 : K;
  const int Rb = b_mn ? K : N, Cb = b_mn ? N : K;
  const uint32_t pa = (uint32_t)((Ra + 7) / 8 * 8) * 128, pb = (uint32_t)((Rb + 7) / 8 * 8) * 128;  // panel bytes
  const uint32_t parta = pa * ((Ca + 63) / 64), partb = pb * ((Cb + 63) / 64);
  uint8_t* a_img = smem;
  uint8_t* b_img = smem + 3 * parta;
  for (uint32_t e = tid; e < (3 * parta + 3 * partb) / 16; e += blockDim.x) reinterpret_cast<uint4*>(smem)[e] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  for (int e = tid; e < Ra * Ca / 2; e += blockDim.x) {
    const int r = e / (Ca / 2), c = 2 * (e - r * (Ca / 2));
    uint32_t w1, w2, w3;
    split3(A[(size_t)r * Ca + c], A[(size_t)r * Ca + c + 1], w1, w2, w3);
    const uint32_t off = (uint32_t)(c >> 6) * pa + sw128_off(r, c & 63);
    *reinterpret_cast<uint32_t*>(a_img + off) = w1;
    *reinterpret_cast<uint32_t*>(a_img + parta + off) = w2;
    *reinterpret_cast<uint32_t*>(a_img + 2 * parta + off) = w3;
  }
  for (int e = tid; e < Rb * Cb / 2; e += blockDim.x) {
    const int r = e / (Cb / 2), c = 2 * (e - r * (Cb / 2));
    uint32_t w1, w2, w3;
    split3(B[(size_t)r * Cb + c], B[(size_t)r * Cb + c + 1], w1, w2, w3);
    const uint32_t off = (uint32_t)(c >> 6) * pb + sw128_off(r, c & 63);
    *reinterpret_cast<uint32_t*>(b_img + off) = w1;
    *reinterpret_cast<uint32_t*>(b_img + partb + off) = w2;
    *reinterpret_cast<uint32_t*>(b_img + 2 * partb + off) = w3;
  }
  uint32_t ncols = 32;
  while (ncols < (uint32_t)N) ncols <<= 1;
  if (warp == 0) tc::tmem_alloc(&tmem_slot, ncols);
  if (tid == 0) {
    tc::mbar_init(&done_bar, 1);
    tc::mbar_init_fence();
  }
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = idesc_bf16(128, N, a_mn, b_mn);
    const uint64_t da = tc::smem_desc(0, a_mn ? pa : 16, 1024, kLayoutSw128);
    const uint64_t db = tc::smem_desc(0, b_mn ? pb : 16, 1024, kLayoutSw128);
    constexpr int PA[6] = {0, 0, 1, 0, 2, 1}, PB[6] = {0, 1, 0, 2, 0, 1};
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint32_t ao = a_mn ? (uint32_t)ks * 2048 : (uint32_t)(ks >> 2) * pa + (ks & 3) * 32;
      const uint32_t bo = b_mn ? (uint32_t)ks * 2048 : (uint32_t)(ks >> 2) * pb + (ks & 3) * 32;
      for (int q = 0; q < 6; ++q) {
        const uint32_t a = tc::smem_u32(a_img) + PA[q] * parta + ao;
        const uint32_t b = tc::smem_u32(b_img) + PB[q] * partb + bo;
        mma_bf16(tmem_base, da + (a >> 4), db + (b >> 4), idesc, (ks | q) != 0);
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
}  // namespace fz

int launch_bf16x3_selftest(const float* A, const float* B, float* D, int N, int K, int a_mn, int b_mn, cudaStream_t st) {
  if (N % 16 || N < 16 || N > 256 || K % 16 || K < 16 || K > 256 || (b_mn && N % 64)) {
    set_error("bf16x3 selftest: unsupported N=%d K=%d", N, K);
    return MR_ERR_INVALID;
  }
  const int Ra = a_mn ? K : 128, Ca = a_mn ? 128 : K, Rb = b_mn ? K : N, Cb = b_mn ? N : K;
  const size_t smem = 3 * ((size_t)((Ra + 7) / 8 * 8) * 128 * ((Ca + 63) / 64) + (size_t)((Rb + 7) / 8 * 8) * 128 * ((Cb + 63) / 64)) + 1024;
  if (smem > 220 * 1024) {
    set_error("bf16x3 selftest: operands too large");
    return MR_ERR_INVALID;
  }
  MR_CUDA(cudaFuncSetAttribute(fz::bf16x3_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  fz::bf16x3_selftest_kernel<<<1, 128, smem, st>>>(A, B, D, N, K, a_mn, b_mn);
  MR_LAUNCH_CHECK("bf16x3_selftest_kernel");
  return MR_OK;
}


}  // namespace mr

using namespace mr;

extern "C" {

const char* mr_diag_last_error(void) { return g_diag_err; }

int mr_tc_gemm_selftest(const float* A, const float* B, float* D, int32_t N, int32_t K, int32_t a_mn, int32_t b_mn,
                        int32_t three_x, void* stream) {
  MR_REQUIRE(A && B && D, "tc selftest: NULL pointer");
  return launch_tc_selftest(A, B, D, N, K, a_mn, b_mn, three_x, (cudaStream_t)stream);
}

int mr_bf16x3_gemm_selftest(const float* A, const float* B, float* D, int32_t N, int32_t K, int32_t a_mn, int32_t b_mn,
                            void* stream) {
  MR_REQUIRE(A && B && D, "bf16x3 selftest: NULL pointer");
  return launch_bf16x3_selftest(A, B, D, N, K, a_mn, b_mn, (cudaStream_t)stream);
}

int mr_tc_probe(const float* raw_a, int32_t n_words, int32_t start_off, int32_t lbo, int32_t sbo, int32_t a_mn, float* D,
                void* stream) {
  MR_REQUIRE(raw_a && D, "tc probe: NULL pointer");
  return launch_tc_probe(raw_a, n_words, start_off, lbo, sbo, a_mn, D, (cudaStream_t)stream);
}

int mr_tc_rate(int32_t N, int32_t iters, int32_t nbuf, int32_t flags, int32_t writers, int32_t write_iters,
               int64_t* out_cycles, int32_t grid, void* stream) {
  MR_REQUIRE(out_cycles != nullptr, "tc_rate: NULL output");
  return launch_tc_rate(N, iters, nbuf, flags, writers, write_iters, reinterpret_cast<long long*>(out_cycles), grid,
                        (cudaStream_t)stream);
}

}  // extern "C"
