/* Diagnostics of libmovierec_b200 (sm_100a): NOT part of the product library.  Built into
 * libmovierec_b200_diag.so by `make diag` (csrc/Makefile, -DMR_DIAGNOSTICS) and loaded only by tests/test_gpu_tc.py
 * and tools/tc_probe.py, tools/tc_rate.py.  They validate, in isolation, the operand layouts, matrix / instruction
 * descriptors, TMEM allocation and the mbarrier + proxy-fence protocol the production kernels rely on. */
#ifndef MOVIEREC_B200_DIAG_H_
#define MOVIEREC_B200_DIAG_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* mr_diag_last_error(void);
/* Single-tile tcgen05 GEMM self-test: D[128 x N] = A . B^T on the tensor cores (TF32 or 3xTF32), A given
 * as [128 x K] (a_mn = 0) or [K x 128] (a_mn = 1), B as [N x K] (b_mn = 0) or [K x N] (b_mn = 1). */
int mr_tc_gemm_selftest(const float* A, const float* B, float* D, int32_t N, int32_t K, int32_t a_mn, int32_t b_mn,
                        int32_t three_x, void* stream);
/* The same single-tile GEMM on the operand form of the fused train kernel: three bf16 parts per operand, six part
 * products, SWIZZLE_128B tiles read K-major (x_mn = 0) or MN-major (x_mn = 1); K % 16 == 0, K <= 256. */
int mr_bf16x3_gemm_selftest(const float* A, const float* B, float* D, int32_t N, int32_t K, int32_t a_mn, int32_t b_mn,
                            void* stream);
/* Descriptor explorer: raw_a (n_words floats) is the shared-memory image of A, B is the 16x8 identity, one
 * M=128,N=16,K=8 TF32 MMA; D is [128 x 16]. */
int mr_tc_probe(const float* raw_a, int32_t n_words, int32_t start_off, int32_t lbo, int32_t sbo, int32_t a_mn,
                float* D, void* stream);
/* Sustained tcgen05 issue rate of the 3xTF32 stage pattern on static operands (tools/tc_rate.py). */
int mr_tc_rate(int32_t N, int32_t iters, int32_t nbuf, int32_t flags, int32_t writers, int32_t write_iters,
               int64_t* out_cycles, int32_t grid, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MOVIEREC_B200_DIAG_H_ */
