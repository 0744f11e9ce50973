// Embedding lookup out[i,:] = table[idx[i],:]  (replaces Embedding+Flatten, movierec/model.py:161-172).
// One warp per row, 128-bit coalesced loads/stores; HBM-bound: 8*dim bytes per row + 4 for the id.
#include "launchers.h"

namespace mr {

template <bool VEC>
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ table, int64_t rows, int dim,
                                                          const int32_t* __restrict__ idx, int64_t n,
                                                          float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
    const int r = __ldg(idx + i);
    const bool ok = (unsigned)r < (uint64_t)rows;
    const float* src = table + (size_t)(ok ? r : 0) * dim;
    float* dst = out + (size_t)i * dim;
    if (VEC) {
      for (int c = lane * 4; c < dim; c += 128) {
        float4 v = ok ? ld_stream4(src + c) : make_float4(nanf(""), nanf(""), nanf(""), nanf(""));
        *reinterpret_cast<float4*>(dst + c) = v;
      }
    } else {
      for (int c = lane; c < dim; c += 32) dst[c] = ok ? __ldg(src + c) : nanf("");
    }
  }
}

int launch_gather_rows(const float* table, int64_t rows, int dim, const int32_t* idx, int64_t n, float* out,
                       cudaStream_t st) {
  if (n == 0) return MR_OK;
  const bool vec = (dim & 3) == 0 && ((reinterpret_cast<uintptr_t>(table) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  int64_t blocks = (n + 7) / 8;  // 8 warps (rows) per CTA
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (vec) gather_rows_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(table, rows, dim, idx, n, out);
  else gather_rows_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(table, rows, dim, idx, n, out);
  MR_LAUNCH_CHECK("gather_rows_kernel");
  return MR_OK;
}

}  // namespace mr
