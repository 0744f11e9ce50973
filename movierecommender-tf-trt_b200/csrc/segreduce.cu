// Deterministic segmented reduction of staged per-sample embedding-row gradients, fused with the
// row update -- the "segmented-reduce scatter-add, with no atomics, feeding a fused sparse-row Adam
// update" of the train step.  Input: sample indices stably sorted by row id (radix_sort.cu), so the
// samples of one row are contiguous and in batch order -- the order in which TensorFlow's
// IndexedSlices densification sums duplicates.
//
// A warp takes 32 consecutive sorted entries, finds the segments that START among them and owns each
// of those to its end (which may lie in later chunks).  Lanes span the columns of the row, so every
// staged row is read with coalesced 128-byte requests, and the samples of a segment are added
// sequentially -> bit-identical results on every run.
#include "launchers.h"

namespace mr {

__device__ __forceinline__ void apply_row_value(const RowUpdate& u, int row, int c, float g) {
  float *p, *m, *v, *gt;
  int d, col;
  if (c < u.d0) {
    p = u.p0; m = u.m0; v = u.v0; gt = u.g0; d = u.d0; col = c;
  } else {
    p = u.p1; m = u.m1; v = u.v1; gt = u.g1; d = u.d1; col = c - u.d0;
  }
  const size_t at = (size_t)row * d + col;
  if (u.mode == MR_TABLES_DENSE) {
    gt[at] = g;  // the dense Adam sweep (optimizer.cu) consumes the gradient table
  } else if (u.optimizer == MR_OPT_ADAM) {
    const float mn = u.beta_1 * m[at] + (1.f - u.beta_1) * g;
    const float vn = u.beta_2 * v[at] + (1.f - u.beta_2) * g * g;
    m[at] = mn;
    v[at] = vn;
    p[at] = p[at] - u.lr_t * mn / (sqrtf(vn) + u.epsilon);
  } else {
    p[at] = p[at] - u.lr * g;
  }
}

__global__ void __launch_bounds__(256) segreduce_kernel(const int32_t* __restrict__ keys,
                                                        const int32_t* __restrict__ index, int64_t n,
                                                        const float* __restrict__ staged, const RowUpdate u) {
  const int lane = threadIdx.x & 31;
  const int ld = u.d0 + u.d1;
  const int64_t nchunks = (n + 31) >> 5;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t chunk = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; chunk < nchunks; chunk += warps) {
    const int64_t e = chunk * 32 + lane;
    const int key = e < n ? __ldg(keys + e) : -1;
    const int prev = (e > 0 && e < n) ? __ldg(keys + e - 1) : -2;
    unsigned starts = __ballot_sync(0xffffffffu, e < n && key != prev);
    while (starts != 0) {
      const int sl = __ffs(starts) - 1;
      starts &= starts - 1;
      const int row = __shfl_sync(0xffffffffu, key, sl);
      const int64_t s = chunk * 32 + sl;
      // end of the segment: first later entry whose key differs (entries past n count as different)
      int64_t end;
      const unsigned diff = __ballot_sync(0xffffffffu, lane > sl && key != row);
      if (diff != 0) {
        end = chunk * 32 + (__ffs(diff) - 1);
      } else {
        int64_t j = (chunk + 1) * 32;
        for (;;) {
          const int64_t e2 = j + lane;
          const int k2 = e2 < n ? __ldg(keys + e2) : -1;
          const unsigned d2 = __ballot_sync(0xffffffffu, k2 != row);
          if (d2 != 0) {
            end = j + (__ffs(d2) - 1);
            break;
          }
          j += 32;
        }
      }
      for (int cb = 0; cb < ld; cb += 32) {
        const int c = cb + lane;
        const bool on = c < ld;
        float acc = 0.f;
        int64_t j = s;
        for (; j + 4 <= end; j += 4) {
          const int i0 = __ldg(index + j), i1 = __ldg(index + j + 1), i2 = __ldg(index + j + 2),
                    i3 = __ldg(index + j + 3);
          float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
          if (on) {
            v0 = __ldg(staged + (size_t)i0 * ld + c);
            v1 = __ldg(staged + (size_t)i1 * ld + c);
            v2 = __ldg(staged + (size_t)i2 * ld + c);
            v3 = __ldg(staged + (size_t)i3 * ld + c);
          }
          acc += v0;
          acc += v1;
          acc += v2;
          acc += v3;
        }
        for (; j < end; ++j) {
          const int i0 = __ldg(index + j);
          if (on) acc += __ldg(staged + (size_t)i0 * ld + c);
        }
        if (on) apply_row_value(u, row, c, acc);
      }
    }
  }
}

int launch_segreduce(const int32_t* sorted_keys, const int32_t* sorted_index, int64_t n, const float* staged,
                     const RowUpdate& u, cudaStream_t st) {
  if (n == 0) return MR_OK;
  const int64_t nchunks = (n + 31) / 32;
  int64_t blocks = (nchunks + 7) / 8;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  segreduce_kernel<<<(unsigned)blocks, 256, 0, st>>>(sorted_keys, sorted_index, n, staged, u);
  MR_LAUNCH_CHECK("segreduce_kernel");
  return MR_OK;
}

}  // namespace mr
