// RankLayer + HR@k / DCG@k (movierec/model.py:336-455) on device.
//   rank_of(i) = #{j : s_j > s_i} + #{j < i : s_j == s_i}      (descending, lower index first on ties --
//                                                               tf.nn.top_k order, test/test_model.py:168-190)
//   pos        = rank_of(label column)  (model.py:447-451), hit = pos < k (:454),
//   dcg        = ln2 / ln(pos + 2) * hit (:414-415).
// One warp per group; sums are reduced in a fixed two-level order (no atomics).
#include "launchers.h"

namespace mr {

constexpr int kRankThreads = 256;
constexpr int kRankGroupsPerCta = 1024;  // groups folded into one partial sum

// Label column of a group: label_col[g] when given, else argmax of the group's labels (first maximum, as
// K.argmax(y_true) at model.py:447-448) when `labels` is given, else the last column (the generator's layout).
__global__ void __launch_bounds__(kRankThreads) rank_positions_kernel(const float* __restrict__ scores, int64_t G,
                                                                      int group, const int32_t* __restrict__ label_col,
                                                                      const float* __restrict__ labels,
                                                                      int32_t* __restrict__ rank,
                                                                      int32_t* __restrict__ pos) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < G; g += warps) {
    const float* s = scores + g * group;
    int lc = label_col != nullptr ? __ldg(label_col + g) : group - 1;
    if (label_col == nullptr && labels != nullptr) {
      const float* y = labels + g * group;
      float best = -INFINITY;
      int bi = group;  // (a NaN label never wins; an all-NaN group falls back to column 0 below)
      for (int j = lane; j < group; j += 32) {
        const float v = __ldg(y + j);
        if (v > best) { best = v; bi = j; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
      }
      lc = bi >= group ? 0 : bi;
    }
    lc = min(max(lc, 0), group - 1);
    const float sp = rank_key(__ldg(s + lc));
    int cnt = 0;
    for (int j = lane; j < group; j += 32) {
      const float sj = rank_key(__ldg(s + j));
      cnt += (sj > sp) || (sj == sp && j < lc);
    }
    cnt = warp_sum_int(cnt);
    if (lane == 0) pos[g] = cnt;
    if (rank != nullptr) {
      for (int i = lane; i < group; i += 32) {
        const float si = rank_key(__ldg(s + i));
        int r = 0;
        for (int j = 0; j < group; ++j) {
          const float sj = rank_key(__ldg(s + j));
          r += (sj > si) || (sj == si && j < i);
        }
        rank[g * group + r] = i;
      }
    }
  }
}

// Small groups (the train step: one positive + a few negatives): ONE THREAD per group -- a warp-per-group launch leaves
// 27 of 32 lanes idle at group = 5 and spends its time in shuffles.  A warp's 32 groups are 32 * group consecutive
// floats, so the `group` strided loads of a thread hit the lines its neighbours load too (L1).  The CTA's
// kRankGroupsPerCta groups are folded into partial[blockIdx] = {hits, dcg} in a fixed order, as below.
template <int MAXG>
__global__ void __launch_bounds__(kRankThreads) rank_small_kernel(const float* __restrict__ scores, int64_t G, int group,
                                                                  const int32_t* __restrict__ label_col,
                                                                  const float* __restrict__ labels, int k,
                                                                  int32_t* __restrict__ pos, float* __restrict__ partial) {
  __shared__ float red_h[kRankThreads / 32], red_d[kRankThreads / 32];
  const int64_t lo = (int64_t)blockIdx.x * kRankGroupsPerCta;
  const int64_t hi = min(G, lo + kRankGroupsPerCta);
  float h = 0.f, d = 0.f;
  const float ln2 = logf(2.f);
  for (int64_t g = lo + threadIdx.x; g < hi; g += blockDim.x) {
    float sc[MAXG];
    int lc = label_col != nullptr ? __ldg(label_col + g) : group - 1;
    float best = -INFINITY;
    int bi = group;
#pragma unroll
    for (int j = 0; j < MAXG; ++j) {
      if (j < group) {
        sc[j] = rank_key(__ldg(scores + g * group + j));
        if (label_col == nullptr && labels != nullptr) {
          const float v = __ldg(labels + g * group + j);
          if (v > best) { best = v; bi = j; }
        }
      }
    }
    if (label_col == nullptr && labels != nullptr) lc = bi >= group ? 0 : bi;
    lc = min(max(lc, 0), group - 1);
    float sp = 0.f;
#pragma unroll
    for (int j = 0; j < MAXG; ++j)
      if (j == lc) sp = sc[j];
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < MAXG; ++j)
      if (j < group) cnt += (sc[j] > sp) || (sc[j] == sp && j < lc);
    pos[g] = cnt;
    if (cnt < k) {
      h += 1.f;
      d += ln2 / logf((float)cnt + 2.f);
    }
  }
  if (partial == nullptr) return;
  h = warp_sum(h);
  d = warp_sum(d);
  if ((threadIdx.x & 31) == 0) {
    red_h[threadIdx.x >> 5] = h;
    red_d[threadIdx.x >> 5] = d;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float hh = 0.f, dd = 0.f;
    for (int w = 0; w < kRankThreads / 32; ++w) {
      hh += red_h[w];
      dd += red_d[w];
    }
    partial[2 * blockIdx.x] = hh;
    partial[2 * blockIdx.x + 1] = dd;
  }
}

// partial[b] = {hits, dcg} of groups [b*1024, (b+1)*1024): fixed strided order + shuffle tree.
__global__ void __launch_bounds__(kRankThreads) rank_metric_partial_kernel(const int32_t* __restrict__ pos, int64_t G,
                                                                           int k, float* __restrict__ partial) {
  __shared__ float red_h[kRankThreads / 32], red_d[kRankThreads / 32];
  const int64_t lo = (int64_t)blockIdx.x * kRankGroupsPerCta;
  const int64_t hi = min(G, lo + kRankGroupsPerCta);
  float h = 0.f, d = 0.f;
  const float ln2 = logf(2.f);
  for (int64_t g = lo + threadIdx.x; g < hi; g += blockDim.x) {
    const int p = pos[g];
    if (p < k) {
      h += 1.f;
      d += ln2 / logf((float)p + 2.f);
    }
  }
  h = warp_sum(h);
  d = warp_sum(d);
  if ((threadIdx.x & 31) == 0) {
    red_h[threadIdx.x >> 5] = h;
    red_d[threadIdx.x >> 5] = d;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float hh = 0.f, dd = 0.f;
    for (int w = 0; w < kRankThreads / 32; ++w) {
      hh += red_h[w];
      dd += red_d[w];
    }
    partial[2 * blockIdx.x] = hh;
    partial[2 * blockIdx.x + 1] = dd;
  }
}

__global__ void __launch_bounds__(1024) rank_metric_final_kernel(const float* __restrict__ partial, int64_t nb,
                                                                 float* __restrict__ sums) {
  __shared__ double red_h[32], red_d[32];
  double h = 0.0, d = 0.0;
  for (int64_t b = threadIdx.x; b < nb; b += blockDim.x) {
    h += partial[2 * b];
    d += partial[2 * b + 1];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    h += __shfl_xor_sync(0xffffffffu, h, o);
    d += __shfl_xor_sync(0xffffffffu, d, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red_h[threadIdx.x >> 5] = h;
    red_d[threadIdx.x >> 5] = d;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double hh = 0.0, dd = 0.0;
    for (int w = 0; w < 32; ++w) {
      hh += red_h[w];
      dd += red_d[w];
    }
    sums[0] = (float)hh;
    sums[1] = (float)dd;
  }
}

size_t rank_partials_count(int64_t G) {
  const int64_t nb = (G + kRankGroupsPerCta - 1) / kRankGroupsPerCta;
  return (size_t)(nb < 1 ? 1 : nb) * 2;
}

int launch_rank_metrics(const int32_t* pos, int64_t G, int k, float* sums, float* partials, cudaStream_t st) {
  if (sums == nullptr) return MR_OK;
  if (G == 0) {
    MR_CUDA(cudaMemsetAsync(sums, 0, 2 * sizeof(float), st));
    return MR_OK;
  }
  const int64_t nb = (G + kRankGroupsPerCta - 1) / kRankGroupsPerCta;
  rank_metric_partial_kernel<<<(unsigned)nb, kRankThreads, 0, st>>>(pos, G, k, partials);
  MR_LAUNCH_CHECK("rank_metric_partial_kernel");
  rank_metric_final_kernel<<<1, 1024, 0, st>>>(partials, nb, sums);
  MR_LAUNCH_CHECK("rank_metric_final_kernel");
  return MR_OK;
}

int launch_rank_scores(const float* scores, int64_t G, int group, int k, const int32_t* label_col, int32_t* rank,
                       int32_t* pos, float* sums, float* partials, cudaStream_t st, const float* labels) {
  if (G == 0) {
    if (sums != nullptr) MR_CUDA(cudaMemsetAsync(sums, 0, 2 * sizeof(float), st));
    return MR_OK;
  }
  if (group <= 8 && rank == nullptr) {  // thread per group, positions + metric partials in one kernel
    const int64_t nb = (G + kRankGroupsPerCta - 1) / kRankGroupsPerCta;
    rank_small_kernel<8><<<(unsigned)nb, kRankThreads, 0, st>>>(scores, G, group, label_col, labels, k, pos,
                                                               sums != nullptr ? partials : nullptr);
    MR_LAUNCH_CHECK("rank_small_kernel");
    if (sums != nullptr) {
      rank_metric_final_kernel<<<1, 1024, 0, st>>>(partials, nb, sums);
      MR_LAUNCH_CHECK("rank_metric_final_kernel");
    }
    return MR_OK;
  }
  int64_t blocks = (G + 7) / 8;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  rank_positions_kernel<<<(unsigned)blocks, kRankThreads, 0, st>>>(scores, G, group, label_col, labels, rank, pos);
  MR_LAUNCH_CHECK("rank_positions_kernel");
  if (sums != nullptr) {
    const int64_t nb = (G + kRankGroupsPerCta - 1) / kRankGroupsPerCta;
    rank_metric_partial_kernel<<<(unsigned)nb, kRankThreads, 0, st>>>(pos, G, k, partials);
    MR_LAUNCH_CHECK("rank_metric_partial_kernel");
    rank_metric_final_kernel<<<1, 1024, 0, st>>>(partials, nb, sums);
    MR_LAUNCH_CHECK("rank_metric_final_kernel");
  }
  return MR_OK;
}

}  // namespace mr
