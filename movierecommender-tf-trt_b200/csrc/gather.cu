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

// ---- row-sharded tables read over NVLink peer pointers ----------------------------------------------------------
// out[i,:] = shard[owner(id)][id / world,:] with owner(id) = id % world, where shard[r] is rank r's slice of the
// table in ITS memory (pointers exchanged through torch's symmetric memory, NVLink 5 / NVSwitch peer access).
// The gather and the "all-to-all of gathered rows" of a row-sharded step are this one kernel: no id exchange, no
// owner-side gather, no staging copies.  One warp per row, 128-bit loads; remote rows cross NVLink once.
template <bool VEC>
__global__ void __launch_bounds__(256) gather_rows_sharded_kernel(const float* const* __restrict__ shards, int world,
                                                                  int64_t total_rows, int dim,
                                                                  const int32_t* __restrict__ ids, int64_t n,
                                                                  float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
    const int r = __ldg(ids + i);
    const bool ok = (unsigned)r < (uint64_t)total_rows;
    const unsigned id = ok ? (unsigned)r : 0u;
    const unsigned local = id / (unsigned)world, owner = id - local * (unsigned)world;
    const float* src = shards[owner] + (size_t)local * dim;
    float* dst = out + (size_t)i * dim;
    if (VEC) {
      for (int c = lane * 4; c < dim; c += 128) {
        float4 v = ok ? *reinterpret_cast<const float4*>(src + c) : make_float4(nanf(""), nanf(""), nanf(""), nanf(""));
        *reinterpret_cast<float4*>(dst + c) = v;
      }
    } else {
      for (int c = lane; c < dim; c += 32) dst[c] = ok ? src[c] : nanf("");
    }
  }
}

int launch_gather_rows_sharded(const float* const* shards, int world, int64_t total_rows, int dim, const int32_t* ids,
                               int64_t n, float* out, cudaStream_t st) {
  if (n == 0) return MR_OK;
  const bool vec = (dim & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;  // shard bases are allocation starts
  int64_t blocks = (n + 7) / 8;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (vec) gather_rows_sharded_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(shards, world, total_rows, dim, ids, n, out);
  else gather_rows_sharded_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(shards, world, total_rows, dim, ids, n, out);
  MR_LAUNCH_CHECK("gather_rows_sharded_kernel");
  return MR_OK;
}

// ---- grouped batches --------------------------------------------------------------------------------------
// The reference's generator lays a batch out as groups of one positive and its negatives, all of one user
// (data_pipeline.py:99-150).  Everything the tower computes from the user row alone is therefore shared by
// the rows of a group: the user half of the first layer in the forward pass, and -- after summing the
// pre-activation gradients of the group -- the user half of the backward pass and of the weight gradient.

// out[g] = sum_{j < group} in[g * group + j], rows of `width` floats, added in order j = 0, 1, ...
// HBM-bound: (group + 1) * width * 4 bytes per group.
__global__ void __launch_bounds__(256) group_sum_rows_kernel(const float* __restrict__ in, int64_t groups, int group,
                                                             int width4, float* __restrict__ out) {
  const int64_t total = groups * width4;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t g = e / width4;
    const int c = (int)(e - g * width4);
    const float4* src = reinterpret_cast<const float4*>(in) + g * group * width4 + c;
    float4 s = ld_stream4(reinterpret_cast<const float*>(src));
    for (int j = 1; j < group; ++j) {
      const float4 v = ld_stream4(reinterpret_cast<const float*>(src + (int64_t)j * width4));
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    reinterpret_cast<float4*>(out)[e] = s;
  }
}

__global__ void __launch_bounds__(256) group_heads_kernel(const int32_t* __restrict__ ids, int64_t groups, int group,
                                                          int32_t* __restrict__ out) {
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (int64_t)gridDim.x * blockDim.x)
    out[g] = __ldg(ids + g * group);
}

__global__ void __launch_bounds__(256) check_grouped_kernel(const int32_t* __restrict__ ids, int64_t n, int group,
                                                            int32_t* __restrict__ flag) {
  bool bad = false;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x)
    bad |= __ldg(ids + r) != __ldg(ids + (r - r % group));
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

static unsigned grid_for(int64_t work_items) {
  int64_t blocks = (work_items + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  return (unsigned)(blocks < 1 ? 1 : blocks);
}

int launch_group_sum_rows(const float* in, int64_t groups, int group, int width, float* out, cudaStream_t st) {
  if (groups == 0) return MR_OK;
  if (width % 4 || group < 1 || ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15)) {
    set_error("group_sum_rows: width=%d must be a multiple of 4 and the buffers 16-byte aligned", width);
    return MR_ERR_INVALID;
  }
  group_sum_rows_kernel<<<grid_for(groups * (width / 4)), 256, 0, st>>>(in, groups, group, width / 4, out);
  MR_LAUNCH_CHECK("group_sum_rows_kernel");
  return MR_OK;
}

int launch_group_heads(const int32_t* ids, int64_t groups, int group, int32_t* out, cudaStream_t st) {
  if (groups == 0) return MR_OK;
  group_heads_kernel<<<grid_for(groups), 256, 0, st>>>(ids, groups, group, out);
  MR_LAUNCH_CHECK("group_heads_kernel");
  return MR_OK;
}

int launch_check_grouped(const int32_t* ids, int64_t n, int group, int32_t* flag, cudaStream_t st) {
  if (n == 0 || group <= 1) return MR_OK;
  check_grouped_kernel<<<grid_for(n), 256, 0, st>>>(ids, n, group, flag);
  MR_LAUNCH_CHECK("check_grouped_kernel");
  return MR_OK;
}

}  // namespace mr
