// Embedding lookup out[i,:] = table[idx[i],:]  (replaces Embedding+Flatten, movierec/model.py:161-172).
// One warp per row, 128-bit coalesced loads/stores; HBM-bound: 8*dim bytes per row + 4 for the id.
#include <stdlib.h>

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
                                                             int width4, int ld4, float* __restrict__ out, int out_ld4) {
  const int64_t total = groups * width4;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t g = e / width4;
    const int c = (int)(e - g * width4);
    const float4* src = reinterpret_cast<const float4*>(in) + g * group * ld4 + c;
    float4 s = ld_stream4(reinterpret_cast<const float*>(src));
    for (int j = 1; j < group; ++j) {
      const float4 v = ld_stream4(reinterpret_cast<const float*>(src + (int64_t)j * ld4));
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    reinterpret_cast<float4*>(out)[g * out_ld4 + c] = s;
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

int launch_group_sum_rows(const float* in, int64_t groups, int group, int width, float* out, cudaStream_t st,
                          int in_ld, int out_ld) {
  if (groups == 0) return MR_OK;
  if (in_ld <= 0) in_ld = width;
  if (out_ld <= 0) out_ld = width;
  if (width % 4 || in_ld % 4 || in_ld < width || out_ld % 4 || out_ld < width || group < 1 ||
      ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15)) {
    set_error("group_sum_rows: width=%d / stride=%d must be multiples of 4 and the buffers 16-byte aligned", width, in_ld);
    return MR_ERR_INVALID;
  }
  group_sum_rows_kernel<<<grid_for(groups * (width / 4)), 256, 0, st>>>(in, groups, group, width / 4, in_ld / 4, out,
                                                                        out_ld / 4);
  MR_LAUNCH_CHECK("group_sum_rows_kernel");
  return MR_OK;
}

// ---- item-projected first layer ------------------------------------------------------------------------------
// The first Dense layer is linear before its ReLU, so its item half depends on the item alone:
//   z1[r] = E_item[i_r] . W1[item rows] + (E_user[u_r] . W1[user rows] + b1) = Pi[i_r] + Zu[group of r].
// When a step has many more rows than there are items (ML-20M shape: 1.3 M rows, 26,744 items) Pi is one small GEMM
// over the item table and the per-row part of the layer is this gather: one warp per row, 128-bit loads of the
// L2-resident Pi row and of the group's Zu row, ReLU, a coalesced store of H1 and the ReLU bits of the backward pass.
// HBM-bound: 4 * width + width / 8 + 4 bytes written / read per row (Pi and Zu rows hit L2).
template <int ROWS>
__global__ void __launch_bounds__(256, 4) h1_from_projection_kernel(const float* __restrict__ Pi, int32_t num_items,
                                                                    const int32_t* __restrict__ items, int64_t row0,
                                                                    int64_t rows, const float* __restrict__ Zu, int group,
                                                                    const int32_t* __restrict__ users, int32_t num_users,
                                                                    int width, float* __restrict__ H1,
                                                                    uint32_t* __restrict__ bits) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int words = width >> 5;
  const int c0 = blockIdx.y * 128;  // a warp covers 128 columns of its rows; wider layers take gridDim.y slabs
  const int col = c0 + 4 * lane;
  const bool active = col < width;
  items += row0;
  if (users != nullptr) users += row0;
  for (int64_t base = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * ROWS; base < rows; base += warps * ROWS) {
    // ids first (every row's loads in flight together), then the rows.  The rows of a group share the user, so the
    // user-side row is loaded once per run of equal source rows (warp-uniform test): with groups of 5 that takes the
    // L2 reads of this kernel -- its bound, ~8 TB/s of L2 traffic in the first version -- from 2 to ~1.4 rows per
    // output row, with the 100-candidate groups of the ranking eval to 1.25.
    int it[ROWS], zi[ROWS];  // zi: source row of the user-side addend, -1: none (row past the end / id out of range)
#pragma unroll
    for (int j = 0; j < ROWS; ++j) {
      const int64_t r = base + j;
      it[j] = -1;
      zi[j] = -1;
      if (r < rows) {
        it[j] = __ldg(items + r);
        if (users == nullptr) {
          zi[j] = (int)((uint32_t)r / (uint32_t)group);
        } else {  // Zu holds one row per USER (user-projected first layer)
          const int u = __ldg(users + r);
          if ((unsigned)u < (unsigned)num_users) zi[j] = u;
        }
      }
    }
    float4 a[ROWS], z[ROWS];
#pragma unroll
    for (int j = 0; j < ROWS; ++j) {
      a[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      z[j] = a[j];
      if (active) {
        if ((unsigned)it[j] < (unsigned)num_items) a[j] = ldg4(Pi + (size_t)it[j] * width + col);  // bad ids: a zero row
        if (zi[j] >= 0 && (j == 0 || zi[j] != zi[j - 1])) z[j] = ldg4(Zu + (size_t)zi[j] * width + col);
      }
    }
    // copies only after every load has been issued (a copy waits for its source row: placed in the loop above it
    // exposed one memory latency per repeated row and made the kernel 40 % slower)
#pragma unroll
    for (int j = 1; j < ROWS; ++j)
      if (zi[j] >= 0 && zi[j] == zi[j - 1]) z[j] = z[j - 1];
#pragma unroll
    for (int j = 0; j < ROWS; ++j) {
      const int64_t r = base + j;
      float4 v;
      v.x = fmaxf(a[j].x + z[j].x, 0.f);
      v.y = fmaxf(a[j].y + z[j].y, 0.f);
      v.z = fmaxf(a[j].z + z[j].z, 0.f);
      v.w = fmaxf(a[j].w + z[j].w, 0.f);
      if (r < rows && active) *reinterpret_cast<float4*>(H1 + (size_t)r * width + col) = v;
      if (bits != nullptr) {  // word q of a row = columns [32 q, 32 q + 32): the 4-bit pieces of lanes 8 q .. 8 q + 7
        uint32_t w = (v.x > 0.f ? 1u : 0u) | (v.y > 0.f ? 2u : 0u) | (v.z > 0.f ? 4u : 0u) | (v.w > 0.f ? 8u : 0u);
        w = active ? w << (4 * (lane & 7)) : 0u;
        w |= __shfl_xor_sync(0xffffffffu, w, 1);
        w |= __shfl_xor_sync(0xffffffffu, w, 2);
        w |= __shfl_xor_sync(0xffffffffu, w, 4);
        const int q = (c0 >> 5) + (lane >> 3);
        if (r < rows && (lane & 7) == 0 && q < words) bits[(size_t)r * words + q] = w;
      }
    }
  }
}

int launch_h1_from_projection(const float* Pi, int32_t num_items, const int32_t* items, int64_t row0, int64_t rows,
                              const float* Zu, int group, const int32_t* users, int32_t num_users, int width, float* H1,
                              uint32_t* bits, cudaStream_t st) {
  if (rows == 0) return MR_OK;
  if (width % 32 || width < 32 || group < 1 ||
      ((reinterpret_cast<uintptr_t>(Pi) | reinterpret_cast<uintptr_t>(Zu) | reinterpret_cast<uintptr_t>(H1)) & 15)) {
    set_error("h1_from_projection: width=%d must be a multiple of 32 and the buffers 16-byte aligned", width);
    return MR_ERR_INVALID;
  }
  constexpr int kRows = 4;  // rows per warp in flight
  int64_t blocks = (rows + 8 * kRows - 1) / (8 * kRows);
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  const dim3 grid((unsigned)blocks, (unsigned)((width + 127) / 128));
  // (evict-first stores of H1, to keep Pi / Pu in L2, measured 2 % slower than plain stores on the ML-20M step)
  h1_from_projection_kernel<kRows><<<grid, 256, 0, st>>>(Pi, num_items, items, row0, rows, Zu, group, users, num_users,
                                                         width, H1, bits);
  MR_LAUNCH_CHECK("h1_from_projection_kernel");
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
