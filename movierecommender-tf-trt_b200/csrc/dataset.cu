// Dataset preparation on the device: the leave-last-two-out split of the reference
// (movierec/data_pipeline.py:190-198) and the per-user sorted, de-duplicated item lists (CSR) the negative sampler
// searches (the reference rebuilds that set with pandas for EVERY positive, data_pipeline.py:103-112).  Integer,
// HBM-bound work on top of the library's stable radix sort: bit-exact against the CPU oracle.
//   split : order = stable sort of the row numbers by user (np.argsort(kind="stable") of the oracle); in that order
//           the last row of a user is its test rating, the one before it the validation rating, the rest train.
//   CSR   : stable sort by item, then by user = rows ordered by (user, item); adjacent duplicates dropped by a
//           flag + block-count + scan + compaction pass; rowptr[u] = first compacted entry whose user is >= u.
#include "launchers.h"

namespace mr {

constexpr int kDsThreads = 256;
constexpr int kDsPerThread = 16;
constexpr int kDsTile = kDsThreads * kDsPerThread;  // elements per CTA of the flag / compaction passes

static unsigned ds_grid(int64_t work_items) {
  int64_t b = (work_items + kDsThreads - 1) / kDsThreads;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (b > cap) b = cap;
  return (unsigned)(b < 1 ? 1 : b);
}

// ids outside [0, limit) become `limit` (they sort last and are dropped); *flag |= 1 when there was one
__global__ void __launch_bounds__(kDsThreads) clamp_ids_kernel(const int32_t* __restrict__ ids, int64_t n, int32_t limit,
                                                               int32_t* __restrict__ out, int32_t* __restrict__ flag) {
  bool bad = false;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t v = __ldg(ids + i);
    const bool b = (unsigned)v >= (unsigned)limit;
    bad |= b;
    out[i] = b ? limit : v;
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

__global__ void __launch_bounds__(kDsThreads) gather_i32_kernel(const int32_t* __restrict__ src,
                                                                const int32_t* __restrict__ idx, int64_t n,
                                                                int32_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __ldg(src + __ldg(idx + i));
}

// order[e] = first[second[e]]: the composition of the two stable sorts
__global__ void __launch_bounds__(kDsThreads) compose_index_kernel(const int32_t* __restrict__ first,
                                                                   const int32_t* __restrict__ second, int64_t n,
                                                                   int32_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __ldg(first + __ldg(second + i));
}

// part[e] for the rows in user order: 2 = last of its user (test), 1 = second last (validation), 0 = train
__global__ void __launch_bounds__(kDsThreads) split_parts_kernel(const int32_t* __restrict__ su, int64_t n,
                                                                 int32_t* __restrict__ part) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int32_t u = __ldg(su + e);
    const bool last = e + 1 >= n || __ldg(su + e + 1) != u;
    const bool second = !last && (e + 2 >= n || __ldg(su + e + 2) != u);
    part[e] = last ? 2 : (second ? 1 : 0);
  }
}

// keep[e] = entry e of the (user, item)-ordered list is the first of its pair and its user is in range
__device__ __forceinline__ bool csr_keep(const int32_t* su, const int32_t* si, int64_t e, int32_t num_users) {
  const int32_t u = __ldg(su + e);
  if ((unsigned)u >= (unsigned)num_users) return false;
  if (e == 0) return true;
  return __ldg(su + e - 1) != u || __ldg(si + e - 1) != __ldg(si + e);
}

__global__ void __launch_bounds__(kDsThreads) csr_count_kernel(const int32_t* __restrict__ su, const int32_t* __restrict__ si,
                                                               int64_t n, int32_t num_users, int32_t num_items,
                                                               unsigned* __restrict__ counts) {
  __shared__ unsigned wsum[kDsThreads / 32];
  const int64_t base = (int64_t)blockIdx.x * kDsTile;
  unsigned c = 0;
  for (int q = 0; q < kDsPerThread; ++q) {
    const int64_t e = base + (int64_t)q * kDsThreads + threadIdx.x;
    if (e < n && (unsigned)__ldg(si + e) < (unsigned)num_items && csr_keep(su, si, e, num_users)) ++c;
  }
  c = (unsigned)warp_sum_int((int)c);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = 0;
    for (int w = 0; w < kDsThreads / 32; ++w) t += wsum[w];
    counts[blockIdx.x] = t;
  }
}

// exclusive scan of counts[0..nb) in place by ONE CTA (nb = n / 4096: 4,883 for the 20 M ratings of ML-20M);
// offsets are 64-bit; total -> *total_out
__global__ void __launch_bounds__(1024) scan_counts_kernel(const unsigned* __restrict__ counts, int64_t nb,
                                                           int64_t* __restrict__ offsets, int64_t* __restrict__ total_out) {
  __shared__ int64_t part[1024];
  const int tid = threadIdx.x;
  const int64_t per = (nb + 1023) / 1024, lo = tid * per, hi = lo + per < nb ? lo + per : nb;
  int64_t s = 0;
  for (int64_t i = lo; i < hi; ++i) s += counts[i];
  part[tid] = s;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {  // Hillis-Steele inclusive scan of the 1024 partial sums
    const int64_t v = tid >= off ? part[tid - off] : 0;
    __syncthreads();
    part[tid] += v;
    __syncthreads();
  }
  int64_t run = tid ? part[tid - 1] : 0;
  for (int64_t i = lo; i < hi; ++i) {
    offsets[i] = run;
    run += counts[i];
  }
  if (tid == 1023) *total_out = part[1023];
}

// compacted (user, item) pairs in order: every CTA re-derives its tile's keep flags, ranks them (warp ballots,
// elements taken thread-major so that the ranks follow the list order) and writes at its scanned offset
__global__ void __launch_bounds__(kDsThreads) csr_compact_kernel(const int32_t* __restrict__ su, const int32_t* __restrict__ si,
                                                                 int64_t n, int32_t num_users, int32_t num_items,
                                                                 const int64_t* __restrict__ offsets,
                                                                 int32_t* __restrict__ cu, int32_t* __restrict__ ci) {
  __shared__ unsigned wsum[kDsThreads / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t base = (int64_t)blockIdx.x * kDsTile;
  int64_t out = offsets[blockIdx.x];
  for (int q = 0; q < kDsPerThread; ++q) {  // rows of 256 consecutive entries: list order = (q, thread)
    const int64_t e = base + (int64_t)q * kDsThreads + threadIdx.x;
    const bool keep = e < n && (unsigned)__ldg(si + e) < (unsigned)num_items && csr_keep(su, si, e, num_users);
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) wsum[w] = __popc(m);
    __syncthreads();
    unsigned before = 0, total = 0;
    for (int k = 0; k < kDsThreads / 32; ++k) {
      const unsigned c = wsum[k];
      if (k < w) before += c;
      total += c;
    }
    if (keep) {
      const int64_t at = out + before + __popc(m & ((1u << lane) - 1));
      cu[at] = __ldg(su + e);
      ci[at] = __ldg(si + e);
    }
    out += total;
    __syncthreads();
  }
}

// rowptr[u] = first compacted entry whose user is >= u (users without entries get empty rows); rowptr[num_users] = m
__global__ void __launch_bounds__(kDsThreads) csr_rowptr_kernel(const int32_t* __restrict__ cu, const int64_t* __restrict__ m_ptr,
                                                                int32_t num_users, int64_t* __restrict__ rowptr) {
  const int64_t m = *m_ptr;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j <= m; j += (int64_t)gridDim.x * blockDim.x) {
    // entry j opens the rows (previous user, its user]; j == m closes the list: rows (last user, num_users]
    const int32_t prev = j == 0 ? -1 : __ldg(cu + j - 1);
    const int32_t cur = j == m ? num_users : __ldg(cu + j);
    for (int32_t u = prev + 1; u <= cur; ++u) rowptr[u] = j;
  }
}

// ---- dense id remapping: ids -> ranks among the distinct ids (the reference has none, SURVEY App. B-6: MovieLens
// ids are 1-based and sparse, so its tables carry rows that are never used) -------------------------------------
// first[e] = sorted entry e opens a run of equal ids (ids outside [0, limit) were clamped to `limit`: dropped)
__device__ __forceinline__ bool run_first(const int32_t* sk, int64_t e) { return e == 0 || __ldg(sk + e - 1) != __ldg(sk + e); }

__global__ void __launch_bounds__(kDsThreads) runs_count_kernel(const int32_t* __restrict__ sk, int64_t n, int32_t limit,
                                                                unsigned* __restrict__ counts) {
  __shared__ unsigned wsum[kDsThreads / 32];
  const int64_t base = (int64_t)blockIdx.x * kDsTile;
  unsigned c = 0;
  for (int q = 0; q < kDsPerThread; ++q) {
    const int64_t e = base + (int64_t)q * kDsThreads + threadIdx.x;
    if (e < n && (unsigned)__ldg(sk + e) < (unsigned)limit && run_first(sk, e)) ++c;
  }
  c = (unsigned)warp_sum_int((int)c);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = 0;
    for (int w = 0; w < kDsThreads / 32; ++w) t += wsum[w];
    counts[blockIdx.x] = t;
  }
}

// rank of every sorted entry = number of runs opened up to and including it, minus one; scattered back through
// the sort's index (dense[idx[e]] = rank), and the id of every run written at its rank
__global__ void __launch_bounds__(kDsThreads) runs_rank_kernel(const int32_t* __restrict__ sk, const int32_t* __restrict__ idx,
                                                               int64_t n, int32_t limit, const int64_t* __restrict__ offsets,
                                                               int32_t* __restrict__ dense, int32_t* __restrict__ unique) {
  __shared__ unsigned wsum[kDsThreads / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t base = (int64_t)blockIdx.x * kDsTile;
  int64_t out = offsets[blockIdx.x];
  for (int q = 0; q < kDsPerThread; ++q) {
    const int64_t e = base + (int64_t)q * kDsThreads + threadIdx.x;
    const bool ok = e < n && (unsigned)__ldg(sk + e) < (unsigned)limit;
    const bool first = ok && run_first(sk, e);
    const unsigned m = __ballot_sync(0xffffffffu, first);
    if (lane == 0) wsum[w] = __popc(m);
    __syncthreads();
    unsigned before = 0, total = 0;
    for (int k = 0; k < kDsThreads / 32; ++k) {
      const unsigned c = wsum[k];
      if (k < w) before += c;
      total += c;
    }
    if (e < n) {
      const int64_t rank = out + before + __popc(m & (0xffffffffu >> (31 - lane))) - 1;  // runs opened up to here - 1
      dense[__ldg(idx + e)] = ok ? (int32_t)rank : -1;
      if (first) unique[rank] = __ldg(sk + e);
    }
    out += total;
    __syncthreads();
  }
}

size_t remap_workspace_bytes(int64_t n) {
  if (n < 1) n = 1;
  const int64_t nb = (n + kDsTile - 1) / kDsTile;
  return sort_workspace_bytes(n) + align_up((size_t)n * 4, 256) * 3 + align_up((size_t)nb * 4, 256) +
         align_up((size_t)nb * 8, 256) + 512;
}

int launch_remap_ids(const int32_t* ids, int64_t n, int32_t limit, int32_t* dense, int32_t* unique, int64_t* num_unique,
                     int32_t* flag, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int64_t nb = (n + kDsTile - 1) / kDsTile;
  Carver cv(ws);
  int32_t* clamped = cv.take<int32_t>(n > 0 ? n : 1);
  int32_t* sk = cv.take<int32_t>(n > 0 ? n : 1);
  int32_t* idx = cv.take<int32_t>(n > 0 ? n : 1);
  unsigned* counts = cv.take<unsigned>(nb > 0 ? nb : 1);
  int64_t* offsets = cv.take<int64_t>(nb > 0 ? nb : 1);
  void* sort_ws = cv.take<char>(sort_workspace_bytes(n > 0 ? n : 1));
  if (ws_bytes < cv.off) {
    set_error("remap workspace too small: %zu < %zu", ws_bytes, cv.off);
    return MR_ERR_WORKSPACE;
  }
  if (n == 0) {
    MR_CUDA(cudaMemsetAsync(num_unique, 0, sizeof(int64_t), st));
    return MR_OK;
  }
  int bits = 1;
  while (bits < 31 && ((int64_t)1 << bits) <= limit) ++bits;
  clamp_ids_kernel<<<ds_grid(n), kDsThreads, 0, st>>>(ids, n, limit, clamped, flag);
  MR_LAUNCH_CHECK("clamp_ids_kernel");
  int rc = launch_sort_pairs(clamped, n, bits, sk, idx, sort_ws, sort_workspace_bytes(n), st);
  if (rc != MR_OK) return rc;
  runs_count_kernel<<<(unsigned)nb, kDsThreads, 0, st>>>(sk, n, limit, counts);
  MR_LAUNCH_CHECK("runs_count_kernel");
  scan_counts_kernel<<<1, 1024, 0, st>>>(counts, nb, offsets, num_unique);
  MR_LAUNCH_CHECK("scan_counts_kernel");
  runs_rank_kernel<<<(unsigned)nb, kDsThreads, 0, st>>>(sk, idx, n, limit, offsets, dense, unique);
  MR_LAUNCH_CHECK("runs_rank_kernel");
  return MR_OK;
}

size_t split_workspace_bytes(int64_t n) {
  if (n < 1) n = 1;
  return sort_workspace_bytes(n) + align_up((size_t)n * 4, 256) * 2 + 256;
}

int launch_split_last_two(const int32_t* users, int64_t n, int32_t num_users, int32_t* order, int32_t* part,
                          int32_t* flag, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (n == 0) return MR_OK;
  Carver cv(ws);
  int32_t* clamped = cv.take<int32_t>(n);
  int32_t* su = cv.take<int32_t>(n);
  void* sort_ws = cv.take<char>(sort_workspace_bytes(n));
  if (ws_bytes < cv.off) {
    set_error("split workspace too small: %zu < %zu", ws_bytes, cv.off);
    return MR_ERR_WORKSPACE;
  }
  int bits = 1;
  while (bits < 31 && ((int64_t)1 << bits) <= num_users) ++bits;  // keys run to num_users inclusive (the sentinel)
  clamp_ids_kernel<<<ds_grid(n), kDsThreads, 0, st>>>(users, n, num_users, clamped, flag);
  MR_LAUNCH_CHECK("clamp_ids_kernel");
  int rc = launch_sort_pairs(clamped, n, bits, su, order, sort_ws, sort_workspace_bytes(n), st);
  if (rc != MR_OK) return rc;
  split_parts_kernel<<<ds_grid(n), kDsThreads, 0, st>>>(su, n, part);
  MR_LAUNCH_CHECK("split_parts_kernel");
  return MR_OK;
}

size_t user_csr_workspace_bytes(int64_t n) {
  if (n < 1) n = 1;
  const int64_t nb = (n + kDsTile - 1) / kDsTile;
  return sort_workspace_bytes(n) + align_up((size_t)n * 4, 256) * 7 + align_up((size_t)nb * 4, 256) +
         align_up((size_t)nb * 8, 256) + 512;
}

int launch_build_user_csr(const int32_t* users, const int32_t* items, int64_t n, int32_t num_users, int32_t num_items,
                          int64_t* rowptr, int32_t* csr_items, int32_t* flag, void* ws, size_t ws_bytes,
                          cudaStream_t st) {
  const int64_t nb = (n + kDsTile - 1) / kDsTile;
  Carver cv(ws);
  int32_t* cl_u = cv.take<int32_t>(n);      // clamped ids
  int32_t* cl_i = cv.take<int32_t>(n);
  int32_t* k1 = cv.take<int32_t>(n);        // sorted keys of a pass (items, then users = su)
  int32_t* idx1 = cv.take<int32_t>(n);      // order by item
  int32_t* u1 = cv.take<int32_t>(n);        // users in item order; later the compacted users
  int32_t* idx2 = cv.take<int32_t>(n);      // order by user of the item-ordered list
  int32_t* si = cv.take<int32_t>(n);        // items in (user, item) order
  unsigned* counts = cv.take<unsigned>(nb > 0 ? nb : 1);
  int64_t* offsets = cv.take<int64_t>(nb > 0 ? nb : 1);
  void* sort_ws = cv.take<char>(sort_workspace_bytes(n > 0 ? n : 1));
  if (ws_bytes < cv.off) {
    set_error("user csr workspace too small: %zu < %zu", ws_bytes, cv.off);
    return MR_ERR_WORKSPACE;
  }
  if (n == 0) {
    MR_CUDA(cudaMemsetAsync(rowptr, 0, ((size_t)num_users + 1) * sizeof(int64_t), st));
    return MR_OK;
  }
  int ubits = 1, ibits = 1;
  while (ubits < 31 && ((int64_t)1 << ubits) <= num_users) ++ubits;
  while (ibits < 31 && ((int64_t)1 << ibits) <= num_items) ++ibits;
  clamp_ids_kernel<<<ds_grid(n), kDsThreads, 0, st>>>(users, n, num_users, cl_u, flag);
  MR_LAUNCH_CHECK("clamp_ids_kernel");
  clamp_ids_kernel<<<ds_grid(n), kDsThreads, 0, st>>>(items, n, num_items, cl_i, flag);
  MR_LAUNCH_CHECK("clamp_ids_kernel");
  int rc = launch_sort_pairs(cl_i, n, ibits, k1, idx1, sort_ws, sort_workspace_bytes(n), st);
  if (rc != MR_OK) return rc;
  gather_i32_kernel<<<ds_grid(n), kDsThreads, 0, st>>>(cl_u, idx1, n, u1);
  MR_LAUNCH_CHECK("gather_i32_kernel");
  int32_t* su = cl_u;  // the clamped users are no longer needed once gathered
  rc = launch_sort_pairs(u1, n, ubits, su, idx2, sort_ws, sort_workspace_bytes(n), st);
  if (rc != MR_OK) return rc;
  gather_i32_kernel<<<ds_grid(n), kDsThreads, 0, st>>>(k1, idx2, n, si);  // k1 = items in item order
  MR_LAUNCH_CHECK("gather_i32_kernel");
  csr_count_kernel<<<(unsigned)nb, kDsThreads, 0, st>>>(su, si, n, num_users, num_items, counts);
  MR_LAUNCH_CHECK("csr_count_kernel");
  scan_counts_kernel<<<1, 1024, 0, st>>>(counts, nb, offsets, rowptr + num_users);
  MR_LAUNCH_CHECK("scan_counts_kernel");
  int32_t* cu = u1;
  csr_compact_kernel<<<(unsigned)nb, kDsThreads, 0, st>>>(su, si, n, num_users, num_items, offsets, cu, csr_items);
  MR_LAUNCH_CHECK("csr_compact_kernel");
  csr_rowptr_kernel<<<ds_grid(n + 1), kDsThreads, 0, st>>>(cu, rowptr + num_users, num_users, rowptr);
  MR_LAUNCH_CHECK("csr_rowptr_kernel");
  return MR_OK;
}

}  // namespace mr
