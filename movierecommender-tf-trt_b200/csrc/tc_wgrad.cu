// tcgen05 weight-gradient kernel: dW[Fa x Fb] += A^T . Z and db[Fb] += colsum(Z), the reduction running
// over batch rows (TensorFlow's MatMul-grad / BiasAddGrad of the Dense layers, model.py:175-181).
//
// Rows are the K dimension of the MMA, so both operands are MN-major: a row-major [rows x F] chunk goes
// to shared memory as is (rows of 32 features, SWIZZLE_128B_BASE32B -- tc_common.cuh) and the tensor core
// transposes it.  A = [user row | item row] gathered by id (first layer) or an activation matrix.
// Each CTA owns a contiguous slice of 16-row chunks, keeps the whole dW in TMEM (Fa/128 accumulators of
// Fb columns) for its slice, and adds it to its private row of the partial buffer at the end; the
// cross-CTA sum happens in dense_reduce_kernel in a fixed order.
//   warps 0-7  producers (every warp fills a slice of every chunk, global loads kWgLoadAhead chunks ahead in
//              rotating register buffers): load, TF32 hi/lo split, st.shared; the threads that load Z also
//              keep column sums for db
//   warp  8    MMA issuer
//   warp  9    L2 prefetch of the rows 24 chunks ahead
//   warps 0-3  epilogue after the last chunk (TMEM -> partial buffer)
#include <stdlib.h>

#include "launchers.h"
#include "tc_common.cuh"

namespace mr {

constexpr int kWgThreads = 320;  // 8 producer warps + MMA warp + L2 prefetch warp
constexpr int kWgPrefetchWarp = 9;
// Chunks the L2 prefetch warp runs ahead of the MMA issuer.  Swept on the ML-20M step (weight-gradient phase, ms):
// 4: 0.470, 8: 0.436, 10: 0.435, 12: 0.442, 16: 0.447, 24: 0.478, 48: 0.545 -- at 24 chunks the 148 CTAs keep 87 MB
// of prefetched lines in the 126 MB L2 and lines are evicted before their load arrives (L2 hit rate 33 %, ncu).
#ifndef MR_WG_PREFETCH_AHEAD
#define MR_WG_PREFETCH_AHEAD 8
#endif
constexpr int kWgPrefetchAhead = MR_WG_PREFETCH_AHEAD;
constexpr int kWgMmaWarp = 8;
constexpr int kWgKC = 16;  // batch rows per pipeline stage
#ifndef MR_WG_LOAD_AHEAD
#define MR_WG_LOAD_AHEAD 1
#endif
constexpr int kWgLoadAhead = MR_WG_LOAD_AHEAD;  // group iterations the global loads run ahead of the conversion
#ifndef MR_WG_GROUPS
#define MR_WG_GROUPS 2
#endif
constexpr int kWgGroups = MR_WG_GROUPS;         // producer groups (8 / kWgGroups warps each) taking chunks round-robin

struct TcWgradParams {
  const float* a_dense;
  const float* user_tab;
  const float* item_tab;
  const int32_t* users;
  const int32_t* items;
  int32_t num_users, num_items, d_u, user_mul;  // d_u = 0 or Fa: one table only; user rows read users[r * user_mul]
  // AM = 2 (projected first layer): A row r = relu(proj_i[items[row0 + r]] + proj_u[proj_ids ? proj_ids[row0 + r] : r / proj_div])
  const float* proj_i;
  const float* proj_u;
  const int32_t* proj_ids;
  int32_t proj_u_rows, proj_div;
  const float* z;
  int32_t Fa, Fb;
  int64_t rows, row0;
  float* dw_partial;
  float* db_partial;
  int64_t partial_stride;
  int32_t stages;
  int32_t debug;  // MR_TC_DEBUG (diagnostics only): 2 = skip global loads, 4 = skip convert + shared stores
};

// NA / NZ: float4 loads per producer thread and chunk for A / Z (Fa / 32 and Fb / 32).
// AM: 0 = A is a dense matrix, 1 = gathered embedding rows, 2 = the projected first layer recomputed from its two
// L2-resident projections (the train step then never stores H1: api.cu).
template <int AM, int NA, int NZ>
__global__ void __launch_bounds__(kWgThreads, 1) tc_wgrad_kernel(const TcWgradParams p) {
  constexpr bool GATHER = AM == 1;
  constexpr bool PROJ = AM == 2;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[8], empty_bar[8], done_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float db_red[256 * 4];
  __shared__ int chunks_issued;  // written by the MMA issuer, paces the prefetch warp

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int Fa = 32 * NA, Fb = 32 * NZ;  // compile-time widths: index arithmetic folds, no integer divisions
  const int S = p.stages;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t a_bytes = (uint32_t)kWgKC * Fa * 4, z_bytes = (uint32_t)kWgKC * Fb * 4;
  const uint32_t stage_bytes = 2 * a_bytes + 2 * z_bytes;
  constexpr int halves = Fa / 128;
  uint32_t acc_cols = 32;
  while (acc_cols < (uint32_t)(halves * Fb)) acc_cols <<= 1;

  const int64_t total_chunks = (p.rows + kWgKC - 1) / kWgKC;
  const int64_t per_cta = (total_chunks + gridDim.x - 1) / gridDim.x;
  const int64_t chunk_lo = min(total_chunks, per_cta * blockIdx.x);
  const int64_t chunk_hi = min(total_chunks, chunk_lo + per_cta);
  const int64_t my_chunks = chunk_hi - chunk_lo;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      tc::mbar_init(&full_bar[s], 8 / kWgGroups);
      tc::mbar_init(&empty_bar[s], 1);
    }
    tc::mbar_init(&done_bar, 1);
    tc::mbar_init_fence();
    chunks_issued = 0;
  }
  if (warp == kWgMmaWarp) tc::tmem_alloc(&tmem_slot, acc_cols);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = tmem_slot;

  float4 dbacc = make_float4(0.f, 0.f, 0.f, 0.f);  // column sums of Z for this thread's fixed float4 column
  constexpr int zq = Fb >> 2;                      // float4 per Z row (8..64, divides 128)

  if (warp < 8) {
    // producers: all eight warps fill every chunk; global loads run kWgLoadAhead chunks ahead of the
    // conversion in rotating register buffers (see tc_dense.cu)
    constexpr int G = kWgGroups, TG = 256 / G;  // threads per group
    const int group = tid / TG, t = tid % TG;
    constexpr int aq = Fa >> 2;    // float4 per A row
    int st_stage = group % S;
    uint32_t st_phase = (uint32_t)((group / S) & 1);
    int64_t ld_chunk = chunk_lo + group;  // launch-local chunk index of the next load
    constexpr int PA = NA * G / 2;                      // float4 of A per thread and chunk (16 rows x Fa/4 over TG threads)
    constexpr int PZ = NZ * G >= 2 ? NZ * G / 2 : 1;    // float4 of Z per thread and chunk (Fb = 32, G = 1: half of the threads idle)

    constexpr int PB = PROJ ? PA : 1;  // PROJ: the user-side pieces ride in a second buffer until the conversion
    auto issue_loads = [&](float4(&xa)[PA], float4(&xz)[PZ], float4(&xb)[PB]) {
      const int64_t crow0 = ld_chunk * kWgKC;  // launch-local first row of the chunk
      ld_chunk += G;
#pragma unroll
      for (int i = 0; i < PA; ++i) {
        const int idx = t + TG * i;
        const int r = idx / aq, c = (idx - r * aq) << 2;
        const int64_t lr = crow0 + r;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (PROJ) {
          float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
          if (lr < p.rows) {
            const int it = __ldg(p.items + p.row0 + lr);
            if ((unsigned)it < (unsigned)p.num_items) x = ldg4(p.proj_i + (size_t)it * Fa + c);  // bad ids: zero rows
            if (p.proj_ids != nullptr) {
              const int u = __ldg(p.proj_ids + p.row0 + lr);
              if ((unsigned)u < (unsigned)p.proj_u_rows) y = ldg4(p.proj_u + (size_t)u * Fa + c);
            } else {
              y = ldg4(p.proj_u + (size_t)((uint32_t)lr / (uint32_t)p.proj_div) * Fa + c);
            }
          }
          xb[i] = y;
        } else if (lr < p.rows && !(p.debug & 2)) {
          if (GATHER) {
            if (c < p.d_u) {
              const int u = __ldg(p.users + (p.row0 + lr) * p.user_mul);
              if ((unsigned)u < (unsigned)p.num_users) x = ldg4(p.user_tab + (size_t)u * p.d_u + c);
            } else {
              const int it = __ldg(p.items + p.row0 + lr);
              if ((unsigned)it < (unsigned)p.num_items) x = ldg4(p.item_tab + (size_t)it * (Fa - p.d_u) + (c - p.d_u));
            }
          } else {
            x = ldg4(p.a_dense + (size_t)lr * Fa + c);
          }
        }
        xa[i] = x;
      }
#pragma unroll
      for (int i = 0; i < PZ; ++i) {
        const int idx = t + TG * i;
        const int r = idx / zq, c = (idx - r * zq) << 2;
        const int64_t lr = crow0 + r;
        xz[i] = (r < kWgKC && lr < p.rows && !(p.debug & 2)) ? ldg4(p.z + (size_t)lr * Fb + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto store_chunk = [&](float4(&xa)[PA], const float4(&xz)[PZ], const float4(&xb)[PB]) {
      const int stage = st_stage;
      const uint32_t phase = st_phase;
      st_stage += G;
      while (st_stage >= S) { st_stage -= S; st_phase ^= 1; }
      tc::mbar_wait(&empty_bar[stage], phase ^ 1);
      uint8_t* st = smem + (size_t)stage * stage_bytes;
      if (!(p.debug & 4)) {
#pragma unroll
        for (int i = 0; i < PA; ++i) {
          const int idx = t + TG * i;
          const int r = idx / aq, c = (idx - r * aq) << 2;
          if (PROJ) {
            xa[i].x = fmaxf(xa[i].x + xb[i].x, 0.f);
            xa[i].y = fmaxf(xa[i].y + xb[i].y, 0.f);
            xa[i].z = fmaxf(xa[i].z + xb[i].z, 0.f);
            xa[i].w = fmaxf(xa[i].w + xb[i].w, 0.f);
          }
          float4 hi, lo;
          tc::split_tf32x4(xa[i], hi, lo);
          const uint32_t off = tc::mn_off(r, c, kWgKC / 4);
          *reinterpret_cast<float4*>(st + off) = hi;
          *reinterpret_cast<float4*>(st + a_bytes + off) = lo;
        }
#pragma unroll
        for (int i = 0; i < PZ; ++i) {  // this thread always sees float4 column t % zq of Z: keep its column sums for db
          const int idx = t + TG * i;
          const int r = idx / zq, c = (idx - r * zq) << 2;
          if (r < kWgKC) {
            dbacc.x += xz[i].x;
            dbacc.y += xz[i].y;
            dbacc.z += xz[i].z;
            dbacc.w += xz[i].w;
            float4 hi, lo;
            tc::split_tf32x4(xz[i], hi, lo);
            const uint32_t off = tc::mn_off(r, c, kWgKC / 4);
            *reinterpret_cast<float4*>(st + 2 * a_bytes + off) = hi;
            *reinterpret_cast<float4*>(st + 2 * a_bytes + z_bytes + off) = lo;
          }
        }
      }
      if (!(p.debug & 16)) tc::fence_proxy_async();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&full_bar[stage]);
    };

    constexpr int D = kWgLoadAhead, NB = kWgLoadAhead + 1;
    float4 ba[NB][PA], bz[NB][PZ], bb[NB][PB];
    const int64_t mine = my_chunks > group ? (my_chunks - group + G - 1) / G : 0;  // chunks of this group
#pragma unroll
    for (int j = 0; j < D; ++j)
      if (j < mine) issue_loads(ba[j], bz[j], bb[j]);
    for (int64_t i0 = 0; i0 < mine; i0 += NB) {
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const int64_t i = i0 + j;
        if (i < mine) {
          if (i + D < mine) issue_loads(ba[(j + D) % NB], bz[(j + D) % NB], bb[(j + D) % NB]);
          store_chunk(ba[j], bz[j], bb[j]);
        }
      }
    }
  } else if (warp == kWgPrefetchWarp) {
    // ---- L2 prefetch of the rows kWgPrefetchAhead chunks ahead of the MMA issuer (see tc_dense.cu)
    if (!(p.debug & 128)) {
      const int64_t row_lo = chunk_lo * kWgKC, row_hi = min(p.rows, chunk_hi * kWgKC);
      for (int64_t r0 = row_lo; r0 < row_hi; r0 += 32) {
        const int64_t n = (r0 - row_lo) / kWgKC;  // chunk this block of 32 rows starts at
        while (n >= (int64_t)*reinterpret_cast<volatile int*>(&chunks_issued) + kWgPrefetchAhead) __nanosleep(256);
        const int64_t lr = r0 + lane;
        if (lr >= row_hi) continue;
        if (PROJ) {  // the projections are L2-resident: only Z is prefetched
        } else if (GATHER) {
          const int u = p.d_u > 0 ? __ldg(p.users + (p.row0 + lr) * p.user_mul) : 0;
          const int it = p.d_u < Fa ? __ldg(p.items + p.row0 + lr) : 0;
          if ((unsigned)u < (unsigned)p.num_users && (unsigned)it < (unsigned)p.num_items) {
            const float* pu = p.user_tab + (size_t)u * p.d_u;
            const float* pi = p.item_tab + (size_t)it * (Fa - p.d_u);
            for (int c = 0; c < p.d_u; c += 32) tc::prefetch_l2(pu + c);
            for (int c = 0; c < Fa - p.d_u; c += 32) tc::prefetch_l2(pi + c);
          }
        } else {
          const float* pa = p.a_dense + (size_t)lr * Fa;
          for (int c = 0; c < Fa; c += 32) tc::prefetch_l2(pa + c);
        }
        const float* pz = p.z + (size_t)lr * Fb;
        for (int c = 0; c < Fb; c += 32) tc::prefetch_l2(pz + c);
      }
    }
  } else {
    // ---- MMA issuer: one elected thread runs the whole loop (see tc::elect_one)
    if (tc::elect_one()) {
      constexpr int kHalves = NA / 4;  // Fa / 128
      const uint32_t idesc = tc::idesc_tf32(128, Fb, 1, 1);
      const uint32_t lbo = (kWgKC / 4) * 512, sbo = 512;
      const uint64_t dbase = tc::smem_desc(0, lbo, sbo, tc::kLayoutSw128Base32);
      const uint32_t s0 = tc::smem_u32(smem);
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t n = 0; n < my_chunks; ++n) {
        *reinterpret_cast<volatile int*>(&chunks_issued) = (int)n + 1;
        tc::mbar_wait(&full_bar[stage], phase);
        tc::fence_after_sync();
        const uint32_t sa = s0 + (uint32_t)stage * stage_bytes;
        const uint32_t sz = sa + 2 * a_bytes;
#pragma unroll
        for (int kk = 0; kk < kWgKC / 8; ++kk) {  // one K=8 step spans 1024 bytes
          const uint64_t zh = dbase + ((sz + kk * 1024) >> 4);
          const uint64_t zl = dbase + ((sz + z_bytes + kk * 1024) >> 4);
#pragma unroll
          for (int h = 0; h < kHalves; ++h) {
            // features [128h, 128h+128) of A = 4 blocks of 32, each kWgKC/4 groups of 512 bytes
            const uint32_t aoff = (uint32_t)h * 4 * lbo + kk * 1024;
            const uint64_t ah = dbase + ((sa + aoff) >> 4);
            const uint64_t al = dbase + ((sa + a_bytes + aoff) >> 4);
            const uint32_t d = tmem_base + (uint32_t)h * Fb;
            tc::mma_tf32(d, ah, zh, idesc, (n | kk) != 0);
            tc::mma_tf32(d, al, zh, idesc, 1);
            tc::mma_tf32(d, ah, zl, idesc, 1);
          }
        }
        tc::mma_commit(&empty_bar[stage]);
        if (n == my_chunks - 1) tc::mma_commit(&done_bar);
        if (++stage == S) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  }

  // ---- bias gradient: fold the 256 producer threads' column sums in a fixed order
  if (warp < 8) *reinterpret_cast<float4*>(db_red + 4 * tid) = dbacc;
  __syncthreads();
  if (tid < zq && p.db_partial != nullptr) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = tid; j < 256; j += zq) {
      const float4 v = *reinterpret_cast<float4*>(db_red + 4 * j);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    float* dst = p.db_partial + (size_t)blockIdx.x * p.partial_stride + 4 * tid;
    dst[0] += s.x;
    dst[1] += s.y;
    dst[2] += s.z;
    dst[3] += s.w;
  }

  // ---- dW: TMEM -> this CTA's row of the partial buffer (accumulated across launches)
  if (warp < 4 && my_chunks > 0) {
    tc::mbar_wait(&done_bar, 0);
    tc::fence_after_sync();
    float* base = p.dw_partial + (size_t)blockIdx.x * p.partial_stride;
    for (int h = 0; h < halves; ++h) {
      const int m = 128 * h + 32 * warp + lane;  // input-feature index = row of dW
      for (int c0 = 0; c0 < Fb; c0 += 16) {
        float v[16];
        tc::tmem_ld16(tmem_base + (uint32_t)h * Fb + ((uint32_t)(32 * warp) << 16) + c0, v);
        float* dst = base + (size_t)m * Fb + c0;
#pragma unroll
        for (int i = 0; i < 16; ++i) dst[i] += v[i];
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == kWgMmaWarp) tc::tmem_dealloc(tmem_base, acc_cols);
}

// ---- TS variant: A^T lives in tensor memory -----------------------------------------------------------------
// dW[Fa x Fb] += A^T . Z with the A operand of the MMA (M = 128 features of A, K = batch rows) read from TMEM:
// TMEM lane = feature, TMEM column = batch row of the chunk.  A producer thread therefore owns ONE feature and
// loads that column of the chunk with 4-byte loads that are coalesced across the warp (lane = consecutive
// feature), splits it into TF32 hi/lo and writes it with tcgen05.st -- no shared-memory traffic for A at all
// (it was 2/3 of the st.shared wavefronts and half of the tensor core's operand fetches), and the MMA runs at
// 64 instead of 105.6 cycles.  Z stays a shared-memory MN-major operand exactly as in the SS kernel.
//   warps 0-7  producers, two groups of four on alternate chunks (thread = TMEM lane for A, a float4 column for Z)
//   warp 8  MMA issuer, warp 9  L2 prefetch
template <bool GATHER, int NA, int NZ, int KC>
__global__ void __launch_bounds__(kWgThreads, 1) tc_wgrad_ts_kernel(const TcWgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[8], empty_bar[8], done_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float db_red[256 * 4];
  __shared__ int chunks_issued;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int Fa = 32 * NA, Fb = 32 * NZ, halves = Fa / 128, zq = Fb >> 2;
  constexpr int kACols = halves * 2 * KC;   // TMEM columns of one A stage: per half KC hi + KC lo
  constexpr uint32_t kAccCols = halves * Fb;  // accumulators first, the A ring behind them
  const int S = p.stages;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr uint32_t z_bytes = (uint32_t)KC * Fb * 4, stage_bytes = 2 * z_bytes;

  const int64_t total_chunks = (p.rows + KC - 1) / KC;
  const int64_t per_cta = (total_chunks + gridDim.x - 1) / gridDim.x;
  const int64_t chunk_lo = min(total_chunks, per_cta * blockIdx.x);
  const int64_t chunk_hi = min(total_chunks, chunk_lo + per_cta);
  const int64_t my_chunks = chunk_hi - chunk_lo;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      tc::mbar_init(&full_bar[s], 4);  // the four warps of the group that owns the chunk
      tc::mbar_init(&empty_bar[s], 1);
    }
    tc::mbar_init(&done_bar, 1);
    tc::mbar_init_fence();
    chunks_issued = 0;
  }
  if (warp == kWgMmaWarp) tc::tmem_alloc(&tmem_slot, 512);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = tmem_slot;
  float4 dbacc = make_float4(0.f, 0.f, 0.f, 0.f);

  if (warp < 8) {
    // ---- producers: two groups of four warps on alternate chunks.  A thread owns feature f of every half of A
    // (TMEM lane 32 * (warp % 4) + lane) and a fixed float4 column of Z; global loads run one group iteration
    // ahead in registers, the row ids one more (lane r of every warp keeps the ids of row r of the chunk).
    const int group = warp >> 2, f = tid & 127, t = tid & 127;
    constexpr int PZ = NZ * KC / 16;  // float4 of Z per thread and chunk: KC rows x Fb/4 over 128 threads
    int64_t ld_chunk = chunk_lo + group;   // chunk whose VALUES are loaded next
    int64_t id_chunk = chunk_lo + group;   // chunk whose IDS are loaded next
    int st_stage = group % S;
    uint32_t st_phase = (uint32_t)((group / S) & 1);
    int uid = 0, iid = 0;  // ids of row `lane` (< 16) of the chunk the next issue_loads will read
    auto load_ids = [&]() {
      if (GATHER) {
        const int64_t lr = id_chunk * KC + (lane & (KC - 1));  // KC <= 32: lane r keeps the ids of row r
        uid = 0;
        iid = 0;
        if (lr < p.rows) {
          if (p.d_u > 0) uid = __ldg(p.users + (p.row0 + lr) * p.user_mul);
          if (p.d_u < Fa) iid = __ldg(p.items + p.row0 + lr);
        }
      }
      id_chunk += 2;
    };
    auto issue_loads = [&](float(&x)[halves * KC], float4(&xz)[PZ]) {
      const int64_t crow0 = ld_chunk * KC;
      ld_chunk += 2;
      const int my_u = uid, my_i = iid;
      load_ids();  // for the chunk after this one: their latency hides behind this chunk's loads
#pragma unroll
      for (int r = 0; r < KC; ++r) {
        const int64_t lr = crow0 + r;
        const bool ok = lr < p.rows && !(p.debug & 2);
        const int u = GATHER ? __shfl_sync(0xffffffffu, my_u, r) : 0;
        const int it = GATHER ? __shfl_sync(0xffffffffu, my_i, r) : 0;
#pragma unroll
        for (int h = 0; h < halves; ++h) {
          const int c = 128 * h + f;
          float v = 0.f;
          if (ok) {
            if (GATHER) {
              if (128 * h < p.d_u) {  // d_u is a multiple of 128 here: a half lies in one table
                if ((unsigned)u < (unsigned)p.num_users) v = __ldg(p.user_tab + (size_t)u * p.d_u + c);
              } else {
                if ((unsigned)it < (unsigned)p.num_items) v = __ldg(p.item_tab + (size_t)it * (Fa - p.d_u) + (c - p.d_u));
              }
            } else {
              v = __ldg(p.a_dense + (size_t)lr * Fa + c);
            }
          }
          x[h * KC + r] = v;
        }
      }
#pragma unroll
      for (int i = 0; i < PZ; ++i) {
        const int idx = t + 128 * i;
        const int r = idx / zq, c = (idx - r * zq) << 2;
        const int64_t lr = crow0 + r;
        xz[i] = (lr < p.rows && !(p.debug & 2)) ? ldg4(p.z + (size_t)lr * Fb + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto store_chunk = [&](const float(&x)[halves * KC], const float4(&xz)[PZ]) {
      const int stage = st_stage;
      const uint32_t phase = st_phase;
      st_stage += 2;
      while (st_stage >= S) { st_stage -= S; st_phase ^= 1; }
      tc::mbar_wait(&empty_bar[stage], phase ^ 1);
      if (!(p.debug & 256)) tc::fence_after_sync();
      uint8_t* st = smem + (size_t)stage * stage_bytes;
      if (!(p.debug & 4)) {
#pragma unroll
        for (int h = 0; h < halves; ++h) {
          float hi[KC], lo[KC];
#pragma unroll
          for (int r = 0; r < KC; ++r) tc::split_tf32_fast(x[h * KC + r], hi[r], lo[r]);
          const uint32_t taddr = tmem_base + kAccCols + (uint32_t)stage * kACols + (uint32_t)h * 2 * KC + ((uint32_t)(32 * (warp & 3)) << 16);
#pragma unroll
          for (int q = 0; q < KC / 16; ++q) {
            tc::tmem_st16(taddr + 16 * q, hi + 16 * q);
            tc::tmem_st16(taddr + KC + 16 * q, lo + 16 * q);
          }
        }
#pragma unroll
        for (int i = 0; i < PZ; ++i) {
          const int idx = t + 128 * i;
          const int r = idx / zq, c = (idx - r * zq) << 2;
          dbacc.x += xz[i].x;
          dbacc.y += xz[i].y;
          dbacc.z += xz[i].z;
          dbacc.w += xz[i].w;
          float4 hi, lo;
          tc::split_tf32x4(xz[i], hi, lo);
          const uint32_t off = tc::mn_off(r, c, KC / 4);
          *reinterpret_cast<float4*>(st + off) = hi;
          *reinterpret_cast<float4*>(st + z_bytes + off) = lo;
        }
        tc::tmem_st_wait();
      }
      tc::fence_proxy_async();
      if (!(p.debug & 256)) tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&full_bar[stage]);
    };
    float xa[2][halves * KC];
    float4 bz[2][PZ];
    const int64_t mine = my_chunks > group ? (my_chunks - group + 1) / 2 : 0;  // chunks of this group
    if (mine > 0) {
      load_ids();
      issue_loads(xa[0], bz[0]);
    }
    for (int64_t i0 = 0; i0 < mine; i0 += 2) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int64_t i = i0 + j;
        if (i < mine) {
          if (i + 1 < mine && !(p.debug & 512)) issue_loads(xa[(j + 1) & 1], bz[(j + 1) & 1]);
          store_chunk(xa[j], bz[j]);
        }
      }
    }
  } else if (warp == kWgPrefetchWarp) {
    if (!(p.debug & 128)) {
      const int64_t row_lo = chunk_lo * KC, row_hi = min(p.rows, chunk_hi * KC);
      for (int64_t r0 = row_lo; r0 < row_hi; r0 += 32) {
        const int64_t n = (r0 - row_lo) / KC;
        while (n >= (int64_t)*reinterpret_cast<volatile int*>(&chunks_issued) + kWgPrefetchAhead) __nanosleep(256);
        const int64_t lr = r0 + lane;
        if (lr >= row_hi) continue;
        if (GATHER) {
          const int u = p.d_u > 0 ? __ldg(p.users + (p.row0 + lr) * p.user_mul) : 0;
          const int it = p.d_u < Fa ? __ldg(p.items + p.row0 + lr) : 0;
          if ((unsigned)u < (unsigned)p.num_users && (unsigned)it < (unsigned)p.num_items) {
            const float* pu = p.user_tab + (size_t)u * p.d_u;
            const float* pi = p.item_tab + (size_t)it * (Fa - p.d_u);
            for (int c = 0; c < p.d_u; c += 32) tc::prefetch_l2(pu + c);
            for (int c = 0; c < Fa - p.d_u; c += 32) tc::prefetch_l2(pi + c);
          }
        } else {
          const float* pa = p.a_dense + (size_t)lr * Fa;
          for (int c = 0; c < Fa; c += 32) tc::prefetch_l2(pa + c);
        }
        const float* pz = p.z + (size_t)lr * Fb;
        for (int c = 0; c < Fb; c += 32) tc::prefetch_l2(pz + c);
      }
    }
  } else {
    // ---- MMA issuer
    if (tc::elect_one()) {
      const uint32_t idesc = tc::idesc_tf32(128, Fb, 0, 1);  // A from TMEM (K-major by construction), Z MN-major
      const uint32_t lbo = (KC / 4) * 512, sbo = 512;
      const uint64_t dbase = tc::smem_desc(0, lbo, sbo, tc::kLayoutSw128Base32);
      const uint32_t s0 = tc::smem_u32(smem);
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t n = 0; n < my_chunks; ++n) {
        *reinterpret_cast<volatile int*>(&chunks_issued) = (int)n + 1;
        tc::mbar_wait(&full_bar[stage], phase);
        tc::fence_after_sync();
        const uint32_t sz = s0 + (uint32_t)stage * stage_bytes;
        const uint32_t a0 = tmem_base + kAccCols + (uint32_t)stage * kACols;
#pragma unroll
        for (int kk = 0; kk < KC / 8; ++kk) {
          const uint64_t zh = dbase + ((sz + kk * 1024) >> 4);
          const uint64_t zl = dbase + ((sz + z_bytes + kk * 1024) >> 4);
#pragma unroll
          for (int h = 0; h < halves; ++h) {
            const uint32_t ah = a0 + (uint32_t)h * 2 * KC + 8 * kk, al = ah + KC;
            const uint32_t d = tmem_base + (uint32_t)h * Fb;
            tc::mma_tf32_ts(d, ah, zh, idesc, (n | kk) != 0);
            tc::mma_tf32_ts(d, al, zh, idesc, 1);
            tc::mma_tf32_ts(d, ah, zl, idesc, 1);
          }
        }
        tc::mma_commit(&empty_bar[stage]);
        if (n == my_chunks - 1) tc::mma_commit(&done_bar);
        if (++stage == S) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  }

  // ---- bias gradient: fold the 256 producer threads' column sums in a fixed order
  if (warp < 8) *reinterpret_cast<float4*>(db_red + 4 * tid) = dbacc;
  __syncthreads();
  if (tid < zq && p.db_partial != nullptr) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = tid; j < 256; j += zq) {
      const float4 v = *reinterpret_cast<float4*>(db_red + 4 * j);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    float* dst = p.db_partial + (size_t)blockIdx.x * p.partial_stride + 4 * tid;
    dst[0] += s.x;
    dst[1] += s.y;
    dst[2] += s.z;
    dst[3] += s.w;
  }

  // ---- dW: TMEM -> this CTA's row of the partial buffer (accumulated across launches)
  if (warp < 4 && my_chunks > 0) {
    tc::mbar_wait(&done_bar, 0);
    tc::fence_after_sync();
    float* base = p.dw_partial + (size_t)blockIdx.x * p.partial_stride;
    for (int h = 0; h < halves; ++h) {
      const int m = 128 * h + 32 * warp + lane;
      for (int c0 = 0; c0 < Fb; c0 += 16) {
        float v[16];
        tc::tmem_ld16(tmem_base + (uint32_t)h * Fb + ((uint32_t)(32 * warp) << 16) + c0, v);
        float* dst = base + (size_t)m * Fb + c0;
#pragma unroll
        for (int i = 0; i < 16; ++i) dst[i] += v[i];
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == kWgMmaWarp) tc::tmem_dealloc(tmem_base, 512);
}

int tc_wgrad_grid() { return sm_count(); }

int launch_tc_wgrad(const TcWgradArgs& a, cudaStream_t st) {
  if (a.Fa % 128 || a.Fa < 128 || a.Fa > 256 || (a.Fb != 32 && a.Fb != 64 && a.Fb != 128 && a.Fb != 256) || (a.Fa / 128) * a.Fb > 512) {
    set_error("tc wgrad: unsupported Fa=%d Fb=%d", a.Fa, a.Fb);
    return MR_ERR_INVALID;
  }
  TcWgradParams p{};
  p.a_dense = a.a_dense;
  p.user_tab = a.user_tab;
  p.item_tab = a.item_tab;
  p.users = a.users;
  p.items = a.items;
  p.num_users = a.num_users;
  p.num_items = a.num_items;
  p.d_u = a.d_u;
  p.user_mul = a.user_mul < 1 ? 1 : a.user_mul;
  p.proj_i = a.proj_i;
  p.proj_u = a.proj_u;
  p.proj_ids = a.proj_ids;
  p.proj_u_rows = a.proj_u_rows;
  p.proj_div = a.proj_div < 1 ? 1 : a.proj_div;
  p.z = a.z;
  p.Fa = a.Fa;
  p.Fb = a.Fb;
  p.rows = a.rows;
  p.row0 = a.row0;
  p.dw_partial = a.dw_partial;
  p.db_partial = a.db_partial;
  p.partial_stride = a.partial_stride;
  // TS variant (A^T in tensor memory): needs the accumulators plus at least two A stages in the 512 TMEM columns
  // and a half of A to lie in one table (d_u a multiple of 128)
  const int halves_h = a.Fa / 128;
  // Opt-in (MR_WGRAD_TS=1): parity-green on B200 but slower than the SS kernel in its first form (ML-20M weight
  // gradients 1.21 ms against 1.00 ms; with every producer action switched off 0.74 ms against 0.57 ms, i.e. the
  // stage hand-off, not the data path, is what costs) -- kept for the next round, see DESIGN.md.
  const bool proj = a.proj_i != nullptr;
  if (proj && (a.Fa != 128 || a.gather || a.proj_u == nullptr || a.items == nullptr)) {
    set_error("tc wgrad: the projected A operand needs Fa = 128, item ids and both projections");
    return MR_ERR_INVALID;
  }
  const bool ts = !proj && getenv("MR_WGRAD_TS") != nullptr && halves_h * a.Fb + 2 * halves_h * 32 <= 512 &&
                  (!a.gather || a.d_u % 128 == 0);
  // TS chunks are 32 rows (12 MMAs per half and hand-off instead of 6: the issuing thread's per-chunk overhead is
  // what bounds that kernel) when the TMEM columns allow at least 3 stages, else 16
  const int ts_kc = ts && (512 - halves_h * a.Fb) / (halves_h * 64) >= 3 && getenv("MR_WGRAD_TS16") == nullptr ? 32 : 16;
  const size_t sb = ts ? (size_t)2 * ts_kc * a.Fb * 4 : (size_t)2 * kWgKC * (a.Fa + a.Fb) * 4;
  int stages = (int)((190 * 1024) / sb);
  if (ts) {
    const int tmem_stages = (512 - halves_h * a.Fb) / (halves_h * 2 * ts_kc);
    if (stages > tmem_stages) stages = tmem_stages;
  }
  if (stages > 8) stages = 8;
  if (stages < 2) {
    set_error("tc wgrad: Fa=%d Fb=%d leave fewer than 2 stages", a.Fa, a.Fb);
    return MR_ERR_INVALID;
  }
  p.stages = stages;
  {
    const char* dbg = getenv("MR_TC_DEBUG");
    p.debug = dbg ? atoi(dbg) : 0;
  }
  const size_t smem = sb * stages + 1024;
  const int grid = tc_wgrad_grid();
  const int na = a.Fa / 32, nz = a.Fb / 32;  // kWgKC * (F / 4) / 128
  int rc = MR_ERR_INVALID;
#define MR_WG_CASE(G, NA_, NZ_)                                                                          \
  if (a.gather == G && na == NA_ && nz == NZ_) {                                                         \
    auto kern = ts ? (ts_kc == 32 ? tc_wgrad_ts_kernel<G, NA_, NZ_, 32> : tc_wgrad_ts_kernel<G, NA_, NZ_, 16>)  \
                   : tc_wgrad_kernel<G, NA_, NZ_>;                                                       \
    MR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));         \
    kern<<<grid, kWgThreads, smem, st>>>(p);                                                             \
    rc = MR_OK;                                                                                          \
  }
#define MR_WG_NZ(G, NA_) MR_WG_CASE(G, NA_, 1) MR_WG_CASE(G, NA_, 2) MR_WG_CASE(G, NA_, 4) MR_WG_CASE(G, NA_, 8)
#define MR_WG_PROJ(NZ_)                                                                                  \
  if (proj && nz == NZ_) {                                                                               \
    auto kern = tc_wgrad_kernel<2, 4, NZ_>;                                                              \
    MR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));         \
    kern<<<grid, kWgThreads, smem, st>>>(p);                                                             \
    rc = MR_OK;                                                                                          \
  }
  if (proj) {
    MR_WG_PROJ(1) MR_WG_PROJ(2) MR_WG_PROJ(4) MR_WG_PROJ(8)
  } else {
    MR_WG_NZ(true, 4) MR_WG_NZ(true, 8) MR_WG_NZ(false, 4) MR_WG_NZ(false, 8)
  }
#undef MR_WG_PROJ
#undef MR_WG_NZ
#undef MR_WG_CASE
  if (rc != MR_OK) {
    set_error("tc wgrad: no kernel for Fa=%d Fb=%d", a.Fa, a.Fb);
    return rc;
  }
  MR_LAUNCH_CHECK("tc_wgrad_kernel");
  return MR_OK;
}

}  // namespace mr
