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
  const float* z;
  int32_t Fa, Fb;
  int64_t rows, row0;
  float* dw_partial;
  float* db_partial;
  int64_t partial_stride;
  int32_t stages;
};

// NA / NZ: float4 loads per producer thread and chunk for A / Z (Fa / 32 and Fb / 32).
template <bool GATHER, int NA, int NZ>
__global__ void __launch_bounds__(kWgThreads, 1) tc_wgrad_kernel(const TcWgradParams p) {
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

    auto issue_loads = [&](float4(&xa)[PA], float4(&xz)[PZ]) {
      const int64_t crow0 = ld_chunk * kWgKC;  // launch-local first row of the chunk
      ld_chunk += G;
#pragma unroll
      for (int i = 0; i < PA; ++i) {
        const int idx = t + TG * i;
        const int r = idx / aq, c = (idx - r * aq) << 2;
        const int64_t lr = crow0 + r;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (lr < p.rows) {
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
        xz[i] = (r < kWgKC && lr < p.rows) ? ldg4(p.z + (size_t)lr * Fb + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto store_chunk = [&](float4(&xa)[PA], const float4(&xz)[PZ]) {
      const int stage = st_stage;
      const uint32_t phase = st_phase;
      st_stage += G;
      while (st_stage >= S) { st_stage -= S; st_phase ^= 1; }
      tc::mbar_wait(&empty_bar[stage], phase ^ 1);
      uint8_t* st = smem + (size_t)stage * stage_bytes;
      {
#pragma unroll
        for (int i = 0; i < PA; ++i) {
          const int idx = t + TG * i;
          const int r = idx / aq, c = (idx - r * aq) << 2;
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
      tc::fence_proxy_async();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&full_bar[stage]);
    };

    constexpr int D = kWgLoadAhead, NB = kWgLoadAhead + 1;
    float4 ba[NB][PA], bz[NB][PZ];
    const int64_t mine = my_chunks > group ? (my_chunks - group + G - 1) / G : 0;  // chunks of this group
#pragma unroll
    for (int j = 0; j < D; ++j)
      if (j < mine) issue_loads(ba[j], bz[j]);
    for (int64_t i0 = 0; i0 < mine; i0 += NB) {
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const int64_t i = i0 + j;
        if (i < mine) {
          if (i + D < mine) issue_loads(ba[(j + D) % NB], bz[(j + D) % NB]);
          store_chunk(ba[j], bz[j]);
        }
      }
    }
  } else if (warp == kWgPrefetchWarp) {
    // ---- L2 prefetch of the rows kWgPrefetchAhead chunks ahead of the MMA issuer (see tc_dense.cu)
    {
      const int64_t row_lo = chunk_lo * kWgKC, row_hi = min(p.rows, chunk_hi * kWgKC);
      for (int64_t r0 = row_lo; r0 < row_hi; r0 += 32) {
        const int64_t n = (r0 - row_lo) / kWgKC;  // chunk this block of 32 rows starts at
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
  p.z = a.z;
  p.Fa = a.Fa;
  p.Fb = a.Fb;
  p.rows = a.rows;
  p.row0 = a.row0;
  p.dw_partial = a.dw_partial;
  p.db_partial = a.db_partial;
  p.partial_stride = a.partial_stride;
  const size_t sb = (size_t)2 * kWgKC * (a.Fa + a.Fb) * 4;
  int stages = (int)((190 * 1024) / sb);
  if (stages > 8) stages = 8;
  if (stages < 2) {
    set_error("tc wgrad: Fa=%d Fb=%d leave fewer than 2 stages", a.Fa, a.Fb);
    return MR_ERR_INVALID;
  }
  p.stages = stages;
  const size_t smem = sb * stages + 1024;
  const int grid = tc_wgrad_grid();
  const int na = a.Fa / 32, nz = a.Fb / 32;  // kWgKC * (F / 4) / 128
  int rc = MR_ERR_INVALID;
#define MR_WG_CASE(G, NA_, NZ_)                                                                          \
  if (a.gather == G && na == NA_ && nz == NZ_) {                                                         \
    auto kern = tc_wgrad_kernel<G, NA_, NZ_>;                                                            \
    MR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));         \
    kern<<<grid, kWgThreads, smem, st>>>(p);                                                             \
    rc = MR_OK;                                                                                          \
  }
#define MR_WG_NZ(G, NA_) MR_WG_CASE(G, NA_, 1) MR_WG_CASE(G, NA_, 2) MR_WG_CASE(G, NA_, 4) MR_WG_CASE(G, NA_, 8)
  MR_WG_NZ(true, 4) MR_WG_NZ(true, 8) MR_WG_NZ(false, 4) MR_WG_NZ(false, 8)
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
