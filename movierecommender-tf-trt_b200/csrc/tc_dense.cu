// tcgen05 dense-layer kernel of the NeuMF tower (forward layers and backward-activation layers):
//
//     OUT[rows x N] = epilogue( A[rows x K] . Bw[N x K]^T )        3xTF32 on the tensor cores, fp32 accumulate
//
//   A    : batch rows.  A_GATHER: row r = [user_mlp[u[r]] | item_mlp[i[r]]] (the embedding gather of
//          model.py:161-172 fused into the first layer); A_DENSE: a row-major fp32 activation matrix;
//          A_PROJ: row r = relu(Pi[i[r]] + Zu[r / group]), the item-projected first layer (api.cu) computed by the
//          producers of the SECOND layer, so that H1 never exists in memory (ranking eval).
//   Bw   : a dense kernel (or its transpose), pre-split into TF32 hi/lo and pre-arranged in core-matrix
//          order by pack_weights_kernel, so one cp.async.bulk per K-chunk drops it into shared memory.
//   epilogue: EPI_BIAS_RELU  h = relu(acc + b)          forward Dense/ReLU (model.py:175-181)
//             EPI_MASK       dz = acc * (h_prev > 0)    backward through the previous ReLU
//             EPI_STAGE      rows of dL/dx0 split into the user / item staging buffers
//             EPI_HEAD_DOT   z[r] = relu(acc + b) . w     the last hidden layer folded into the output unit's dot
//                                                         product (ranking eval: thread = row, so the dot needs no
//                                                         shuffles and the layer's output never reaches HBM)
//
// Persistent, warp-specialised CTA (416 threads), 128 batch rows per tile, K streamed in chunks of 32:
//   warps 0-7  producers: every warp fills a slice of every K-chunk; the global loads of a chunk are issued
//              kTcLoadAhead chunks before it is converted (rotating register buffers):
//              load A fp32 (gather or dense), split x = hi + lo (TF32), st.shared in core-matrix layout;
//              lane 0 of warp 0 also issues the bulk copy of the weight chunk
//   warp  8    one elected thread issues tcgen05.mma: acc += Ahi.Bhi + Alo.Bhi + Ahi.Blo per K-step of 8
//   warps 9-12 epilogue: tcgen05.ld the accumulator (double-buffered in TMEM so the next tile's MMAs run
//              under this tile's epilogue), apply the epilogue, store to global
//   warp 13    L2 prefetch of the A rows two tiles ahead
// Stages are recycled with mbarriers: full[s] (8 producer arrivals + the bulk copy's bytes),
// empty[s] (tcgen05.commit), acc_full/acc_empty per TMEM buffer.
#include "launchers.h"
#include "tc_common.cuh"

namespace mr {

// Warp roles: [0, NPROD) producers, NPROD the MMA issuer, NPROD + 1 the L2 prefetch warp, then NEPI epilogue warps.
// Two shapes, 14 warps each: 8 producers + 4 epilogue warps (layers bound by their producers / the MMA operand
// fetches), or 4 producers + 8 epilogue warps for the layers ncu showed EPILOGUE-bound (the backward layer: K = 64,
// N = 128 -- producers idle 64 % of their samples, the four epilogue warps busy 85 %): a TMEM lane quarter is then
// served by two warps that take alternate 32-column blocks.
#ifndef MR_TC_LOAD_AHEAD
#define MR_TC_LOAD_AHEAD 1
#endif
constexpr int kTcLoadAhead = MR_TC_LOAD_AHEAD;    // group iterations the producers' global loads run ahead of the conversion
#ifndef MR_TC_GROUPS
#define MR_TC_GROUPS 2
#endif
constexpr int kTcGroups = MR_TC_GROUPS;           // producer groups (8 / kTcGroups warps each) taking chunks round-robin
constexpr int kTcThreads = 32 * 14;
#ifndef MR_TC_PREFETCH_AHEAD
#define MR_TC_PREFETCH_AHEAD 2
#endif
constexpr int kTcPrefetchAhead = MR_TC_PREFETCH_AHEAD;  // tiles the prefetch warp runs ahead of the MMA issuer
constexpr int kTcTileRows = 128;
constexpr int kTcKC = 32;  // K elements per pipeline stage
constexpr int kEpiLd = 36;  // floats per row of an epilogue warp's 32x32 staging tile (16-byte aligned, conflict-free)

enum { A_GATHER = 0, A_DENSE = 1, A_PROJ = 2 };
enum { EPI_BIAS_RELU = 0, EPI_MASK = 1, EPI_STAGE = 2, EPI_HEAD_DOT = 3 };

struct TcDenseParams {
  // A operand
  const float* a_dense;  // [rows x K] (A_DENSE)
  const float* proj_i;   // (A_PROJ) [num_items x K] item half of the first layer, indexed by items[]
  const float* proj_u;   // (A_PROJ) [rows / proj_div x K] user half + bias, launch-local row r reads row r / proj_div;
  int32_t proj_div;      //          with proj_ids: [proj_u_rows x K], row r reads row proj_ids[row0 + r]
  const int32_t* proj_ids;
  int32_t proj_u_rows;
  const float* user_tab; // (A_GATHER) user rows of width d_u, item rows of width K - d_u
  const float* item_tab;
  const int32_t* users;
  const int32_t* items;
  int32_t num_users, num_items, d_u, user_div, user_mul;  // row r reads users[r * user_mul / user_div]; d_u = 0 or K: one table only
  // B operand: packed [K/32 chunks][hi, lo][N x 32] core-matrix order
  const float* b_packed;
  int32_t N, K;
  int64_t rows;          // rows of this launch
  int64_t row0;          // global index of the first row (ids, staging and outputs are indexed globally)
  // epilogue
  const float* bias;     // EPI_BIAS_RELU, optional
  const float* addend;   // EPI_BIAS_RELU, optional: [rows / addend_div x N] added before the activation (launch-local rows)
  int32_t addend_div;
  int32_t relu;          // EPI_BIAS_RELU: apply the ReLU (0 = linear output)
  const float* head_w;   // EPI_HEAD_DOT: [N] output-unit weights of this layer's columns; out = [rows] partial logits
  const uint32_t* mask_bits;  // EPI_MASK: [rows x N/32] ReLU bits of the previous layer's output, launch-local rows
  uint32_t* bits_out;    // EPI_BIAS_RELU, optional: [rows x N/32] bits (h > 0) for the backward pass
  float* out;            // EPI_BIAS_RELU / EPI_MASK: [rows x N], launch-local rows, row stride out_ld floats
  int32_t out_ld;
  float* stage_u;        // EPI_STAGE: global rows, widths su / si, split at d_u
  float* stage_i;
  int32_t su, si;
  int32_t stages;        // pipeline depth
};

template <int AMODE, int EPI, int NPROD = 8, int NEPI = 4>
__global__ void __launch_bounds__(kTcThreads, 1) tc_dense_kernel(const TcDenseParams p) {
  static_assert(NPROD + NEPI + 2 == kTcThreads / 32, "14 warps");
  constexpr int kTcMmaWarp = NPROD, kTcPrefetchWarp = NPROD + 1, kTcEpiWarp0 = NPROD + 2;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full_bar[8], empty_bar[8], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_slot;
  __shared__ int tiles_started;  // written by the MMA issuer, paces the prefetch warp

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = p.N, K = p.K, S = p.stages;
  const int nchunks = K / kTcKC;
  const uint32_t a_bytes = kTcTileRows * kTcKC * 4;  // one of hi / lo
  const uint32_t b_bytes = (uint32_t)N * kTcKC * 4;
  const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
  float* epi_smem = reinterpret_cast<float*>(smem + (size_t)S * stage_bytes);  // 4 warps x 32 x kEpiLd floats
  const int64_t ntiles = (p.rows + kTcTileRows - 1) / kTcTileRows;
  uint32_t acc_cols = 32;
  while (acc_cols < (uint32_t)N) acc_cols <<= 1;  // columns per accumulator buffer (power of two)

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      tc::mbar_init(&full_bar[s], NPROD / (NPROD == 8 ? kTcGroups : 1) + 1);  // the group's producer warps + the weight copy's expect_tx
      tc::mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(&acc_full[b], 1);
      tc::mbar_init(&acc_empty[b], NEPI);
    }
    tc::mbar_init_fence();
    tiles_started = 0;
  }
  if (warp == kTcMmaWarp) tc::tmem_alloc(&tmem_slot, 2 * acc_cols);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = tmem_slot;

  if (warp < kTcMmaWarp) {
    // ================================ producers =================================================
    // All eight warps fill every chunk (a thread owns 2 rows x 2 sixteen-byte pieces of it).  The global loads
    // run kTcLoadAhead chunks ahead of the conversion in a rotating set of register buffers, so 3 x 16 KB per
    // SM are in flight and a load has three chunk periods to arrive (ncu: the producers of the first version,
    // one chunk ahead, sat in long-scoreboard stalls for more than half of their samples).
    constexpr int G = NPROD == 8 ? kTcGroups : 1, WPG = NPROD / G;  // producer groups, warps per group
    constexpr int RG = 16 / WPG;                  // 8-row groups per thread and chunk (a warp covers 128 / WPG rows)
    const int group = warp / WPG, pw = warp % WPG;
    const int rsub = lane & 7, csub = lane >> 3;  // 8 rows x 4 sixteen-byte chunks per warp instruction
    const int64_t my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t total = my_tiles * nchunks;
    // Running positions of the load stream and of the store stream (chunk in tile, tile, stage, phase) are
    // advanced incrementally: the 64-bit divisions n / nchunks, n % S of the first version kept the XU pipe
    // 80 % busy (ncu) and were most of the producers' instructions.
    int64_t ld_tile = 0, cached_tile = -1;
    int ld_c = group, st_c = group, st_stage = group % S;
    uint32_t st_phase = (uint32_t)((group / S) & 1);
    while (ld_c >= nchunks) { ld_c -= nchunks; ++ld_tile; }
    while (st_c >= nchunks) st_c -= nchunks;
    const float* src_u[RG];
    const float* src_i[RG];
    bool ok[RG];

    auto issue_loads = [&](float4(&x)[2 * RG]) {
      const int64_t tl = ld_tile;
      const int c = ld_c;
      ld_c += G;
      while (ld_c >= nchunks) { ld_c -= nchunks; ++ld_tile; }
      if (tl != cached_tile) {  // this lane's RG rows of the tile: 8*RG*pw + 8*g + rsub
        cached_tile = tl;
        const int64_t trow0 = (blockIdx.x + tl * gridDim.x) * kTcTileRows;
#pragma unroll
        for (int g = 0; g < RG; ++g) {
          const int64_t lr = trow0 + 8 * RG * pw + 8 * g + rsub;  // launch-local row
          ok[g] = lr < p.rows;
          src_u[g] = nullptr;
          src_i[g] = nullptr;
          if (AMODE == A_PROJ) {
            if (ok[g]) {
              const int it = __ldg(p.items + p.row0 + lr);
              if (p.proj_ids == nullptr) {
                src_u[g] = p.proj_u + (size_t)((uint32_t)lr / (uint32_t)p.proj_div) * K;
              } else {
                const int u = __ldg(p.proj_ids + p.row0 + lr);
                src_u[g] = (unsigned)u < (unsigned)p.proj_u_rows ? p.proj_u + (size_t)u * K : nullptr;  // bad id: zero row
              }
              src_i[g] = (unsigned)it < (unsigned)p.num_items ? p.proj_i + (size_t)it * K : nullptr;  // bad id: zero row
            }
          } else if (AMODE == A_GATHER) {
            if (ok[g]) {
              const uint32_t grow = (uint32_t)(p.row0 + lr);  // B < 2^31
              const bool has_u = p.d_u > 0, has_i = p.d_u < K;
              const int u = has_u ? __ldg(p.users + (p.user_div == 1 ? grow * (uint32_t)p.user_mul : grow / (uint32_t)p.user_div)) : 0;
              const int it = has_i ? __ldg(p.items + grow) : 0;
              if ((unsigned)u < (unsigned)p.num_users && (unsigned)it < (unsigned)p.num_items) {
                src_u[g] = p.user_tab + (size_t)u * p.d_u;
                src_i[g] = p.item_tab + (size_t)it * (K - p.d_u);
              } else {
                ok[g] = false;  // flagged by the head kernel; treated as a zero row here
              }
            }
          } else {
            src_u[g] = p.a_dense + (size_t)lr * K;
          }
        }
      }
      const int col0 = c * kTcKC;
#pragma unroll
      for (int g = 0; g < RG; ++g) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int col = col0 + 4 * (4 * h + csub);
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ok[g]) {
            if (AMODE == A_GATHER) v = (col < p.d_u) ? ldg4(src_u[g] + col) : ldg4(src_i[g] + (col - p.d_u));
            else if (AMODE == A_DENSE || src_u[g] != nullptr) v = ldg4(src_u[g] + col);  // A_PROJ: the user-side row
          }
          x[2 * g + h] = v;
        }
      }
    };
    // A_PROJ: the item-side rows of the chunk issue_loads just took (same tile cache, same columns)
    auto issue_loads_item = [&](int c, float4(&y)[2 * RG]) {
      const int col0 = c * kTcKC;
#pragma unroll
      for (int g = 0; g < RG; ++g) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int col = col0 + 4 * (4 * h + csub);
          y[2 * g + h] = (ok[g] && src_i[g] != nullptr) ? ldg4(src_i[g] + col) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    };
    auto store_chunk = [&](const float4(&x)[2 * RG]) {
      const int c = st_c;
      const int stage = st_stage;
      const uint32_t phase = st_phase;
      st_c += G;
      while (st_c >= nchunks) st_c -= nchunks;
      st_stage += G;
      while (st_stage >= S) { st_stage -= S; st_phase ^= 1; }
      tc::mbar_wait(&empty_bar[stage], phase ^ 1);
      uint8_t* st = smem + (size_t)stage * stage_bytes;
      if (pw == 0 && lane == 0) {
        tc::mbar_arrive_expect_tx(&full_bar[stage], 2 * b_bytes);
        tc::bulk_g2s(st + 2 * a_bytes, p.b_packed + (size_t)c * 2 * N * kTcKC, 2 * b_bytes, &full_bar[stage]);
      }
#pragma unroll
      for (int g = 0; g < RG; ++g) {
        const int r = 8 * RG * pw + 8 * g + rsub;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int cc = 4 * h + csub;  // 16-byte chunk 0..7 inside the 32-wide K chunk
          float4 hi, lo;
          tc::split_tf32x4(x[2 * g + h], hi, lo);
          const uint32_t off = (uint32_t)(((r >> 3) * 8 + cc) * 128 + (r & 7) * 16);
          *reinterpret_cast<float4*>(st + off) = hi;
          *reinterpret_cast<float4*>(st + a_bytes + off) = lo;
        }
      }
      tc::fence_proxy_async();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&full_bar[stage]);
    };

    constexpr int D = kTcLoadAhead, NB = kTcLoadAhead + 1;
    float4 buf[NB][2 * RG];
    const int64_t mine = total > group ? (total - group + G - 1) / G : 0;  // chunks of this group
    if (AMODE == A_PROJ) {
      // two source rows per element: both register buffers hold ONE chunk (no load-ahead; Pi and Zu are L2- and
      // L1-resident, and the other producer group's chunk overlaps this one's latency)
      for (int64_t i = 0; i < mine; ++i) {
        const int c = ld_c;
        issue_loads(buf[0]);
        issue_loads_item(c, buf[NB - 1]);
#pragma unroll
        for (int e = 0; e < 2 * RG; ++e) {
          buf[0][e].x = fmaxf(buf[0][e].x + buf[NB - 1][e].x, 0.f);
          buf[0][e].y = fmaxf(buf[0][e].y + buf[NB - 1][e].y, 0.f);
          buf[0][e].z = fmaxf(buf[0][e].z + buf[NB - 1][e].z, 0.f);
          buf[0][e].w = fmaxf(buf[0][e].w + buf[NB - 1][e].w, 0.f);
        }
        store_chunk(buf[0]);
      }
    } else {
#pragma unroll
    for (int j = 0; j < D; ++j)
      if (j < mine) issue_loads(buf[j]);
    for (int64_t i0 = 0; i0 < mine; i0 += NB) {
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const int64_t i = i0 + j;
        if (i < mine) {
          if (i + D < mine) issue_loads(buf[(j + D) % NB]);
          store_chunk(buf[j]);
        }
      }
    }
    }
  } else if (warp == kTcMmaWarp) {
    // ================================ MMA issuer ================================================
    // One elected thread runs the whole loop (waits included); see tc::elect_one for why not `lane == 0`.
    if (tc::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t idesc = tc::idesc_tf32(128, N, 0, 0);
      const uint64_t dbase = tc::smem_desc(0, 128, 1024);  // address field added per operand (smem < 256 KB: no carry)
      const uint32_t s0 = tc::smem_u32(smem);
      int64_t it = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int b = (int)(it & 1);
        *reinterpret_cast<volatile int*>(&tiles_started) = (int)it + 1;
        tc::mbar_wait(&acc_empty[b], (uint32_t)((it >> 1) & 1) ^ 1);
        tc::fence_after_sync();
        const uint32_t d_tmem = tmem_base + (uint32_t)b * acc_cols;
        for (int c = 0; c < nchunks; ++c) {
          tc::mbar_wait(&full_bar[stage], phase);
          tc::fence_after_sync();
          const uint32_t sa = s0 + (uint32_t)stage * stage_bytes;
          const uint32_t sb = sa + 2 * a_bytes;
          const uint64_t ah = dbase + (sa >> 4), al = dbase + ((sa + a_bytes) >> 4);
          const uint64_t bh = dbase + (sb >> 4), bl = dbase + ((sb + b_bytes) >> 4);
#pragma unroll
          for (int kk = 0; kk < kTcKC / 8; ++kk) {  // one K=8 step = two 128-byte core matrices = 256 bytes
            tc::mma_tf32(d_tmem, ah + 16 * kk, bh + 16 * kk, idesc, (c | kk) != 0);
            tc::mma_tf32(d_tmem, al + 16 * kk, bh + 16 * kk, idesc, 1);
            tc::mma_tf32(d_tmem, ah + 16 * kk, bl + 16 * kk, idesc, 1);
          }
          tc::mma_commit(&empty_bar[stage]);  // stage reusable once these MMAs have read it
          if (c == nchunks - 1) tc::mma_commit(&acc_full[b]);
          if (++stage == S) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == kTcPrefetchWarp) {
    // ================================ L2 prefetch ===============================================
    // The producers keep one chunk per group in flight in registers (32 KB per SM), which covers L2 latency
    // but not DRAM latency at full bandwidth.  This warp pulls the A rows of the tile kTcPrefetchAhead tiles
    // ahead of the MMA issuer into L2 (prefetch.global.L2 holds no registers), so the producers' loads hit L2.
    if (AMODE != A_PROJ) {  // (A_PROJ reads L2-resident projections)
      int64_t it = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        while (it >= (int64_t)*reinterpret_cast<volatile int*>(&tiles_started) + kTcPrefetchAhead) __nanosleep(256);
#pragma unroll 1
        for (int rr = 0; rr < kTcTileRows / 32; ++rr) {
          const int64_t lr = tile * kTcTileRows + 32 * rr + lane;
          if (lr >= p.rows) continue;
          if (AMODE == A_GATHER) {
            const uint32_t grow = (uint32_t)(p.row0 + lr);
            const bool has_u = p.d_u > 0, has_i = p.d_u < K;
            const int u = has_u ? __ldg(p.users + (p.user_div == 1 ? grow * (uint32_t)p.user_mul : grow / (uint32_t)p.user_div)) : 0;
            const int itm = has_i ? __ldg(p.items + grow) : 0;
            if ((unsigned)u < (unsigned)p.num_users && (unsigned)itm < (unsigned)p.num_items) {
              const float* pu = p.user_tab + (size_t)u * p.d_u;
              const float* pi = p.item_tab + (size_t)itm * (K - p.d_u);
              for (int c = 0; c < p.d_u; c += 32) tc::prefetch_l2(pu + c);
              for (int c = 0; c < K - p.d_u; c += 32) tc::prefetch_l2(pi + c);
            }
          } else {
            const float* pa = p.a_dense + (size_t)lr * K;
            for (int c = 0; c < K; c += 32) tc::prefetch_l2(pa + c);
          }
        }
      }
    }
  } else {
    // ================================ epilogue ==================================================
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access (the 4 epilogue warps cover 0..3)
    int64_t it = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int b = (int)(it & 1);
      const int64_t wrow0 = tile * kTcTileRows + 32 * quarter;  // launch-local first row of this warp
      uint32_t mbits[8];  // EPI_MASK: this thread's row, all N/32 words, fetched before the accumulator is ready
      if (EPI == EPI_MASK) {
#pragma unroll
        for (int q = 0; q < 8; ++q)
          mbits[q] = (q < (N >> 5) && wrow0 + lane < p.rows) ? __ldg(p.mask_bits + (size_t)(wrow0 + lane) * (N >> 5) + q) : 0u;
      }
      tc::mbar_wait(&acc_full[b], (uint32_t)((it >> 1) & 1));
      tc::fence_after_sync();
      const uint32_t taddr = tmem_base + (uint32_t)b * acc_cols + ((uint32_t)(32 * quarter) << 16);
      // 32 columns per pass: TMEM -> registers (thread = row) -> shared tile -> coalesced global stores
      // (thread-per-row stores wrote 16-byte pieces of 32 different rows per instruction and made the
      // epilogue the bottleneck of the backward layers).
      float* tile_s = epi_smem + (size_t)(warp - kTcEpiWarp0) * (32 * kEpiLd);
      const int64_t my_lr = wrow0 + lane;  // the row this thread owns while the data is thread-per-row
      const bool my_ok = my_lr < p.rows;
      // optional per-group addend row of this thread's row (the rows of a group read the same line)
      const float* arow = (EPI == EPI_BIAS_RELU && p.addend != nullptr && my_ok)
                              ? p.addend + (size_t)((uint32_t)my_lr / (uint32_t)p.addend_div) * N : nullptr;
      float zdot = 0.f;  // EPI_HEAD_DOT
      // fully unrolled over the (at most 8) 32-column blocks so that mbits[] is indexed at compile time: with a
      // run-time index the array lived in local memory and ncu showed the epilogue warps of the backward layers,
      // which bound that kernel, waiting on those loads for a third of their time
      const int ncols_epi = N;
#pragma unroll
      for (int cb = 0; cb < 8; ++cb) {
        const int c0 = 32 * cb;
        if (c0 >= ncols_epi) break;
        if (NEPI == 8 && (cb & 1) != ((warp - kTcEpiWarp0) >> 2)) continue;  // the quarter's other warp takes this block
        float v[32];
        tc::tmem_ld16(taddr + c0, v);
        tc::tmem_ld16(taddr + c0 + 16, v + 16);
        if (EPI == EPI_HEAD_DOT) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 bv = ldg4(p.bias + c0 + 4 * q), wv = ldg4(p.head_w + c0 + 4 * q);
            zdot = fmaf(fmaxf(v[4 * q + 0] + bv.x, 0.f), wv.x, zdot);
            zdot = fmaf(fmaxf(v[4 * q + 1] + bv.y, 0.f), wv.y, zdot);
            zdot = fmaf(fmaxf(v[4 * q + 2] + bv.z, 0.f), wv.z, zdot);
            zdot = fmaf(fmaxf(v[4 * q + 3] + bv.w, 0.f), wv.w, zdot);
          }
          continue;  // nothing to stage: one float per row leaves after the last column block
        }
        if (EPI == EPI_BIAS_RELU) {
          uint32_t bits = 0;
          if (p.bias != nullptr) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 bv = ldg4(p.bias + c0 + 4 * q);
              v[4 * q + 0] += bv.x; v[4 * q + 1] += bv.y; v[4 * q + 2] += bv.z; v[4 * q + 3] += bv.w;
            }
          }
          if (arow != nullptr) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 av = ldg4(arow + c0 + 4 * q);
              v[4 * q + 0] += av.x; v[4 * q + 1] += av.y; v[4 * q + 2] += av.z; v[4 * q + 3] += av.w;
            }
          }
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
          }
          if (p.bits_out != nullptr) {
#pragma unroll
            for (int i = 0; i < 32; ++i) bits |= (v[i] > 0.f ? 1u : 0u) << i;
            if (my_ok) p.bits_out[(size_t)my_lr * (N >> 5) + (c0 >> 5)] = bits;
          }
        } else if (EPI == EPI_MASK) {
          const uint32_t bits = mbits[cb];
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (!((bits >> i) & 1u)) v[i] = 0.f;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<float4*>(tile_s + lane * kEpiLd + 4 * q) =
              make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        __syncwarp();
        {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int idx = lane + 32 * j, r = idx >> 3, c = c0 + 4 * (idx & 7);
            const int64_t lr = wrow0 + r;
            if (lr >= p.rows) continue;
            const float4 x = *reinterpret_cast<const float4*>(tile_s + r * kEpiLd + 4 * (idx & 7));
            float* dst;
            if (EPI == EPI_STAGE) {
              dst = (c < p.d_u) ? p.stage_u + (size_t)(p.row0 + lr) * p.su + c
                                : p.stage_i + (size_t)(p.row0 + lr) * p.si + (c - p.d_u);
            } else {
              dst = p.out + (size_t)lr * p.out_ld + c;
            }
            *reinterpret_cast<float4*>(dst) = x;
          }
        }
        __syncwarp();
      }
      if (EPI == EPI_HEAD_DOT && my_ok) p.out[my_lr] = zdot;
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&acc_empty[b]);
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == kTcMmaWarp) tc::tmem_dealloc(tmem_base, 2 * acc_cols);
}

// Pre-split a kernel into TF32 hi/lo and lay it out for the bulk copies of tc_dense_kernel.
//   src: W (K_in x N_out) row-major (Keras layout).
//   transpose = 0: operand rows n = output unit, cols k = input unit  (forward:  Bw[n][k] = W[k][n])
//   transpose = 1: operand rows n = input unit,  cols k = output unit (backward: Bw[n][k] = W[n][k])
//   dst: [Kop/32 chunks][hi, lo][Nop x 32] floats in core-matrix order.
__global__ void pack_weights_kernel(const float* __restrict__ W, int K_in, int N_out, int transpose,
                                    float* __restrict__ dst) {
  const int Nop = transpose ? K_in : N_out, Kop = transpose ? N_out : K_in;
  const int total = Nop * Kop;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int n = e / Kop, k = e - n * Kop;
    const float x = transpose ? __ldg(W + (size_t)n * N_out + k) : __ldg(W + (size_t)k * N_out + n);
    float hi, lo;
    tc::split_tf32(x, hi, lo);
    const int c = k / kTcKC, kc = k - c * kTcKC;
    const size_t base = (size_t)c * 2 * Nop * kTcKC;
    const uint32_t off = tc::core_off_rg_major(n, kc, kTcKC / 4) / 4;
    dst[base + off] = hi;
    dst[base + (size_t)Nop * kTcKC + off] = lo;
  }
}

// ---- host ------------------------------------------------------------------------------------------

static size_t dense_stage_bytes(int N) { return (size_t)2 * kTcTileRows * kTcKC * 4 + (size_t)2 * N * kTcKC * 4; }

template <int AMODE, int EPI, int NPROD = 8, int NEPI = 4>
static int launch_dense_t(TcDenseParams p, cudaStream_t st) {
  const size_t sb = dense_stage_bytes(p.N);
  const size_t epi_bytes = (size_t)NEPI * 32 * kEpiLd * sizeof(float);
  int stages = (int)((226 * 1024 - epi_bytes - 2048) / sb);
  if (stages > 8) stages = 8;
  if (stages < 2) {
    set_error("tc dense: N=%d leaves fewer than 2 pipeline stages", p.N);
    return MR_ERR_INVALID;
  }
  p.stages = stages;
  const size_t smem = sb * stages + epi_bytes;
  auto kern = tc_dense_kernel<AMODE, EPI, NPROD, NEPI>;
  MR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t ntiles = (p.rows + kTcTileRows - 1) / kTcTileRows;
  if (ntiles == 0) return MR_OK;
  int64_t grid = sm_count();
  if (grid > ntiles) grid = ntiles;
  kern<<<(unsigned)grid, kTcThreads, smem, st>>>(p);
  MR_LAUNCH_CHECK("tc_dense_kernel");
  return MR_OK;
}

int launch_tc_dense(const TcDenseArgs& a, cudaStream_t st) {
  TcDenseParams p{};
  p.a_dense = a.a_dense;
  p.user_tab = a.user_tab;
  p.item_tab = a.item_tab;
  p.users = a.users;
  p.items = a.items;
  p.num_users = a.num_users;
  p.num_items = a.num_items;
  p.d_u = a.d_u;
  p.user_div = a.user_div < 1 ? 1 : a.user_div;
  p.user_mul = a.user_mul < 1 ? 1 : a.user_mul;
  p.proj_i = a.proj_i;
  p.proj_u = a.proj_u;
  p.proj_div = a.proj_div < 1 ? 1 : a.proj_div;
  p.proj_ids = a.proj_ids;
  p.proj_u_rows = a.proj_u_rows;
  p.addend = a.addend;
  p.addend_div = a.addend_div < 1 ? 1 : a.addend_div;
  p.relu = a.linear ? 0 : 1;
  p.b_packed = a.b_packed;
  p.N = a.N;
  p.K = a.K;
  p.rows = a.rows;
  p.row0 = a.row0;
  p.bias = a.bias;
  p.head_w = a.head_w;
  p.mask_bits = a.mask_bits;
  p.bits_out = a.bits_out;
  p.out = a.out;
  p.out_ld = a.out_ld > 0 ? a.out_ld : a.N;
  p.stage_u = a.stage_u;
  p.stage_i = a.stage_i;
  p.su = a.su;
  p.si = a.si;
  if (a.N % 32 || a.N < 32 || a.N > 256 || a.K % kTcKC || a.K < kTcKC) {
    set_error("tc dense: unsupported N=%d K=%d", a.N, a.K);
    return MR_ERR_INVALID;
  }
  if (a.gather) {
    if (a.epilogue != TC_EPI_BIAS_RELU) return MR_ERR_INVALID;
    return launch_dense_t<A_GATHER, EPI_BIAS_RELU>(p, st);
  }
  if (a.proj_i != nullptr) {  // A = relu(Pi[item] + Zu[row / proj_div])
    if (a.proj_u == nullptr || a.items == nullptr) return MR_ERR_INVALID;
    if (a.epilogue == TC_EPI_BIAS_RELU) return launch_dense_t<A_PROJ, EPI_BIAS_RELU>(p, st);
    if (a.epilogue == TC_EPI_HEAD_DOT && a.bias && a.head_w && a.out) return launch_dense_t<A_PROJ, EPI_HEAD_DOT>(p, st);
    return MR_ERR_INVALID;
  }
  switch (a.epilogue) {
    case TC_EPI_BIAS_RELU: return launch_dense_t<A_DENSE, EPI_BIAS_RELU>(p, st);
    case TC_EPI_MASK:  // epilogue-bound (K = 64, N = 128 in the ML-20M tower): 4 producer + 8 epilogue warps
      if (a.N < 64) return launch_dense_t<A_DENSE, EPI_MASK>(p, st);
      return launch_dense_t<A_DENSE, EPI_MASK, 4, 8>(p, st);
    case TC_EPI_STAGE: return launch_dense_t<A_DENSE, EPI_STAGE>(p, st);
    case TC_EPI_HEAD_DOT:
      if (a.bias == nullptr || a.head_w == nullptr || a.out == nullptr) return MR_ERR_INVALID;
      return launch_dense_t<A_DENSE, EPI_HEAD_DOT>(p, st);
  }
  return MR_ERR_INVALID;
}

int launch_pack_weights(const float* W, int K_in, int N_out, int transpose, float* dst, cudaStream_t st) {
  const int total = K_in * N_out;
  pack_weights_kernel<<<(total + 255) / 256, 256, 0, st>>>(W, K_in, N_out, transpose, dst);
  MR_LAUNCH_CHECK("pack_weights_kernel");
  return MR_OK;
}

}  // namespace mr
