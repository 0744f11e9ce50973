// Fused NeuMF tile kernel (fp32 SIMT, any layer sizes): gather -> concat -> Dense/ReLU tower ->
// GMF product -> head -> sigmoid -> BCE, and in TRAIN mode the whole backward pass on the same
// shared-memory tile: dense-layer gradient partials (CTA-private, reduced in a fixed order by
// dense_reduce_kernel) and per-row embedding gradients staged for the sort + segmented reduction.
// Takes over the graph of movierec/model.py:154-188 and 213-215 (and its autodiff).
#include "neumf_tile.cuh"
#include "launchers.h"

namespace mr {

constexpr int kTileThreads = 256;

struct TileParams {
  MrModel m;
  const float* Wt[MR_MAX_LAYERS];  // Wt[l] = W[l]^T, (L[l], L[l-1]) row-major (train only)
  const int32_t* users;
  const int32_t* items;
  const float* labels;
  int64_t B;
  int64_t num_tiles;
  int32_t user_div;
  float inv_batch;
  float* logits;
  float* probs;
  float* loss_partial;   // (grid)
  float* dense_partial;  // (grid, dense_stride)
  int64_t dense_stride;
  float* stage_u;        // (B, d_u + f)
  float* stage_i;        // (B, d_i + f)
  int32_t* flags;        // [0] |= 1 when an id is out of range
  int32_t ld[MR_MAX_LAYERS];
  int32_t act_off[MR_MAX_LAYERS];
  int32_t gu_off, gi_off, ldf;
  int32_t ga_off, gb_off, ldg;
  int32_t misc_off;
};

template <int TM, bool TRAIN>
__global__ void __launch_bounds__(kTileThreads) neumf_tile_kernel(const TileParams p) {
  extern __shared__ __align__(16) float smem[];
  const MrModel& m = p.m;
  const int n = m.n_layers, f = m.mf_dim;
  const int L0 = m.L[0], d_u = L0 / 2, d_i = L0 - d_u;
  const int Ln = m.L[n - 1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;

  float* act0 = smem + p.act_off[0];
  float* actL = smem + p.act_off[n - 1];
  float* gu = smem + p.gu_off;
  float* gi = smem + p.gi_off;
  int* s_user = reinterpret_cast<int*>(smem + p.misc_off);
  int* s_item = s_user + TM;
  float* s_y = reinterpret_cast<float*>(s_item + TM);
  float* s_dz = s_y + TM;
  float* s_loss = s_dz + TM;

  const bool vec_mlp = ((d_u | d_i) & 3) == 0 && ((reinterpret_cast<uintptr_t>(m.user_mlp) |
                                                    reinterpret_cast<uintptr_t>(m.item_mlp)) & 15) == 0;
  const bool vec_gmf = f > 0 && (f & 3) == 0 && ((reinterpret_cast<uintptr_t>(m.user_gmf) |
                                                  reinterpret_cast<uintptr_t>(m.item_gmf)) & 15) == 0;
  const int su = d_u + f, si = d_i + f;  // staged row widths

  float loss_acc = 0.f;
  bool first = true;
  for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * TM;
    const int valid = (int)min((int64_t)TM, p.B - row0);

    // ---- ids and labels of the tile -------------------------------------------------------
    if (tid < TM) {
      int u = -1, it = -1;
      float y = 0.f;
      if (tid < valid) {
        u = __ldg(p.users + (row0 + tid) / p.user_div);
        it = __ldg(p.items + row0 + tid);
        if ((unsigned)u >= (unsigned)m.num_users || (unsigned)it >= (unsigned)m.num_items) {
          atomicOr(p.flags, 1);  // integer flag only; the row is treated as padding
          u = -1;
          it = -1;
        }
        if (p.labels != nullptr) y = __ldg(p.labels + row0 + tid);
      }
      s_user[tid] = u;
      s_item[tid] = it;
      s_y[tid] = y;
    }
    __syncthreads();

    // ---- embedding gather: x0 = [user row | item row], 128-bit coalesced when widths allow ---
    if (vec_mlp) {
      const int q = L0 >> 2;
      for (int e = tid; e < TM * q; e += blockDim.x) {
        const int r = e / q, c = (e - r * q) << 2;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const int u = s_user[r];
        if (u >= 0)
          v = (c < d_u) ? ld_stream4(m.user_mlp + (size_t)u * d_u + c)
                        : ld_stream4(m.item_mlp + (size_t)s_item[r] * d_i + (c - d_u));
        *reinterpret_cast<float4*>(act0 + r * p.ld[0] + c) = v;
      }
    } else {
      for (int e = tid; e < TM * L0; e += blockDim.x) {
        const int r = e / L0, c = e - r * L0;
        float v = 0.f;
        const int u = s_user[r];
        if (u >= 0)
          v = (c < d_u) ? __ldg(m.user_mlp + (size_t)u * d_u + c)
                        : __ldg(m.item_mlp + (size_t)s_item[r] * d_i + (c - d_u));
        act0[r * p.ld[0] + c] = v;
      }
    }
    if (f > 0) {
      if (vec_gmf) {
        const int q = f >> 2;
        for (int e = tid; e < TM * q; e += blockDim.x) {
          const int r = e / q, c = (e - r * q) << 2;
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
          const int u = s_user[r];
          if (u >= 0) {
            a = ld_stream4(m.user_gmf + (size_t)u * f + c);
            b = ld_stream4(m.item_gmf + (size_t)s_item[r] * f + c);
          }
          *reinterpret_cast<float4*>(gu + r * p.ldf + c) = a;
          *reinterpret_cast<float4*>(gi + r * p.ldf + c) = b;
        }
      } else {
        for (int e = tid; e < TM * f; e += blockDim.x) {
          const int r = e / f, c = e - r * f;
          float a = 0.f, b = 0.f;
          const int u = s_user[r];
          if (u >= 0) {
            a = __ldg(m.user_gmf + (size_t)u * f + c);
            b = __ldg(m.item_gmf + (size_t)s_item[r] * f + c);
          }
          gu[r * p.ldf + c] = a;
          gi[r * p.ldf + c] = b;
        }
      }
    }
    __syncthreads();

    // ---- Dense/ReLU tower ---------------------------------------------------------------------
    for (int l = 1; l < n; ++l) {
      TileOut o{};
      o.smem = smem + p.act_off[l];
      o.ld = p.ld[l];
      tile_gemm<TM, true>(smem + p.act_off[l - 1], p.ld[l - 1], m.L[l - 1], m.W[l], m.L[l], m.b[l], o);
      __syncthreads();
    }

    // ---- head: z = b + w[:f].(gu*gi) + w[f:].x ; p = sigmoid(z) ; BCE ; dz ---------------------
    const int ldL = p.ld[n - 1];
    for (int r = warp; r < TM; r += nwarps) {
      float s = 0.f;
      for (int j = lane; j < f; j += 32) s = fmaf(__ldg(m.w_out + j), gu[r * p.ldf + j] * gi[r * p.ldf + j], s);
      for (int c = lane; c < Ln; c += 32) s = fmaf(__ldg(m.w_out + f + c), actL[r * ldL + c], s);
      s = warp_sum(s);
      if (lane == 0) {
        const float z = s + __ldg(m.b_out);
        const float pr = sigmoidf_stable(z);
        float dz = 0.f, lo = 0.f;
        if (r < valid && s_user[r] >= 0) {
          if (p.logits != nullptr) p.logits[row0 + r] = z;
          if (p.probs != nullptr) p.probs[row0 + r] = pr;
          if (p.labels != nullptr) {
            lo = bce_logits(z, s_y[r]);
            dz = (pr - s_y[r]) * p.inv_batch;
          }
        } else if (r < valid) {
          if (p.logits != nullptr) p.logits[row0 + r] = nanf("");
          if (p.probs != nullptr) p.probs[row0 + r] = nanf("");
        }
        s_dz[r] = dz;
        s_loss[r] = lo;
      }
    }
    __syncthreads();
    if (warp == 0) {
      float s = 0.f;
      for (int r = lane; r < TM; r += 32) s += s_loss[r];
      s = warp_sum(s);
      loss_acc += s;
    }

    if (TRAIN) {
      float* dp = p.dense_partial + (size_t)blockIdx.x * p.dense_stride;
      float* gcur = smem + p.ga_off;
      float* gnext = smem + p.gb_off;
      const int ldg = p.ldg;
      // head gradients: dw_out[j] = sum_r dz[r] * h[r][j], db_out = sum_r dz[r]
      float* d_wout = dp + (m.w_out - m.dense);
      for (int j = tid; j < f + Ln; j += blockDim.x) {
        float s = 0.f;
        if (j < f) {
          for (int r = 0; r < TM; ++r) s = fmaf(s_dz[r], gu[r * p.ldf + j] * gi[r * p.ldf + j], s);
        } else {
          for (int r = 0; r < TM; ++r) s = fmaf(s_dz[r], actL[r * ldL + (j - f)], s);
        }
        d_wout[j] = first ? s : (d_wout[j] + s);
      }
      if (tid == 0) {
        float s = 0.f;
        for (int r = 0; r < TM; ++r) s += s_dz[r];
        float* d_bout = dp + (m.b_out - m.dense);
        *d_bout = first ? s : (*d_bout + s);
      }
      // GMF row gradients go straight to the staged buffers (after the MLP part of each row)
      for (int e = tid; e < TM * f; e += blockDim.x) {
        const int r = e / f, j = e - r * f;
        if (r < valid) {
          const float g = s_dz[r] * __ldg(m.w_out + j);
          p.stage_u[(row0 + r) * (int64_t)su + d_u + j] = g * gi[r * p.ldf + j];
          p.stage_i[(row0 + r) * (int64_t)si + d_i + j] = g * gu[r * p.ldf + j];
        }
      }
      // gradient entering the last tower layer, through the ReLU when it is a hidden layer
      for (int e = tid; e < TM * Ln; e += blockDim.x) {
        const int r = e / Ln, c = e - r * Ln;
        float v = s_dz[r] * __ldg(m.w_out + f + c);
        if (n > 1) {
          if (!(actL[r * ldL + c] > 0.f)) v = 0.f;
          gcur[r * ldg + c] = v;
        } else if (r < valid) {
          if (c < d_u) p.stage_u[(row0 + r) * (int64_t)su + c] = v;
          else p.stage_i[(row0 + r) * (int64_t)si + (c - d_u)] = v;
        }
      }
      __syncthreads();
      for (int l = n - 1; l >= 1; --l) {
        tile_outer_acc<TM>(smem + p.act_off[l - 1], p.ld[l - 1], m.L[l - 1], gcur, ldg, m.L[l],
                           dp + (m.W[l] - m.dense), dp + (m.b[l] - m.dense), first);
        TileOut o{};
        if (l - 1 >= 1) {
          o.smem = gnext;
          o.ld = ldg;
          o.mask = smem + p.act_off[l - 1];
          o.mask_ld = p.ld[l - 1];
        } else {
          o.ga = p.stage_u;
          o.gb = p.stage_i;
          o.lda = su;
          o.ldb = si;
          o.split = d_u;
          o.row0 = row0;
          o.valid_rows = valid;
        }
        tile_gemm<TM, false>(gcur, ldg, m.L[l], p.Wt[l], m.L[l - 1], nullptr, o);
        __syncthreads();
        float* t = gcur;
        gcur = gnext;
        gnext = t;
      }
    }
    first = false;
    __syncthreads();
  }
  if (tid == 0 && p.loss_partial != nullptr) p.loss_partial[blockIdx.x] = loss_acc;
}

// Wt[l][n][k] = W[l][k][n] for every hidden kernel, one launch (the dense block is tiny).
__global__ void transpose_kernels_kernel(MrModel m, float* __restrict__ wt) {
  for (int l = 1; l < m.n_layers; ++l) {
    const int K = m.L[l - 1], N = m.L[l];
    const float* src = m.W[l];
    float* dst = wt + (m.W[l] - m.dense);
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < K * N; e += gridDim.x * blockDim.x) {
      const int nn = e / K, k = e - nn * K;  // consecutive threads write consecutive Wt elements
      dst[e] = __ldg(src + (size_t)k * N + nn);
    }
  }
}

// grads.dense[j] = sum over CTAs (fixed order) of the CTA-private partials, + 2*l2[l]*W[l][j].
// A CTA owns 32 consecutive outputs; warp s of its eight adds the rows s, s + 8, s + 16, ... (a serial chain over
// all rows took 56 us for the 512 rows of the default-tower step), the eight sums are added in warp order.
__global__ void __launch_bounds__(256) dense_reduce_kernel(MrModel m, const float* __restrict__ partial, int64_t stride,
                                                           int grid_ctas, float* __restrict__ out, int with_l2) {
  __shared__ float red[8][32];
  const int lane = threadIdx.x & 31, sl = threadIdx.x >> 5;
  for (int64_t j0 = (int64_t)blockIdx.x * 32; j0 < m.dense_count; j0 += (int64_t)gridDim.x * 32) {
    const int64_t j = j0 + lane;
    float s = 0.f;
    if (j < m.dense_count) {
#pragma unroll 4
      for (int c = sl; c < grid_ctas; c += 8) s += partial[(size_t)c * stride + j];
    }
    red[sl][lane] = s;
    __syncthreads();
    if (sl == 0 && j < m.dense_count) {
      float t = red[0][lane];
#pragma unroll
      for (int w = 1; w < 8; ++w) t += red[w][lane];
      for (int l = 1; with_l2 && l < m.n_layers; ++l) {
        const int64_t off = m.W[l] - m.dense;
        if (m.l2[l] != 0.f && j >= off && j < off + (int64_t)m.L[l - 1] * m.L[l]) t += 2.f * m.l2[l] * m.dense[j];
      }
      out[j] = t;
    }
    __syncthreads();
  }
}

// out[0] = sum of `n` per-CTA partials in index order (single warp; n is a few hundred).
__global__ void sum_partials_kernel(const float* __restrict__ partial, int n, float* __restrict__ out) {
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 32) s += partial[i];
  s = warp_sum(s);
  if (threadIdx.x == 0) *out = s;
}

// sum of squares of a buffer, scaled: out += coef * sum(x^2); single CTA, fixed order (used only
// when an l2 coefficient is non-zero -- the penalty term of the reported loss, model.py:163,168,178).
__global__ void l2_penalty_kernel(const float* __restrict__ x, int64_t n, float coef, float* __restrict__ out) {
  __shared__ float red[32];
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s = fmaf(x[i], x[i], s);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) *out += coef * t;
  }
}

// ---- host side -----------------------------------------------------------------------------------

struct TileLayout {
  int tm;
  size_t smem_bytes;
  TileParams p;
};

static size_t layout_for(const MrModel& m, bool train, int tm, TileParams* p) {
  int off = 0;
  for (int l = 0; l < m.n_layers; ++l) {
    p->ld[l] = pad_ld(m.L[l]);
    p->act_off[l] = off;
    off += tm * p->ld[l];
  }
  p->ldf = pad_ld(m.mf_dim > 0 ? m.mf_dim : 1);
  p->gu_off = off;
  if (m.mf_dim > 0) off += tm * p->ldf;
  p->gi_off = off;
  if (m.mf_dim > 0) off += tm * p->ldf;
  int gmax = 1;
  for (int l = 1; l < m.n_layers; ++l) gmax = m.L[l] > gmax ? m.L[l] : gmax;
  p->ldg = pad_ld(gmax);
  p->ga_off = off;
  if (train) off += tm * p->ldg;
  p->gb_off = off;
  if (train) off += tm * p->ldg;
  p->misc_off = off;
  off += 5 * tm;
  return (size_t)off * sizeof(float);
}

template <int TM, bool TRAIN>
static int launch_tile(const TileParams& p, size_t smem, int max_ctas, cudaStream_t st, int* grid_out) {
  auto kern = neumf_tile_kernel<TM, TRAIN>;
  MR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  MR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kTileThreads, smem));
  if (occ < 1) {
    set_error("fused tile kernel does not fit on an SM (TM=%d, smem=%zu)", TM, smem);
    return MR_ERR_INVALID;
  }
  int64_t grid = (int64_t)sm_count() * occ;
  if (grid > p.num_tiles) grid = p.num_tiles;
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  *grid_out = (int)grid;
  kern<<<(unsigned)grid, kTileThreads, smem, st>>>(p);
  MR_LAUNCH_CHECK("neumf_tile_kernel");
  return MR_OK;
}

int choose_tile_rows(const MrModel& m, bool train) {
  // Largest tile whose shared memory still lets two CTAs share an SM; else whatever fits.
  TileParams tmp;
  const int cands[4] = {32, 16, 8, 4};
  for (int i = 0; i < 4; ++i)
    if (layout_for(m, train, cands[i], &tmp) <= 112 * 1024) return cands[i];
  for (int i = 0; i < 4; ++i)
    if (layout_for(m, train, cands[i], &tmp) <= 226 * 1024) return cands[i];
  return 0;
}

int max_tile_ctas() { return sm_count() * 4; }

int launch_neumf_tiles(const TileLaunch& a, cudaStream_t st, int* grid_out) {
  const MrModel& m = *a.model;
  const int tm = choose_tile_rows(m, a.train);
  if (tm == 0) {
    set_error("layer widths too large for the fused tile kernel");
    return MR_ERR_INVALID;
  }
  TileParams p{};
  const size_t smem = layout_for(m, a.train, tm, &p);
  p.m = m;
  for (int l = 1; l < m.n_layers; ++l) p.Wt[l] = a.wt != nullptr ? a.wt + (m.W[l] - m.dense) : nullptr;
  p.users = a.users;
  p.items = a.items;
  p.labels = a.labels;
  p.B = a.B;
  p.num_tiles = (a.B + tm - 1) / tm;
  p.user_div = a.user_div < 1 ? 1 : a.user_div;
  p.inv_batch = a.inv_batch;
  p.logits = a.logits;
  p.probs = a.probs;
  p.loss_partial = a.loss_partial;
  p.dense_partial = a.dense_partial;
  p.dense_stride = a.dense_stride;
  p.stage_u = a.stage_u;
  p.stage_i = a.stage_i;
  p.flags = a.flags;
  if (p.num_tiles == 0) {
    *grid_out = 0;
    return MR_OK;
  }
  const int cap = max_tile_ctas();
#define MR_DISPATCH(TMV)                                                              \
  case TMV:                                                                           \
    return a.train ? launch_tile<TMV, true>(p, smem, cap, st, grid_out)               \
                   : launch_tile<TMV, false>(p, smem, cap, st, grid_out);
  switch (tm) {
    MR_DISPATCH(32)
    MR_DISPATCH(16)
    MR_DISPATCH(8)
    MR_DISPATCH(4)
  }
#undef MR_DISPATCH
  return MR_ERR_INVALID;
}

int launch_transpose_kernels(const MrModel& m, float* wt, cudaStream_t st) {
  if (m.n_layers <= 1) return MR_OK;
  transpose_kernels_kernel<<<64, 256, 0, st>>>(m, wt);
  MR_LAUNCH_CHECK("transpose_kernels_kernel");
  return MR_OK;
}

int launch_dense_reduce(const MrModel& m, const float* partial, int64_t stride, int grid_ctas, float* out,
                        cudaStream_t st, bool with_l2) {
  const int blocks = (int)((m.dense_count + 31) / 32);
  dense_reduce_kernel<<<blocks < 1 ? 1 : blocks, 256, 0, st>>>(m, partial, stride, grid_ctas, out, with_l2 ? 1 : 0);
  MR_LAUNCH_CHECK("dense_reduce_kernel");
  return MR_OK;
}

int launch_sum_partials(const float* partial, int n, float* out, cudaStream_t st) {
  sum_partials_kernel<<<1, 32, 0, st>>>(partial, n, out);
  MR_LAUNCH_CHECK("sum_partials_kernel");
  return MR_OK;
}

int launch_l2_penalty(const float* x, int64_t n, float coef, float* out, cudaStream_t st) {
  l2_penalty_kernel<<<1, 1024, 0, st>>>(x, n, coef, out);
  MR_LAUNCH_CHECK("l2_penalty_kernel");
  return MR_OK;
}

}  // namespace mr
