// Tile-level building blocks of the fused NeuMF kernels (fp32 SIMT path, any layer sizes).
//
// A CTA owns a tile of TM batch rows whose activations live in shared memory, row-major with a
// padded leading dimension (multiple of 4 floats, so rows are 16-byte aligned).  Dense kernels are
// streamed from global memory through L1/L2 (the whole dense block is <= a few hundred KB and is
// shared by every CTA, so it stays cache-resident); nothing about a row leaves the SM between the
// gather and the staged gradient rows.
#pragma once

#include "common.cuh"

namespace mr {

__host__ __device__ inline int pad_ld(int width) { return ((width + 3) & ~3) + 4; }

// Destination of a tile GEMM: shared memory (optionally masked by a ReLU derivative) or the two
// staged row-gradient buffers in global memory (columns [0,split) -> a, [split,N) -> b).
struct TileOut {
  float* smem;        // non-null: out[r*ld + c]
  int ld;
  const float* mask;  // optional (same ld as `mask_ld`): out *= (mask[r][c] > 0)
  int mask_ld;
  float* ga;          // global: ga[(row0+r)*lda + c]           for c <  split
  float* gb;          // global: gb[(row0+r)*ldb + (c - split)] for c >= split
  int lda, ldb, split;
  int64_t row0;
  int valid_rows;     // rows >= valid_rows are not written to global
};

// out[r][c] = act( sum_k in[r][k] * W[k*N + c] + bias[c] ),  r < TM, c < N.
// Each thread owns 4x4 micro-tiles: four rows broadcast from shared memory, one 128-bit column
// slice of W per k from global.  Works for any K, N; the vector path needs N % 4 == 0.
template <int TM, bool RELU>
__device__ void tile_gemm(const float* __restrict__ in, int ld_in, int K, const float* __restrict__ W,
                          int N, const float* __restrict__ bias, const TileOut& o) {
  const int ncg = (N + 3) >> 2;
  const int ntile = (TM / 4) * ncg;
  const bool vecN = (N & 3) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0;
  const int K4 = K & ~3;
  for (int t = threadIdx.x; t < ntile; t += blockDim.x) {
    const int rg = t / ncg, cg = t - rg * ncg;
    const int r0 = rg * 4, c0 = cg * 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const float* in0 = in + r0 * ld_in;
    if (vecN) {
      const float* w = W + c0;
      int k = 0;
      for (; k < K4; k += 4) {
        float4 a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(in0 + i * ld_in + k);
#pragma unroll
        for (int q = 0; q < 4; ++q) b[q] = ldg4(w + (size_t)(k + q) * N);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float av[4] = {a[i].x, a[i].y, a[i].z, a[i].w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            acc[i][0] = fmaf(av[q], b[q].x, acc[i][0]);
            acc[i][1] = fmaf(av[q], b[q].y, acc[i][1]);
            acc[i][2] = fmaf(av[q], b[q].z, acc[i][2]);
            acc[i][3] = fmaf(av[q], b[q].w, acc[i][3]);
          }
        }
      }
      for (; k < K; ++k) {
        const float4 b = ldg4(w + (size_t)k * N);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float av = in0[i * ld_in + k];
          acc[i][0] = fmaf(av, b.x, acc[i][0]);
          acc[i][1] = fmaf(av, b.y, acc[i][1]);
          acc[i][2] = fmaf(av, b.z, acc[i][2]);
          acc[i][3] = fmaf(av, b.w, acc[i][3]);
        }
      }
    } else {
      for (int k = 0; k < K; ++k) {
        float b[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = (c0 + j < N) ? __ldg(W + (size_t)k * N + c0 + j) : 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float av = in0[i * ld_in + k];
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av, b[j], acc[i][j]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = r0 + i;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = c0 + j;
        if (c >= N) continue;
        float v = acc[i][j];
        if (bias != nullptr) v += __ldg(bias + c);
        if (RELU) v = fmaxf(v, 0.f);
        if (o.smem != nullptr) {
          if (o.mask != nullptr && !(o.mask[r * o.mask_ld + c] > 0.f)) v = 0.f;
          o.smem[r * o.ld + c] = v;
        } else if (r < o.valid_rows) {
          if (c < o.split) o.ga[(o.row0 + r) * (int64_t)o.lda + c] = v;
          else o.gb[(o.row0 + r) * (int64_t)o.ldb + (c - o.split)] = v;
        }
      }
    }
  }
}

// dW[k][n] (+)= sum_r A[r][k] * Z[r][n]  into a CTA-private global accumulator gW (K x N row-major),
// and db[n] (+)= sum_r Z[r][n].  8x4 micro-tiles; `first` overwrites instead of accumulating so the
// accumulator never needs zeroing.  Rows of the tile that are padding carry Z == 0.
template <int TM>
__device__ void tile_outer_acc(const float* __restrict__ A, int ldA, int K, const float* __restrict__ Z,
                               int ldZ, int N, float* __restrict__ gW, float* __restrict__ gb, bool first) {
  const int nng = (N + 3) >> 2;
  const int nkg = (K + 7) >> 3;
  const int ntile = nkg * nng;
  for (int t = threadIdx.x; t < ntile; t += blockDim.x) {
    const int kg = t / nng, ng = t - kg * nng;
    const int k0 = kg * 8, n0 = ng * 4;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    // padded leading dimensions guarantee that k0+7 and n0+3 stay inside the row allocation
    // only when K, N are multiples of 8 / 4; otherwise fall back to guarded scalar loads.
    const bool fast = (k0 + 8 <= K) && (n0 + 4 <= N);
    if (fast) {
#pragma unroll 4
      for (int r = 0; r < TM; ++r) {
        const float4 a0 = *reinterpret_cast<const float4*>(A + r * ldA + k0);
        const float4 a1 = *reinterpret_cast<const float4*>(A + r * ldA + k0 + 4);
        const float4 z = *reinterpret_cast<const float4*>(Z + r * ldZ + n0);
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          acc[i][0] = fmaf(av[i], z.x, acc[i][0]);
          acc[i][1] = fmaf(av[i], z.y, acc[i][1]);
          acc[i][2] = fmaf(av[i], z.z, acc[i][2]);
          acc[i][3] = fmaf(av[i], z.w, acc[i][3]);
        }
      }
    } else {
      for (int r = 0; r < TM; ++r) {
        float av[8], zv[4];
#pragma unroll
        for (int i = 0; i < 8; ++i) av[i] = (k0 + i < K) ? A[r * ldA + k0 + i] : 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) zv[j] = (n0 + j < N) ? Z[r * ldZ + n0 + j] : 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], zv[j], acc[i][j]);
      }
    }
    if (fast && (N & 3) == 0 && (reinterpret_cast<uintptr_t>(gW) & 15) == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float4* dst = reinterpret_cast<float4*>(gW + (size_t)(k0 + i) * N + n0);
        float4 v = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        if (!first) {
          const float4 old = *dst;
          v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w;
        }
        *dst = v;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (k0 + i < K && n0 + j < N) {
            float* dst = gW + (size_t)(k0 + i) * N + n0 + j;
            *dst = first ? acc[i][j] : (*dst + acc[i][j]);
          }
        }
    }
  }
  if (gb != nullptr) {
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
      float s = 0.f;
      for (int r = 0; r < TM; ++r) s += Z[r * ldZ + n];
      gb[n] = first ? s : (gb[n] + s);
    }
  }
}

}  // namespace mr
